#!/bin/bash
# A/B on one multi-GPU box: gradient all-reduce on a dedicated high-priority NCCL communicator (OBT_NCCL_HIGH_PRIORITY=1,
# MLMTrainer._gradient_group) against the default-priority world group; NCCL parity tests first.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02p}; N=${2:-4}
timeout -k 10 400 python -m pytest tests/test_multigpu_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/${TAG}_pytest_multigpu.log 2>&1
echo "pytest multigpu exit $?"; tail -n 4 gpurun_out/${TAG}_pytest_multigpu.log | cut -c1-400
port=29700
for hp in 0 1 0 1; do
  port=$((port + 1))
  OBT_NCCL_HIGH_PRIORITY=$hp timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 8 --warmup 3 --skip-extras --skip-cpu-baseline \
    --skip-masked-rows-head > gpurun_out/${TAG}_bench_n${N}_hp${hp}_$port.log 2> gpurun_out/${TAG}_bench_n${N}_hp${hp}_$port.err
  echo "hp=$hp exit $?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_n${N}_hp${hp}_$port.log").read().strip().splitlines()[-1])
print("hp=$hp", d["n_gpus"], round(d["value"]), round(d["ms_per_step"],2), "exposed_ar_ms", round(d["exposed_allreduce_ms_last_step"],3), d["clocks"])
PY
done
