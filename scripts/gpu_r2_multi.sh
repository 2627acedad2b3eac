#!/bin/bash
# Multi-GPU pass on ONE 8 x B200 box: NCCL tests (2 ranks), the default bench at N = 8 (with the BASELINE config 4
# leg appended), config 4 on its own, N = 2 / 4 for the scaling table, the reference arm under torchrun.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02m}
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
run() {  # run <n> <port> <log> <args...>
  local n=$1 port=$2 log=$3; shift 3
  timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $n "$@" > gpurun_out/$log.log 2> gpurun_out/$log.err
  echo "$log exit $?"; tail -n 1 gpurun_out/$log.log | cut -c1-300
}
timeout -k 10 600 python -m pytest tests/test_multigpu_gpu.py -q --timeout 500 -p no:cacheprovider > gpurun_out/${TAG}_pytest_multigpu.log 2>&1
echo "pytest multigpu exit $?"; tail -n 4 gpurun_out/${TAG}_pytest_multigpu.log | cut -c1-400
run 8 29601 ${TAG}_bench_n8 --steps 10 --warmup 3
run 8 29602 ${TAG}_bench_large_n8 --config large --steps 3 --warmup 2 --skip-masked-rows-head
run 4 29603 ${TAG}_bench_n4 --steps 5 --warmup 3 --skip-masked-rows-head
run 2 29604 ${TAG}_bench_n2 --steps 5 --warmup 3 --skip-masked-rows-head
run 8 29605 ${TAG}_bench_ref_n8 --impl reference --steps 5 --warmup 2
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/*_bench_*n[248].log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsed", e); continue
    print(f, d.get("n_gpus"), round(d["value"]), d.get("ms_per_step"), "exposed_ar_ms", d.get("exposed_allreduce_ms_last_step"),
          "mfu", (d.get("roofline") or {}).get("step_model_flops_frac_of_2.25PF"), "large", json.dumps(d.get("large_config"))[:700])
PY
