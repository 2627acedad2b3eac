#!/bin/bash
# N = 4, 2 and 1 on ONE box with the final code (scaling table on one set of GPUs).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r03c}
for n in 4 2; do
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $((29900 + n)) bench.py --gpus $n --steps 5 --warmup 3 --skip-masked-rows-head \
    > gpurun_out/${TAG}_bench_n$n.log 2> gpurun_out/${TAG}_bench_n$n.err
  echo "n$n exit $?"
done
timeout -k 10 600 python bench.py --gpus 1 --steps 3 --warmup 3 --skip-extras --skip-cpu-baseline --skip-masked-rows-head \
  > gpurun_out/${TAG}_bench_n1.log 2> gpurun_out/${TAG}_bench_n1.err
echo "n1 exit $?"
python - <<PY
import json
for n in (1, 2, 4):
    d = json.loads(open(f"gpurun_out/${TAG}_bench_n{n}.log").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "exposed_ar_ms",
          round(d["exposed_allreduce_ms_last_step"], 3), d["clocks"]["sm_mhz"])
PY
