#!/bin/bash
# The driver's N = 8 invocation (default arguments, BASELINE config 4 leg appended) on one 8 x B200 box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02v}
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
  --master-port 29811 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n8.log 2> gpurun_out/${TAG}_bench_n8.err
echo "n8 exit $?"; tail -n 3 gpurun_out/${TAG}_bench_n8.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_n8.log").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "exposed_ar_ms", d["exposed_allreduce_ms_last_step"], d["clocks"], "mfu", d["roofline"]["step_model_flops_frac_of_2.25PF"])
print("large", json.dumps(d.get("large_config"))[:900])
print("masked_rows", d.get("masked_rows_head",{}).get("value"))
PY
