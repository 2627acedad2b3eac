#!/bin/bash
# One-GPU regression + measurement pass after a kernel change: whole GPU test suite, attention probe launch list,
# bench (with the per-shape GEMM table), ncu launch list of one step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02g}
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -n 5 gpurun_out/${TAG}_pytest.log | cut -c1-300
PROBE_REPS=3 timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_probe.csv python scripts/attn_probe.py > gpurun_out/${TAG}_probe.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_probe.csv 6 > gpurun_out/${TAG}_probe.txt 2>&1; cat gpurun_out/${TAG}_probe.txt
timeout -k 10 900 python bench.py --steps 6 --warmup 3 --skip-extras > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "mfu", round(r["step_model_flops_frac_of_2.25PF"],4), "gemm", round(r["achieved"],1), "share", round(r["share_of_step"],3), "clk", d["clocks"])
for k,v in r["by_shape"].items(): print(f"  {k:36s} {v}")
print("cpu", d.get("cpu_baseline"))
PY
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --global-batch 32 \
  --skip-cpu-baseline --skip-masked-rows-head --skip-extras > gpurun_out/${TAG}_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt 2>&1; head -n 22 gpurun_out/${TAG}_launches.txt
