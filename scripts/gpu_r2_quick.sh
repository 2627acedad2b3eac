#!/bin/bash
# Short 1-GPU check after a small kernel change: the attention / model test modules and the attention probe launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02t}
timeout -k 10 600 python -m pytest tests/test_attention_tc_gpu.py tests/test_model_gpu.py tests/test_parity_holes_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/${TAG}_pytest.log | cut -c1-300
PROBE_REPS=3 timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_probe.csv python scripts/attn_probe.py > gpurun_out/${TAG}_probe.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_probe.csv 6 > gpurun_out/${TAG}_probe.txt 2>&1; cat gpurun_out/${TAG}_probe.txt
