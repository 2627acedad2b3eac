#!/bin/bash
# GEMM epilogue study on the block's K = 1024 / N = 1024 shape (attention c_proj): every epilogue stand-alone, then one
# `ncu --set full` capture with source of the residual-dropout epilogue. Also: keep-mask generator with interval skip
# (tests + attention probe launch list).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02n}
timeout -k 10 600 python -m pytest tests/test_attention_tc_gpu.py tests/test_parity_holes_gpu.py tests/test_gemm_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -n 5 gpurun_out/${TAG}_pytest.log | cut -c1-300
for shape in "32768 1024 1024" "32768 1024 4096" "32768 4096 1024"; do
  timeout -k 10 120 python scripts/epi_probe.py $shape >> gpurun_out/${TAG}_epi_probe.txt 2>&1
done
cat gpurun_out/${TAG}_epi_probe.txt
PROBE_REPS=3 timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_probe.csv python scripts/attn_probe.py > gpurun_out/${TAG}_probe.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_probe.csv 6 > gpurun_out/${TAG}_probe.txt 2>&1; cat gpurun_out/${TAG}_probe.txt
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -f \
  -o gpurun_out/${TAG}_gemm_epi5 python scripts/epi_probe.py 32768 1024 1024 resid_dropout > gpurun_out/${TAG}_ncu_epi5.log 2>&1
echo "ncu epi5 exit $?"
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -f \
  -o gpurun_out/${TAG}_gemm_epi0 python scripts/epi_probe.py 32768 1024 1024 plain > gpurun_out/${TAG}_ncu_epi0.log 2>&1
echo "ncu epi0 exit $?"
ls -la gpurun_out | grep ${TAG}
