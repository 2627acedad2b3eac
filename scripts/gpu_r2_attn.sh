#!/bin/bash
# Attention-backward A/B pass: tests of both layouts, the stand-alone probe for every combination, ncu (details +
# source-level stall counters) of both dQ / dK/dV variants, a short bench with each, and an ncu capture of ln_bwd.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02b}
# guard: a new kernel variant that deadlocks must cost one minute, not the whole call
OBT_ATTN_BWD_WARPS=16 PROBE_REPS=2 OBT_ATTN_VARIANT=guard16 timeout -k 5 90 python scripts/attn_probe.py > gpurun_out/${TAG}_guard16.log 2>&1
rc=$?; echo "guard 16-warp probe exit $rc"; tail -n 2 gpurun_out/${TAG}_guard16.log | cut -c1-300
if [ $rc -ne 0 ]; then export OBT_SKIP_W16=1; fi
timeout -k 10 600 python -m pytest tests/test_attention_tc_gpu.py tests/test_gemm_gpu.py::test_gemm_delta_epilogue tests/test_kernels_gpu.py tests/test_parity_holes_gpu.py \
  tests/test_shapes_gpu.py tests/test_fullscale_gpu.py::test_large_width_block_at_full_context -q --timeout 120 \
  -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -n 8 gpurun_out/${TAG}_pytest.log
VARIANTS=("8 8" "16 16" "16 8" "8 16"); WIDTHS="8 16"
if [ -n "$OBT_SKIP_W16" ]; then VARIANTS=("8 8"); WIDTHS="8"; fi
for v in "${VARIANTS[@]}"; do
  set -- $v
  OBT_ATTN_BWD_WARPS=$1 OBT_ATTN_DKV_WARPS=$2 OBT_ATTN_VARIANT="dq$1_dkv$2" timeout -k 10 300 python scripts/attn_probe.py >> gpurun_out/${TAG}_attn_probe.log 2>&1
done
cat gpurun_out/${TAG}_attn_probe.log | cut -c1-400
for w in $WIDTHS; do
  OBT_ATTN_BWD_WARPS=$w PROBE_REPS=2 timeout -k 10 600 ncu --set full --clock-control none --import-source on \
    -k regex:'attn_tc_d' -s 6 -c 2 -o gpurun_out/${TAG}_attn_bwd_w$w python scripts/attn_probe.py > gpurun_out/${TAG}_ncu_attn_w$w.log 2>&1
  echo "ncu attn w$w exit $?"
  ncu -i gpurun_out/${TAG}_attn_bwd_w$w.ncu-rep --page details > gpurun_out/${TAG}_attn_bwd_w$w.details.txt 2>&1
  ncu -i gpurun_out/${TAG}_attn_bwd_w$w.ncu-rep --page source --csv > gpurun_out/${TAG}_attn_bwd_w$w.source.csv 2>&1
  ls -la gpurun_out/${TAG}_attn_bwd_w$w.*
done
for w in $WIDTHS; do
  OBT_ATTN_BWD_WARPS=$w timeout -k 10 600 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-masked-rows-head \
    --skip-extras > gpurun_out/${TAG}_bench_w$w.log 2> gpurun_out/${TAG}_bench_w$w.err
  echo "bench w$w exit $?"; tail -n 1 gpurun_out/${TAG}_bench_w$w.log | cut -c1-200
done
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'ln_bwd_kernel' -s 20 -c 2 \
  -o gpurun_out/${TAG}_ln_bwd python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline \
  --skip-masked-rows-head --skip-extras > gpurun_out/${TAG}_ncu_ln_bwd.log 2>&1
echo "ncu ln_bwd exit $?"
ncu -i gpurun_out/${TAG}_ln_bwd.ncu-rep --page details > gpurun_out/${TAG}_ln_bwd.details.txt 2>&1
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --global-batch 32 \
  --skip-cpu-baseline --skip-masked-rows-head --skip-extras > gpurun_out/${TAG}_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt 2>&1; head -n 16 gpurun_out/${TAG}_launches.txt
du -sh gpurun_out
