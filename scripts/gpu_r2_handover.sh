#!/bin/bash
# dS hand-over A/B pass: guard, attention tests in both modes, per-kernel ncu durations, ncu details of the new kernels,
# bench in both modes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02f}
OBT_ATTN_DS_HANDOVER=1 PROBE_REPS=3 OBT_ATTN_VARIANT=guard timeout -k 5 120 python scripts/attn_probe.py > gpurun_out/${TAG}_guard.log 2>&1
rc=$?; echo "guard hand-over probe exit $rc"; tail -n 1 gpurun_out/${TAG}_guard.log | cut -c1-400
timeout -k 10 600 python -m pytest tests/test_attention_tc_gpu.py tests/test_kernels_gpu.py tests/test_parity_holes_gpu.py \
  tests/test_shapes_gpu.py tests/test_model_gpu.py -q --timeout 120 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -n 6 gpurun_out/${TAG}_pytest.log | cut -c1-300
for m in 0 1; do
  OBT_ATTN_DS_HANDOVER=$m PROBE_REPS=3 timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_probe_ds$m.csv python scripts/attn_probe.py > gpurun_out/${TAG}_probe_ds$m.log 2>&1
  python scripts/launch_summary.py gpurun_out/${TAG}_probe_ds$m.csv 6 > gpurun_out/${TAG}_probe_ds$m.txt 2>&1
  echo "== ds hand-over=$m"; cat gpurun_out/${TAG}_probe_ds$m.txt; tail -n 1 gpurun_out/${TAG}_probe_ds$m.log | cut -c1-330
done
OBT_ATTN_DS_HANDOVER=1 PROBE_REPS=2 timeout -k 10 600 ncu --set full --clock-control none --import-source on \
  -k regex:'attn_tc_d' -s 6 -c 2 -o gpurun_out/${TAG}_attn_bwd_ds python scripts/attn_probe.py > gpurun_out/${TAG}_ncu_attn.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/${TAG}_attn_bwd_ds.ncu-rep --page details > gpurun_out/${TAG}_attn_bwd_ds.details.txt 2>&1
for m in 0 1 0 1; do
  OBT_ATTN_DS_HANDOVER=$m timeout -k 10 600 python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --skip-masked-rows-head \
    --skip-extras >> gpurun_out/${TAG}_bench_ds$m.log 2>> gpurun_out/${TAG}_bench_ds$m.err
  echo "bench ds=$m exit $?"; tail -n 1 gpurun_out/${TAG}_bench_ds$m.log | cut -c1-170
done
du -sh gpurun_out
