#!/bin/bash
# Persistent attention-backward A/B pass: guard (a deadlocking variant must cost minutes, not the call), tests, probe of
# each kernel combination, per-kernel ncu durations, ncu details of the persistent kernels, bench with and without.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02d}
OBT_ATTN_PERSIST=3 PROBE_REPS=3 OBT_ATTN_VARIANT=guard timeout -k 5 120 python scripts/attn_probe.py > gpurun_out/${TAG}_guard.log 2>&1
rc=$?; echo "guard persistent probe exit $rc"; tail -n 1 gpurun_out/${TAG}_guard.log | cut -c1-500
if [ $rc -ne 0 ]; then
  for m in 1 2; do
    OBT_ATTN_PERSIST=$m PROBE_REPS=3 OBT_ATTN_VARIANT=guard$m timeout -k 5 120 python scripts/attn_probe.py > gpurun_out/${TAG}_guard$m.log 2>&1
    echo "guard persist=$m exit $?"; tail -n 1 gpurun_out/${TAG}_guard$m.log | cut -c1-500
  done
  export OBT_ATTN_PERSIST=0
fi
timeout -k 10 900 python -m pytest tests/test_attention_tc_gpu.py tests/test_gemm_gpu.py::test_gemm_delta_epilogue tests/test_kernels_gpu.py \
  tests/test_parity_holes_gpu.py tests/test_shapes_gpu.py -q --timeout 120 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -n 6 gpurun_out/${TAG}_pytest.log | cut -c1-300
if [ $rc -eq 0 ]; then
  for m in 0 1 2 3; do
    OBT_ATTN_PERSIST=$m OBT_ATTN_VARIANT="persist$m" PROBE_REPS=50 timeout -k 10 300 python scripts/attn_probe.py 2>/dev/null | tail -n 1 >> gpurun_out/${TAG}_attn_probe.log
  done
  cut -c1-420 gpurun_out/${TAG}_attn_probe.log
  for m in 0 3; do
    OBT_ATTN_PERSIST=$m PROBE_REPS=3 timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/${TAG}_probe_launches_p$m.csv python scripts/attn_probe.py > /dev/null 2>&1
    python scripts/launch_summary.py gpurun_out/${TAG}_probe_launches_p$m.csv 8 > gpurun_out/${TAG}_probe_launches_p$m.txt 2>&1
    echo "== persist=$m"; cat gpurun_out/${TAG}_probe_launches_p$m.txt
  done
  OBT_ATTN_PERSIST=3 PROBE_REPS=2 timeout -k 10 600 ncu --set full --clock-control none --import-source on \
    -k regex:'attn_tc_d' -s 6 -c 2 -o gpurun_out/${TAG}_attn_bwd_persist python scripts/attn_probe.py > gpurun_out/${TAG}_ncu_attn.log 2>&1
  echo "ncu persist exit $?"
  ncu -i gpurun_out/${TAG}_attn_bwd_persist.ncu-rep --page details > gpurun_out/${TAG}_attn_bwd_persist.details.txt 2>&1
  ncu -i gpurun_out/${TAG}_attn_bwd_persist.ncu-rep --page source --csv > gpurun_out/${TAG}_attn_bwd_persist.source.csv 2>&1
fi
for m in 0 3; do
  if [ $rc -ne 0 ] && [ $m -eq 3 ]; then continue; fi
  OBT_ATTN_PERSIST=$m timeout -k 10 600 python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --skip-masked-rows-head \
    --skip-extras > gpurun_out/${TAG}_bench_p$m.log 2> gpurun_out/${TAG}_bench_p$m.err
  echo "bench persist=$m exit $?"; tail -n 1 gpurun_out/${TAG}_bench_p$m.log | cut -c1-200
done
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --global-batch 32 \
  --skip-cpu-baseline --skip-masked-rows-head --skip-extras > gpurun_out/${TAG}_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt 2>&1; head -n 18 gpurun_out/${TAG}_launches.txt
du -sh gpurun_out
