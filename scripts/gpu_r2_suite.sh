#!/bin/bash
# Round-2 GPU pass: tests, smoke, both bench arms, attention probe, launch list, ncu captures of the kernels that
# changed. Every step is bounded by `timeout`; outputs land in gpurun_out/ (merged back by gpurun).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02a}
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -n 15 gpurun_out/${TAG}_pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?"; tail -n 2 gpurun_out/${TAG}_smoke.log
timeout -k 10 300 python scripts/attn_probe.py > gpurun_out/${TAG}_attn_probe.log 2>&1
echo "attn probe exit $?"; tail -n 1 gpurun_out/${TAG}_attn_probe.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -n 1 gpurun_out/${TAG}_bench.log | cut -c1-1500; tail -n 5 gpurun_out/${TAG}_bench.err
timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.log 2>&1
echo "bench ref exit $?"; tail -n 1 gpurun_out/${TAG}_bench_ref.log | cut -c1-600
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --global-batch 32 \
    --skip-cpu-baseline --skip-masked-rows-head --skip-extras > gpurun_out/${TAG}_ncu_launches.log 2>&1
  echo "ncu launch list exit $?"
  python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt 2>&1; head -n 30 gpurun_out/${TAG}_launches.txt
  timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'ln_bwd_kernel|adamw_kernel|ln_fwd_kernel' \
    -s 40 -c 4 -o gpurun_out/${TAG}_rowwise python bench.py --steps 1 --warmup 3 --global-batch 32 \
    --skip-cpu-baseline --skip-masked-rows-head --skip-extras > gpurun_out/${TAG}_ncu_rowwise.log 2>&1
  echo "ncu rowwise exit $?"
  ncu -i gpurun_out/${TAG}_rowwise.ncu-rep --page details > gpurun_out/${TAG}_rowwise.details.txt 2>&1
fi
