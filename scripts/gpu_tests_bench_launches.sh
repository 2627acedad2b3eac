#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/t_all.log 2>&1
echo "tests exit $?"; tail -n 12 gpurun_out/t_all.log
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --global-batch 128 --skip-cpu-baseline > gpurun_out/bench_gb128.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_gb128.log | cut -c1-250
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline --skip-masked-rows-head > gpurun_out/bench_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline --skip-masked-rows-head > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
