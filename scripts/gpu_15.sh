#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/t15.log 2>&1
echo "tests exit $?"; tail -n 5 gpurun_out/t15.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke15.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke15.log
timeout -k 10 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full15.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_full15.log | cut -c1-400
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref15.log 2>&1; echo "ref exit $?"; tail -n 1 gpurun_out/bench_ref15.log | cut -c1-300
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches15.csv python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
