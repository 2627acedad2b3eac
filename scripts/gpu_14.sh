#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_attention_tc_gpu.py tests/test_model_gpu.py tests/test_shapes_gpu.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/t14.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/t14.log
timeout -k 10 1200 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_full14.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_full14.log | cut -c1-200
timeout -k 10 1200 python bench.py --steps 2 --warmup 3 --dropout 0.0 --skip-cpu-baseline > gpurun_out/bench_full14_nodrop.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_full14_nodrop.log | cut -c1-200
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches14.csv python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
