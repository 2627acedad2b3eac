#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_attention_tc_gpu.py tests/test_shapes_gpu.py -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/t19a.log 2>&1
echo "attn tests exit $?"; tail -n 15 gpurun_out/t19a.log
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches19.csv python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
