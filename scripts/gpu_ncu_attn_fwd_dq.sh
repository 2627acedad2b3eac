#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_attention_tc_gpu.py tests/test_shapes_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/t29.log 2>&1
echo "attn tests exit $?"; tail -n 3 gpurun_out/t29.log
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline --skip-masked-rows-head > gpurun_out/bench_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches29.csv python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline --skip-masked-rows-head > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_dq -s 24 -c 1 -f -o gpurun_out/attn_dq_v8 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline --skip-masked-rows-head > gpurun_out/ncu_b.log 2>&1
echo "ncu dq exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 24 -c 1 -f -o gpurun_out/attn_fwd_v8 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline --skip-masked-rows-head > gpurun_out/ncu_a.log 2>&1
echo "ncu fwd exit $?"
