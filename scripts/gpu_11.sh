#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/t11_gemm.log 2>&1
echo "gemm tests exit $?"; tail -n 4 gpurun_out/t11_gemm.log
timeout -k 10 900 python -m pytest tests/test_attention_tc_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_shapes_gpu.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/t11.log 2>&1
echo "tests exit $?"; tail -n 5 gpurun_out/t11.log
L=gpurun_out/probe11.log; : > $L
run() { echo "--- $*" >> $L; timeout -k 5 120 python scripts/gemm_probe.py "$@" >> $L 2>&1; echo "exit $?" >> $L; }
run 3 0 0 32768 4096 1024 t
run 1 0 0 32768 4096 1024 t
run 3 0 0 32768 1024 4096 t
run 3 0 1 32768 1024 4096 t
run 3 1 1 4096 1024 32768 t
run 3 0 0 32768 65536 1024 t
run 3 0 0 32768 1024 1024 t
run 1 0 0 32768 1024 1024 t
grep -E "^---|time|cuBLAS|bad=|exit [1-9]" $L
timeout -k 10 1200 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_full11.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_full11.log | cut -c1-300
timeout -k 10 1200 python bench.py --steps 2 --warmup 3 --dropout 0.0 --skip-cpu-baseline > gpurun_out/bench_full11_nodrop.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_full11_nodrop.log | cut -c1-300
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches11.csv python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
