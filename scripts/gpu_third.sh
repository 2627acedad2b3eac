#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_attention_tc_gpu.py -q -m gpu -p no:cacheprovider --tb=short -s > gpurun_out/test_attention_tc_gpu.log 2>&1
echo "attn_tc exit $?"; tail -n 30 gpurun_out/test_attention_tc_gpu.log
scripts/gpu_suite.sh 900
L=gpurun_out/probe3.log; : > $L
run() { echo "--- $*" >> $L; timeout -k 5 120 python scripts/gemm_probe.py "$@" >> $L 2>&1; echo "exit $?" >> $L; }
run 1 0 0 8192 4096 1024 t
run 2 0 0 8192 4096 1024 t
run 1 0 0 32768 4096 1024 t
run 2 0 0 32768 4096 1024 t
run 1 0 0 32768 1024 4096 t
run 2 0 0 32768 1024 4096 t
run 1 0 1 32768 1024 4096 t
run 2 0 1 32768 1024 4096 t
run 1 1 1 4096 1024 32768 t
run 2 1 1 4096 1024 32768 t
run 1 0 0 16384 65536 1024 t
run 2 0 0 16384 65536 1024 t
cat $L
timeout -k 10 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "bench exit $?"; tail -n 3 gpurun_out/bench_full.log
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
