#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/t21.log 2>&1
echo "tests exit $?"; tail -n 12 gpurun_out/t21.log
L=gpurun_out/probe21.log; : > $L
run() { echo "--- $*" >> $L; timeout -k 5 120 python scripts/gemm_probe.py "$@" >> $L 2>&1; echo "exit $?" >> $L; }
for cg in 1 3; do
run $cg 0 0 32768 3072 1024 t
run $cg 0 0 32768 1024 1024 t
run $cg 0 0 32768 4096 1024 t
run $cg 0 0 32768 1024 4096 t
run $cg 0 0 32768 65536 1024 t
run $cg 0 1 32768 1024 4096 t
run $cg 0 1 32768 4096 1024 t
run $cg 0 1 32768 1024 3072 t
run $cg 0 1 32768 1024 65536 t
run $cg 1 1 4096 1024 32768 t
run $cg 1 1 3072 1024 32768 t
run $cg 1 1 65536 1024 32768 t
done
grep -E "^---|time|exit [1-9]" $L | paste - - | awk '{print $2,$3,$4,$5,$6,$7, $(NF-1), $NF}'
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --global-batch 128 --skip-cpu-baseline > gpurun_out/bench21.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench21.log | cut -c1-250
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches21.csv python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
