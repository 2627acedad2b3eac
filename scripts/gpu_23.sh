#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/t23.log 2>&1
echo "gemm tests exit $?"; tail -n 4 gpurun_out/t23.log
L=gpurun_out/probe23.log; : > $L
run() { echo "--- $*" >> $L; timeout -k 5 120 python scripts/gemm_probe.py "$@" >> $L 2>&1; echo "exit $?" >> $L; }
for cg in 2; do
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
