#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/t_gemm.log 2>&1
echo "tests exit $?"; tail -n 2 gpurun_out/t_gemm.log
L=gpurun_out/probe_gc.log; : > $L
run() { echo "--- $*" >> $L; timeout -k 5 60 python scripts/gemm_probe.py "$@" >> $L 2>&1; echo "exit $?" >> $L; }
run 2 0 0 32768 4096 1024 t
run 2 0 0 32768 1024 1024 t
run 2 1 1 4096 1024 32768 t
grep -E "^---|time|exit [1-9]" $L | paste - - | awk '{print $2,$3,$4,$5,$6,$7, $(NF-1), $NF}'
