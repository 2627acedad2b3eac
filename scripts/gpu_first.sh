#!/bin/bash
# First contact with the GPU: isolated GEMM probes, then the pytest files.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/probe.log; : > $L
run() { echo "--- $*" >> $L; timeout -k 5 90 python scripts/gemm_probe.py "$@" >> $L 2>&1; echo "exit $?" >> $L; }
run 1 0 0 128 256 64
run 1 0 0 256 512 256
run 1 0 1 256 512 256
run 1 1 1 256 512 256
run 1 1 0 256 512 256
run 2 0 0 256 256 64
run 2 0 0 512 512 256
run 2 0 1 512 512 256
run 2 1 1 512 512 256
run 2 1 0 512 512 256
run 1 0 0 8192 4096 1024 t
run 2 0 0 8192 4096 1024 t
run 2 0 1 8192 1024 4096 t
run 2 1 1 4096 1024 32768 t
run 2 0 0 32768 4096 1024 t
cat $L
scripts/gpu_suite.sh 900
