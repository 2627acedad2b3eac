#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -f -o gpurun_out/gemm_cg2_v3 python scripts/gemm_probe.py 2 0 0 32768 4096 1024 t > gpurun_out/ncu_c.log 2>&1
echo "ncu cg2 exit $?"
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -f -o gpurun_out/gemm_cg1_v4 python scripts/gemm_probe.py 1 0 0 32768 4096 1024 t > gpurun_out/ncu_d.log 2>&1
echo "ncu cg1 exit $?"
