#!/bin/bash
# `ncu --set full` of ONE in-step launch of the attention output projection (32768x1024x1024, residual-dropout
# epilogue): GEMM launch 9 of the first step = the attention projection of layer 2 (four forward GEMMs per layer).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r03b}
timeout -k 10 500 ncu --set full --clock-control none -k regex:gemm_bf16 -s 9 -c 1 -f \
  -o gpurun_out/${TAG}_gemm_cproj_instep python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline \
  --skip-masked-rows-head --skip-extras > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/${TAG}_ncu.log | cut -c1-200
ncu -i gpurun_out/${TAG}_gemm_cproj_instep.ncu-rep --page details > gpurun_out/${TAG}_gemm_cproj_instep.details.txt 2>&1
grep -E "gemm_bf16|Duration|DRAM Throughput|Memory Throughput|L2 Hit|Executed Ipc Active|Issue Slots Busy|SM Frequency" gpurun_out/${TAG}_gemm_cproj_instep.details.txt | head -12
