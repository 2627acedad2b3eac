#!/bin/bash
# Source-level captures of the GELU (+derivative) and rotary epilogues on their block shapes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02w}
timeout -k 10 120 python scripts/epi_probe.py 32768 3072 1024 > gpurun_out/${TAG}_epi_probe.txt 2>&1; cat gpurun_out/${TAG}_epi_probe.txt
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -f \
  -o gpurun_out/${TAG}_gemm_epi9 python scripts/epi_probe.py 32768 4096 1024 gelu_dg > gpurun_out/${TAG}_ncu_epi9.log 2>&1
echo "ncu epi9 exit $?"
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -f \
  -o gpurun_out/${TAG}_gemm_epi7 python scripts/epi_probe.py 32768 3072 1024 rope > gpurun_out/${TAG}_ncu_epi7.log 2>&1
echo "ncu epi7 exit $?"
