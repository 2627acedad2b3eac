#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/epi_probe.py 32768 4096 1024 > gpurun_out/epi.log 2>&1; cat gpurun_out/epi.log
python scripts/epi_probe.py 32768 1024 1024 >> gpurun_out/epi.log 2>&1; tail -5 gpurun_out/epi.log
timeout -k 10 900 python -m pytest tests/test_shapes_gpu.py -q -m gpu -p no:cacheprovider --tb=short -s > gpurun_out/t6.log 2>&1
echo "shape tests exit $?"; tail -n 12 gpurun_out/t6.log
timeout -k 10 600 python scripts/bench_encode.py > gpurun_out/encode.log 2>&1; echo "encode exit $?"; tail -n 6 gpurun_out/encode.log
python scripts/epi_probe.py 32768 4096 1024 gelu > gpurun_out/ncu_plain_gelu.log 2>&1 &&
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -f -o gpurun_out/gemm_gelu python scripts/epi_probe.py 32768 4096 1024 gelu > gpurun_out/ncu_gelu.log 2>&1
echo "ncu gelu exit $?"
