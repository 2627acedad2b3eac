#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/ddp_check.py > gpurun_out/ddp_check.log 2>&1
echo "ddp_check exit $?"; tail -n 4 gpurun_out/ddp_check.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "bench n2 exit $?"; tail -n 1 gpurun_out/bench_n2.log | cut -c1-400
timeout -k 10 900 python bench.py --gpus 1 --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_n1.log 2>&1
echo "bench n1 exit $?"; tail -n 1 gpurun_out/bench_n1.log | cut -c1-300
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/t8.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/t8.log
