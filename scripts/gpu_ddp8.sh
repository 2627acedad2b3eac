#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench_n8.log 2>&1
echo "bench n8 exit $?"; tail -n 1 gpurun_out/bench_n8.log | cut -c1-300
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 4 --warmup 3 --skip-masked-rows-head > gpurun_out/bench_n4.log 2>&1
echo "bench n4 exit $?"; tail -n 1 gpurun_out/bench_n4.log | cut -c1-300
