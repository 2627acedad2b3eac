#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
for n in 8 4 2; do
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 4 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
echo "bench n$n exit $?"; tail -n 1 gpurun_out/bench_n$n.log | cut -c1-200
done
