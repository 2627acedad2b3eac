#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 72 -c 3 -f -o gpurun_out/attn_v3 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?"; tail -3 gpurun_out/ncu_attn.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29513 scripts/ddp_check.py > gpurun_out/ddp_check1.log 2>&1; echo "ddp1 exit $?"; tail -2 gpurun_out/ddp_check1.log
