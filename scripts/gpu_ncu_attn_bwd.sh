#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_d -s 48 -c 2 -f -o gpurun_out/attn_bwd_v7 python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline > gpurun_out/ncu_b.log 2>&1
echo "ncu bwd exit $?"
