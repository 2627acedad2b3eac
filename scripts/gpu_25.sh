#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full25.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_full25.log | cut -c1-400
timeout -k 10 1200 python bench.py --steps 3 --warmup 3 --dropout 0.0 --skip-cpu-baseline > gpurun_out/bench_full25_nodrop.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_full25_nodrop.log | cut -c1-200
timeout -k 10 600 python scripts/bench_encode.py > gpurun_out/encode25.jsonl 2>gpurun_out/encode25.err; echo "encode exit $?"; cat gpurun_out/encode25.jsonl | cut -c1-200
