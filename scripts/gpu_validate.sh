#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/t_final.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/t_final.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/smoke_final.log
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; echo "ref exit $?"; tail -n 1 gpurun_out/bench_ref_final.log | cut -c1-200
timeout -k 10 1200 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_final.log | cut -c1-300
timeout -k 10 1200 python bench.py --dropout 0.0 --skip-cpu-baseline --skip-masked-rows-head > gpurun_out/bench_final_nodrop.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_final_nodrop.log | cut -c1-200
timeout -k 10 600 python scripts/bench_encode.py > gpurun_out/encode_final.jsonl 2>gpurun_out/encode_final.err; echo "encode exit $?"
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -f -o gpurun_out/gemm_cg2_v4 python scripts/gemm_probe.py 2 0 0 32768 4096 1024 t > gpurun_out/ncu_c.log 2>&1
echo "ncu gemm exit $?"
timeout -k 10 600 ncu --set full --clock-control none -k regex:attn_tc_ -s 68 -c 5 -f -o gpurun_out/attn_final python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline --skip-masked-rows-head > gpurun_out/ncu_a.log 2>&1
echo "ncu attn exit $?"
