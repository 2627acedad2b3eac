#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_attention_tc_gpu.py -q -m gpu -p no:cacheprovider --tb=short -s > gpurun_out/test_attention_tc_gpu.log 2>&1
echo "attn_tc exit $?"; tail -n 30 gpurun_out/test_attention_tc_gpu.log
scripts/gpu_suite.sh 900
timeout -k 10 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log
timeout -k 10 900 python bench.py --steps 2 --warmup 3 --global-batch 64 --skip-cpu-baseline > gpurun_out/bench_small.log 2>&1; echo "bench exit $?"; tail -n 5 gpurun_out/bench_small.log
timeout -k 10 900 python bench.py --steps 2 --warmup 3 --global-batch 64 --dropout 0.0 --skip-cpu-baseline > gpurun_out/bench_small_nodrop.log 2>&1; echo "bench exit $?"; tail -n 5 gpurun_out/bench_small_nodrop.log
for cg in 1 2; do
  python scripts/gemm_probe.py $cg 0 0 8192 4096 1024 t > gpurun_out/ncu_plain_$cg.log 2>&1 &&
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 1 -f -o gpurun_out/gemm_cg$cg python scripts/gemm_probe.py $cg 0 0 8192 4096 1024 > gpurun_out/ncu_cg$cg.log 2>&1
  echo "ncu cg$cg exit $?"
done
