#!/bin/bash
# Round-2 validation pass on one B200: whole GPU test suite, smoke, both bench arms, ncu launch list of one step, and
# `ncu --set full` captures of the dominant GEMM (DRAM bytes for roofline.traffic), ln_bwd and adamw.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02e}
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -n 6 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout -k 10 300 python -m pytest tests/test_model_gpu.py -q -s -k "clip_and_muadamw or mlm_loss_and_gradients" -p no:cacheprovider 2>&1 | grep -E "update rel err|bf16_h" | cut -c1-1500 > gpurun_out/${TAG}_tolerances.log
cat gpurun_out/${TAG}_tolerances.log | cut -c1-600
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?"; tail -n 1 gpurun_out/${TAG}_smoke.log
timeout -k 10 300 python scripts/attn_probe.py 2>/dev/null | tail -n 1 > gpurun_out/${TAG}_attn_probe.log; cut -c1-400 gpurun_out/${TAG}_attn_probe.log
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -n 1 gpurun_out/${TAG}_bench.log | cut -c1-900; tail -n 3 gpurun_out/${TAG}_bench.err
timeout -k 10 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_ref.log 2>&1
echo "bench ref exit $?"; tail -n 1 gpurun_out/${TAG}_bench_ref.log | cut -c1-400
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --global-batch 32 \
  --skip-cpu-baseline --skip-masked-rows-head --skip-extras > gpurun_out/${TAG}_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt 2>&1; head -n 26 gpurun_out/${TAG}_launches.txt
timeout -k 10 300 ncu --set full --clock-control none -k regex:gemm_bf16 -s 3 -c 1 -f -o gpurun_out/${TAG}_gemm_cfc \
  python scripts/gemm_probe.py 2 0 0 32768 4096 1024 t > gpurun_out/${TAG}_ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
ncu -i gpurun_out/${TAG}_gemm_cfc.ncu-rep --page details > gpurun_out/${TAG}_gemm_cfc.details.txt 2>&1
ncu -i gpurun_out/${TAG}_gemm_cfc.ncu-rep --page raw --csv > gpurun_out/${TAG}_gemm_cfc.raw.csv 2>&1
timeout -k 10 600 ncu --set full --clock-control none -k regex:'ln_bwd_kernel|adamw_kernel' -s 34 -c 3 -f \
  -o gpurun_out/${TAG}_rowwise python bench.py --steps 1 --warmup 3 --global-batch 32 --skip-cpu-baseline \
  --skip-masked-rows-head --skip-extras > gpurun_out/${TAG}_ncu_rowwise.log 2>&1
echo "ncu rowwise exit $?"
ncu -i gpurun_out/${TAG}_rowwise.ncu-rep --page details > gpurun_out/${TAG}_rowwise.details.txt 2>&1
PROBE_REPS=1 timeout -k 10 600 ncu --set full --clock-control none -k regex:'attn_tc|attn_keep_mask' -s 4 -c 6 -f \
  -o gpurun_out/${TAG}_attn python scripts/attn_probe.py > gpurun_out/${TAG}_ncu_attn.log 2>&1
echo "ncu attn exit $?"
ncu -i gpurun_out/${TAG}_attn.ncu-rep --page details > gpurun_out/${TAG}_attn.details.txt 2>&1
PROBE_REPS=1 timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:attn_keep_mask -s 1 -c 1 -f \
  -o gpurun_out/${TAG}_keepmask python scripts/attn_probe.py > gpurun_out/${TAG}_ncu_keepmask.log 2>&1
ncu -i gpurun_out/${TAG}_keepmask.ncu-rep --page details > gpurun_out/${TAG}_keepmask.details.txt 2>&1
du -sh gpurun_out
