#!/bin/bash
# Runs the GPU test files one process each (a hung kernel cannot take the others down); logs land in gpurun_out/.
# usage: scripts/gpu_suite.sh [per-file timeout seconds] [files...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-600}; shift
FILES=${@:-"tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py"}
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc_all=0
for f in $FILES; do
  name=$(basename $f .py)
  echo "=== $f"
  timeout -k 10 $T python -m pytest $f -q -m gpu -p no:cacheprovider -x --tb=short -s > gpurun_out/$name.log 2>&1
  rc=$?
  echo "exit $rc"; tail -n 25 gpurun_out/$name.log
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
