#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_trainer_gpu.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/t33.log 2>&1
echo "trainer tests exit $?"; tail -n 25 gpurun_out/t33.log
