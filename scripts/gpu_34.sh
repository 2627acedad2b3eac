#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/t34.log 2>&1
echo "gpu tests exit $?"; tail -n 6 gpurun_out/t34.log
timeout -k 10 600 python -m pytest tests -q -m "not gpu" -p no:cacheprovider --tb=short > gpurun_out/t34c.log 2>&1
echo "cpu tests on the box exit $?"; tail -n 4 gpurun_out/t34c.log
