#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/t35.log 2>&1
echo "gpu tests exit $?"; tail -n 3 gpurun_out/t35.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke35.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/smoke35.log
