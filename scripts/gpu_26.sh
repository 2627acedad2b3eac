#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/t26.log 2>&1
echo "tests exit $?"; tail -n 5 gpurun_out/t26.log
timeout -k 10 1200 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench26.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench26.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['masked_rows_head'], d['roofline']['achieved'], d['clocks'])"
