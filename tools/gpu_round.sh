#!/bin/bash
TAG=${1:-rXX}
mkdir -p gpurun_out
echo "=== tests"
timeout 600 python -m pytest tests/test_gpu_loss.py -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -E "passed|failed|Error|assert|^E |^tests" | head -20
for f in tests/test_gpu_attention.py tests/test_gpu_model.py; do timeout 900 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -E "passed|failed|Error|assert|^E " | head -12; done
echo "=== diag hf"; timeout 600 python tools/diag_hf.py 2>&1 | grep -v Warning | tail -25
echo "=== bench full"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.log 2>&1; tail -1 gpurun_out/bench_${TAG}.log | cut -c1-3500
