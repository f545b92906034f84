#!/bin/bash
# round-2 first contact: the whole -m gpu suite (incl. the new full-size parity + AC tests), the incumbent harness, bench N=1
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider --tb=short -x 2>&1 | tail -25
echo "=== incumbent"
timeout 900 python tools/incumbent.py --out gpurun_out/r02_incumbent_start.json > gpurun_out/incumbent_start.log 2>&1; tail -5 gpurun_out/incumbent_start.log
echo "=== bench N=1"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_start.log 2>&1; tail -1 gpurun_out/bench_r02_start.log | cut -c1-2500
