#!/bin/bash
TAG=${1:-rXX}
mkdir -p gpurun_out
echo "=== driver-style: pytest -m gpu (one process)"
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
echo "=== frontend timing"; python tools/time_frontend.py 2>&1 | tail -1; TN_FBANK_DFT=1 python tools/time_frontend.py 2>&1 | tail -1
echo "=== bench full"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.log 2>&1; tail -1 gpurun_out/bench_${TAG}.log | cut -c1-3300
echo "=== ncu gemm traffic (metrics only, L=2)"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"gemm" --csv --log-file gpurun_out/gemm_traffic_${TAG}.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_traffic_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_traffic_${TAG}.log | cut -c1-80
