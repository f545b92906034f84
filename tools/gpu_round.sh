#!/bin/bash
# tests + bench + ncu evidence in one lease.  usage: tools/gpu_round.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
echo "=== attention/model tests"
for f in tests/test_gpu_attention.py tests/test_gpu_model.py tests/test_gpu_gemm.py; do timeout 900 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider --tb=short -x 2>&1 | tail -4; done
echo "=== bench full"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.log 2>&1; tail -1 gpurun_out/bench_${TAG}.log | cut -c1-2500
echo "=== ncu launch list (L=2)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_launch_${TAG}.log | cut -c1-100
