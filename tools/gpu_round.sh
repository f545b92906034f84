#!/bin/bash
TAG=${1:-rXX}
mkdir -p gpurun_out
echo "=== tests"
for f in tests/test_gpu_attention.py tests/test_bestrq.py tests/test_gpu_model.py; do timeout 900 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -E "passed|failed|^E " | head -8; done
echo "=== ncu launch list (L=2)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_launch_${TAG}.log | cut -c1-100
echo "=== ncu full (bounded: 16 launches of the last step)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair_kernel|attn_fwd_kernel|attn_bwd_kernel" -s 75 -c 16 -o gpurun_out/prof_${TAG} python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_full_${TAG}.log | cut -c1-100
du -sh gpurun_out; ls -la gpurun_out | tail -6
