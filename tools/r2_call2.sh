#!/bin/bash
# call 2: attention tests with the reworked softmax paths, timings, GEMM L2-hint A/B, ncu of the attention kernels
mkdir -p gpurun_out
for f in tests/test_gpu_attention.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_parity.py tests/test_gpu_model.py; do
  echo "=== $f"
  timeout 900 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider --tb=short -W ignore 2>&1 | tail -6
done
echo "=== attn bench"
timeout 300 python tools/attn_bench.py --case cfg2,cfg3,cfg4,cfg5 2>&1 | tail -4
echo "=== bench N=1 (hints on)"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_c2_hints.log 2>&1; tail -1 gpurun_out/bench_r02_c2_hints.log | cut -c1-2600
echo "=== bench N=1 (hints off)"
TN_GEMM_L2_HINTS=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02_c2_nohints.log 2>&1; tail -1 gpurun_out/bench_r02_c2_nohints.log | cut -c1-700
echo "=== ncu attention (set full)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 4 -o gpurun_out/r02_attn_cfg2 python tools/attn_bench.py --case cfg2 --iters 1 > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
ls -la gpurun_out/*.ncu-rep
