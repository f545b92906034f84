#!/bin/bash
# round-2 first contact: every -m gpu test file in its own process (a device-side trap poisons only that process),
# the new persistent attention forward vs the round-1 kernel, the incumbent harness, bench N=1
mkdir -p gpurun_out
for f in tests/test_gpu_attention.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_parity.py tests/test_gpu_loss.py tests/test_gpu_model.py tests/test_gpu_elementwise.py tests/test_gpu_gemm.py tests/test_gpu_frontend.py tests/test_gpu_layout.py tests/test_gpu_optim.py tests/test_bestrq.py; do
  echo "=== $f"
  timeout 900 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | tail -${TAIL:-25}
done
echo "=== attention v1 (TN_ATTN_FWD_V1=1) parity"
TN_ATTN_FWD_V1=1 timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q --no-header -p no:cacheprovider --tb=line 2>&1 | tail -3
echo "=== attn bench v2 / v1"
timeout 300 python tools/attn_bench.py --case cfg2,cfg3,cfg4 2>&1 | tail -4
TN_ATTN_FWD_V1=1 timeout 300 python tools/attn_bench.py --case cfg2,cfg4 2>&1 | tail -3
echo "=== incumbent"
timeout 900 python tools/incumbent.py --out gpurun_out/r02_incumbent_start.json > gpurun_out/incumbent_start.log 2>&1; tail -5 gpurun_out/incumbent_start.log
echo "=== bench N=1"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_start.log 2>&1; tail -1 gpurun_out/bench_r02_start.log | cut -c1-3000
