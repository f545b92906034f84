#!/bin/bash
# 1-GPU: whole -m gpu suite with the half-tile GEMM tail / fused dswiglu / loss changes, bench A/B of the tail split
mkdir -p gpurun_out
for f in tests/test_gpu_gemm.py tests/test_gpu_fullsize_parity.py tests/test_gpu_model.py tests/test_gpu_loss.py tests/test_gpu_attention.py tests/test_gpu_fullsize.py tests/test_gpu_elementwise.py tests/test_gpu_frontend.py tests/test_gpu_layout.py tests/test_gpu_optim.py tests/test_bestrq.py tests/test_gpu_peer_collective.py; do
  echo "=== $f"
  timeout 900 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider --tb=short -W ignore 2>&1 | tail -5
done
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench N=1 (split tail on)"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incumbent > gpurun_out/bench_r02_c6_split.log 2>&1; tail -1 gpurun_out/bench_r02_c6_split.log | cut -c1-700
echo "=== bench N=1 (split tail off)"
TN_GEMM_SPLIT_TAIL=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-incumbent > gpurun_out/bench_r02_c6_nosplit.log 2>&1; tail -1 gpurun_out/bench_r02_c6_nosplit.log | cut -c1-700
