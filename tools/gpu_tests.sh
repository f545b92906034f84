#!/bin/bash
# Run every -m gpu test file in its own process (a device-side trap poisons the CUDA context of that process only).
mkdir -p gpurun_out
for f in tests/test_gpu_elementwise.py tests/test_gpu_gemm.py tests/test_gpu_frontend.py tests/test_gpu_attention.py tests/test_gpu_model.py; do
  echo "=== $f"
  timeout 900 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | tail -${TAIL:-60}
done
