#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_attn.py 2>&1 | tail -12
for f in tests/test_gpu_attention.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_parity.py tests/test_gpu_model.py; do
  echo "=== $f"
  timeout 900 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider --tb=line -W ignore 2>&1 | tail -4
done
echo "=== attn bench"
timeout 300 python tools/attn_bench.py --case cfg2,cfg3,cfg4,cfg5 2>&1 | tail -4
echo "=== ncu attention (set full)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 5 -o gpurun_out/r02_attn_cfg2_c python tools/attn_bench.py --case cfg2 --iters 1 > gpurun_out/ncu_attn.log 2>&1; tail -1 gpurun_out/ncu_attn.log
