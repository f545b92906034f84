#!/bin/bash
# A/B of the GEMM raster autotuner on the full bench workload (same box, back to back, twice).
mkdir -p gpurun_out
for rep in 1 2; do
for v in 0 1; do
  TN_GEMM_AUTOTUNE=$v TN_GEMM_AUTOTUNE_LOG=1 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_tune$v.log 2> gpurun_out/ab_tune$v.err
  echo "autotune=$v: $(grep '^{' gpurun_out/ab_tune$v.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])')"
done; done
grep "raster group" gpurun_out/ab_tune1.err | cut -c1-150
