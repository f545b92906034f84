#!/bin/bash
# profiles call: GEMM DRAM traffic of the current step (metrics only) + --set full of the HBM-bound kernels and the top GEMMs
mkdir -p gpurun_out
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm --csv --log-file gpurun_out/r02_gemm_traffic.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-incumbent > gpurun_out/ncu_traffic.log 2>&1; tail -1 gpurun_out/ncu_traffic.log | cut -c1-100
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"rmsnorm|swiglu_bwd|pack_ce|colsum|cast_f32|gemm_pair" -s 120 -c 40 -o gpurun_out/r02_step_kernels python bench.py --layers 1 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-incumbent > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-100
ls -la gpurun_out/r02_step_kernels.ncu-rep gpurun_out/r02_gemm_traffic.csv
