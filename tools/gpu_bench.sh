#!/bin/bash
# smoke + bench + ncu evidence in one box lease
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== model test (bf16 master)"; timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --no-header -p no:cacheprovider -k "llama_forward" --tb=short 2>&1 | tail -5
echo "=== bench L=4"; timeout 600 python bench.py --layers 4 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
echo "=== bench full"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1; tail -2 gpurun_out/bench_full.log
echo "=== ncu launch list (L=2)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-200
echo "=== ncu full: gemm / attn"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attn_fwd_kernel|attn_bwd_kernel" -s 60 -c 12 -o gpurun_out/prof_r01 python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/
