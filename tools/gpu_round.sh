#!/bin/bash
TAG=${1:-rXX}
mkdir -p gpurun_out
for v in 1 0 1 0; do echo "=== bench fused_qkv=$v"; TN_FUSED_QKV=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks'], {k:v for k,v in d['extras']['ms_by_entry_point_timed_region'].items() if 'gemm' in k})"; done
echo "=== ncu launch list (L=2)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_launch_${TAG}.log | cut -c1-100
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair_kernel|attn_fwd_kernel|attn_bwd_kernel|pack_ce|rmsnorm|frontend_kernel" -s 80 -c 40 -o gpurun_out/prof_${TAG} python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_full_${TAG}.log | cut -c1-100
ls -la gpurun_out | tail -5
