#!/bin/bash
# Round-end evidence on one GPU: driver-style tests, smoke, full bench line, ncu launch list + GEMM DRAM traffic.
TAG=${1:-r01_final}
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
echo "=== pytest -m gpu (one process, as the driver runs it)"
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_${TAG}.log 2>&1; grep -E "passed|failed|error" gpurun_out/gpu_tests_${TAG}.log | tail -3
echo "=== bench"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.log 2>&1; grep '^{' gpurun_out/bench_${TAG}.log > gpurun_out/bench_${TAG}.json; cut -c1-400 gpurun_out/bench_${TAG}.json
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
echo "=== ncu launch list (bench.py --layers 2)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_launch_${TAG}.log | cut -c1-120
echo "=== ncu GEMM DRAM traffic"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"gemm" --csv --log-file gpurun_out/gemm_traffic_${TAG}.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_traffic_${TAG}.log 2>&1; python tools/gemm_traffic.py gpurun_out/gemm_traffic_${TAG}.csv gpurun_out/gemm_traffic_${TAG}.json | cut -c1-300
ls -la gpurun_out | head -30
