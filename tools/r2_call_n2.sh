#!/bin/bash
# 2-GPU call: multi-GPU parity tests, FSDP2 peer-memory collectives parity + A/B, bench N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
echo "=== test_gpu_multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | tail -8
echo "=== check_fsdp NCCL"; timeout 300 $TR tools/check_fsdp.py 2>&1 | tail -3
echo "=== check_fsdp PEER"; TN_FSDP_PEER=1 timeout 300 $TR tools/check_fsdp.py 2>&1 | tail -6
echo "=== bench N=2 NCCL"; timeout 600 $TR bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/bench_n2_nccl.log 2>&1; tail -1 gpurun_out/bench_n2_nccl.log | cut -c1-400
echo "=== bench N=2 PEER"; TN_FSDP_PEER=1 timeout 600 $TR bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/bench_n2_peer.log 2>&1; tail -1 gpurun_out/bench_n2_peer.log | cut -c1-400
echo "=== bench N=2 PEER 16 CTAs"; TN_FSDP_PEER=1 TN_FSDP_PEER_CTAS=16 timeout 600 $TR bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/bench_n2_peer16.log 2>&1; tail -1 gpurun_out/bench_n2_peer16.log | cut -c1-400
