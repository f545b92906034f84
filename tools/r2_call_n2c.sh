#!/bin/bash
# 2-GPU call C: FSDP2 collectives as copy-engine pushes (parity, A/B vs NCCL on the same box), then TP / CP on hardware
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
echo "=== dswiglu test"; timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q --no-header -p no:cacheprovider -W ignore -k dswiglu 2>&1 | tail -2
echo "=== check_fsdp PUSH"; TN_FSDP_PEER=push timeout 300 $TR tools/check_fsdp.py 2>&1 | tail -4
echo "=== bench N=2 NCCL"; timeout 600 $TR bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/bench_n2_nccl.log 2>&1; tail -1 gpurun_out/bench_n2_nccl.log | cut -c1-330
echo "=== bench N=2 PUSH"; TN_FSDP_PEER=push timeout 600 $TR bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/bench_n2_push.log 2>&1; tail -1 gpurun_out/bench_n2_push.log | cut -c1-330
bash tools/r2_call_n2b.sh
