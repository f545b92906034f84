#!/bin/bash
# 2-GPU call D: direct-push reduce-scatter (parity, A/B vs staged push), loss parallel under TP
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
echo "=== check_fsdp PUSH direct"; TN_FSDP_PEER=push timeout 300 $TR tools/check_fsdp.py 2>&1 | tail -4
echo "=== check_tp (+ loss parallel)"; timeout 300 $TR tools/check_tp.py 2>&1 | grep -E "TP|Error|error" | tail -6
echo "=== test_gpu_multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider --tb=short -W ignore 2>&1 | tail -4
echo "=== bench N=2 push direct"; timeout 600 $TR bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2_push_direct.log 2>&1; tail -1 gpurun_out/bench_n2_push_direct.log | cut -c1-330
echo "=== bench N=2 push staged"; TN_FSDP_DIRECT=0 timeout 600 $TR bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/bench_n2_push_staged.log 2>&1; tail -1 gpurun_out/bench_n2_push_staged.log | cut -c1-330
echo "=== bench N=2 tp2 loss-parallel (T=16384)"; timeout 600 $TR bench.py --gpus 2 --tp 2 --seq-len 16384 --steps 6 --warmup 3 --no-e2e > gpurun_out/bench_n2_tp2_lp.log 2>&1; tail -1 gpurun_out/bench_n2_tp2_lp.log | cut -c1-330
grep -iE "Traceback|Error" gpurun_out/bench_n2_push_direct.log gpurun_out/bench_n2_tp2_lp.log | head -5
