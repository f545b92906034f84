#!/bin/bash
# 8-GPU call B: default mesh with the direct-push reduce-scatter, cfg 4 / cfg 5 with the final defaults
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) "$@"; }
echo "=== bench N=8 default (FSDP2 direct push)"; timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 5 --warmup 3" > gpurun_out/bench_n8_fsdp_direct.log 2>&1; tail -1 gpurun_out/bench_n8_fsdp_direct.log | cut -c1-330
echo "=== bench N=8 cfg5: TP=2 x FSDP2=4, T=16384 (loss parallel)"; timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --tp 2 --seq-len 16384 --steps 4 --warmup 3 --no-e2e" > gpurun_out/bench_n8_cfg5_tp2_lp.log 2>&1; tail -1 gpurun_out/bench_n8_cfg5_tp2_lp.log | cut -c1-330
echo "=== bench N=8 cfg4: CP=4 x FSDP2=2, T=32768"; timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --cp 4 --seq-len 32768 --steps 4 --warmup 3 --no-e2e" > gpurun_out/bench_n8_cfg4_cp4_direct.log 2>&1; tail -1 gpurun_out/bench_n8_cfg4_cp4_direct.log | cut -c1-330
echo "=== bench N=4 default"; timeout 600 bash -c "$(declare -f run); run 4 bench.py --gpus 4 --steps 4 --warmup 3 --no-e2e" > gpurun_out/bench_n4_fsdp_direct.log 2>&1; tail -1 gpurun_out/bench_n4_fsdp_direct.log | cut -c1-330
for f in gpurun_out/bench_n8_fsdp_direct.log gpurun_out/bench_n8_cfg5_tp2_lp.log gpurun_out/bench_n8_cfg4_cp4_direct.log gpurun_out/bench_n4_fsdp_direct.log; do grep -v '^{' $f | grep -iE "error|Traceback|unavailable" | head -3; done
