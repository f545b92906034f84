#!/bin/bash
# 8-GPU call: default FSDP2 mesh (push collectives) at N=8 and N=4, BASELINE cfg 4 (CP=4 x FSDP2=2, T=32768) and cfg 5
# (TP=2 x FSDP2=4, T=16384) at full size, CP=4 parity check.
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) "$@"; }
echo "=== check_cp world=4 (halo)"; timeout 300 bash -c "$(declare -f run); run 4 tools/check_cp.py" 2>&1 | tail -2
echo "=== bench N=8 default (FSDP2 push)"; timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e" > gpurun_out/bench_n8_fsdp_push.log 2>&1; tail -1 gpurun_out/bench_n8_fsdp_push.log | cut -c1-330
echo "=== bench N=8 cfg4: CP=4 x FSDP2=2, T=32768"; timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --cp 4 --seq-len 32768 --steps 4 --warmup 3 --no-e2e" > gpurun_out/bench_n8_cfg4_cp4.log 2>&1; tail -1 gpurun_out/bench_n8_cfg4_cp4.log | cut -c1-330
echo "=== bench N=8 cfg5: TP=2 x FSDP2=4, T=16384"; timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --tp 2 --seq-len 16384 --steps 4 --warmup 3 --no-e2e" > gpurun_out/bench_n8_cfg5_tp2.log 2>&1; tail -1 gpurun_out/bench_n8_cfg5_tp2.log | cut -c1-330
echo "=== bench N=4 default (FSDP2 push)"; timeout 600 bash -c "$(declare -f run); run 4 bench.py --gpus 4 --steps 4 --warmup 3 --no-e2e" > gpurun_out/bench_n4_fsdp_push.log 2>&1; tail -1 gpurun_out/bench_n4_fsdp_push.log | cut -c1-330
echo "=== bench N=8 NCCL (A/B)"; TN_FSDP_PEER=0 timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e" > gpurun_out/bench_n8_fsdp_nccl.log 2>&1; tail -1 gpurun_out/bench_n8_fsdp_nccl.log | cut -c1-330
for f in gpurun_out/bench_n8_*.log gpurun_out/bench_n4_fsdp_push.log; do echo "--- $f"; grep -v '^{' $f | grep -iE "error|Traceback|unavailable" | head -5; done
