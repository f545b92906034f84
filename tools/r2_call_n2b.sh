#!/bin/bash
# 2-GPU call B: tensor / context parallel on hardware - NCCL forms and the peer-memory / halo forms (parity, then timing)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
echo "=== check_tp NCCL"; timeout 300 $TR tools/check_tp.py 2>&1 | tail -3
echo "=== check_tp PEER"; TN_TP_PEER=1 timeout 300 $TR tools/check_tp.py 2>&1 | tail -6
echo "=== check_cp allgather"; timeout 300 $TR tools/check_cp.py 2>&1 | tail -3
echo "=== check_cp HALO"; TN_CP_HALO=1 timeout 300 $TR tools/check_cp.py 2>&1 | tail -6
for v in "tp2_nccl:--tp 2:" "tp2_peer:--tp 2:TN_TP_PEER=1" "cp2_ag:--cp 2:" "cp2_halo:--cp 2:TN_CP_HALO=1"; do
  name=${v%%:*}; rest=${v#*:}; flag=${rest%%:*}; envs=${rest#*:}
  echo "=== bench N=2 $name (T=16384)"
  env $envs timeout 600 $TR bench.py --gpus 2 $flag --seq-len 16384 --steps 6 --warmup 3 --no-e2e > gpurun_out/bench_n2_$name.log 2>&1
  tail -1 gpurun_out/bench_n2_$name.log | cut -c1-330
done
