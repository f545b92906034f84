#!/bin/bash
# N-GPU sweep: NCCL CTA budget x persistent-grid SM margin (does keeping SMs free for NCCL beat sharing them?).
N=${1:-2}
mkdir -p gpurun_out
for v in "0 0" "8 0" "8 8" "4 4" "16 16" "2 2"; do
  set -- $v
  log=gpurun_out/nccl_n${N}_c$1_m$2.log
  if [ "$1" != "0" ]; then export NCCL_MAX_CTAS=$1; else unset NCCL_MAX_CTAS; fi
  TN_FSDP_RESHARD=0 TN_SM_MARGIN=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3 > $log 2>&1
  echo "max_ctas=$1 margin=$2: $(grep '^{' $log | tail -1 | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"])
except Exception as e: print("FAILED")')"
done
