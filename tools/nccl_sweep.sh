#!/bin/bash
# N-GPU sweep: NCCL protocol / CTA budget x persistent-grid SM margin.  usage: nccl_sweep.sh N "proto ctas margin" ...
N=${1:-2}; shift
mkdir -p gpurun_out
for v in "$@"; do
  set -- $v
  log=gpurun_out/nccl_n${N}_$1_c$2_m$3.log
  if [ "$1" != "default" ]; then export NCCL_PROTO=$1; else unset NCCL_PROTO; fi
  if [ "$2" != "0" ]; then export NCCL_MAX_CTAS=$2; else unset NCCL_MAX_CTAS; fi
  TN_FSDP_RESHARD=0 TN_SM_MARGIN=$3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3 --no-e2e > $log 2>&1
  echo "proto=$1 max_ctas=$2 margin=$3: $(grep '^{' $log | tail -1 | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"])
except Exception as e: print("FAILED")')"
done
