#!/bin/bash
# N-GPU sweep of the FSDP2 knobs bench.py exposes (TN_FSDP_RESHARD, TN_FSDP_PREFETCH); logs under gpurun_out/.
N=${1:-2}
mkdir -p gpurun_out
for v in "1 0" "0 0" "0 2" "1 2"; do
  set -- $v
  log=gpurun_out/fsdp_n${N}_r$1_p$2.log
  TN_FSDP_RESHARD=$1 TN_FSDP_PREFETCH=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3 > $log 2>&1
  echo "reshard=$1 prefetch=$2: $(grep '^{' $log | tail -1 | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"])
except Exception as e: print("FAILED")')"
  grep -v Warning $log | grep -i "error" | tail -3
done
