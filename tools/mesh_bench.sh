#!/bin/bash
# bench.py on alternative meshes (TP / CP), N GPUs.  usage: mesh_bench.sh N "tp cp seq_len" ...
N=${1:-2}; shift
mkdir -p gpurun_out
for v in "$@"; do
  set -- $v
  log=gpurun_out/mesh_n${N}_tp$1_cp$2_T$3.log
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --tp $1 --cp $2 --seq-len $3 --steps 4 --warmup 3 --no-e2e > $log 2>&1
  grep '^{' $log | tail -1 > gpurun_out/mesh_n${N}_tp$1_cp$2_T$3.json
  echo "tp=$1 cp=$2 T=$3: $(python -c 'import sys,json
try:
    d=json.loads(open(sys.argv[1]).read()); print(d["value"], d["ms_per_step"], d["extras"]["mem_gb"])
except Exception as e: print("FAILED")' gpurun_out/mesh_n${N}_tp$1_cp$2_T$3.json)"
  grep -v Warning $log | grep "Error" | tail -3
done
