#!/bin/bash
# final 1-GPU call: the driver's own commands (one-process -m gpu suite, smoke, default bench), then the ncu launch list
mkdir -p gpurun_out
echo "=== pytest -m gpu (one process, as the driver runs it)"
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider -W ignore 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench default"
timeout 900 python bench.py > gpurun_out/bench_r02_final.log 2>&1; tail -1 gpurun_out/bench_r02_final.log | cut -c1-1200
echo "=== bench --with-optimizer"
timeout 600 python bench.py --with-optimizer --steps 4 --no-e2e --no-cpu-baseline --no-incumbent > gpurun_out/bench_r02_with_opt.log 2>&1; tail -1 gpurun_out/bench_r02_with_opt.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['extras']['with_optimizer'])" 2>&1 | cut -c1-600
echo "=== ncu launch list (layers 2)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_L2.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-incumbent > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-120; wc -l gpurun_out/r02_launches_L2.csv
