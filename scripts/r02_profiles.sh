#!/bin/bash
# Round-2 ncu evidence (run on the GPU box under gpurun; every command below has exited 0 without ncu before).
set -x
ARGS="--steps 5 --warmup 3 --no-cpu --no-scan --no-epoch --no-backbone --no-chain-variant"
python bench.py $ARGS > gpurun_out/r02_bench_cfg2_steps5.json 2> gpurun_out/r02_bench_cfg2_steps5.err || exit 1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_cfg2.csv python bench.py $ARGS > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:seg_reduce -c 4 -f -o gpurun_out/r02_pool_bwd python bench.py $ARGS > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:pool_fwd_kernel -s 12 -c 1 -f -o gpurun_out/r02_pool_fwd python bench.py $ARGS --no-config2 > /dev/null 2>&1
ls -la gpurun_out/
