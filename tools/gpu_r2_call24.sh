#!/bin/bash
# round-2 GPU call 24: launch list + ncu --set full of the shipped fused kernel (v7, compact reduction) on the headline workload
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
ARGS="--steps 1 --warmup 1 --no-cpu-baseline --extras none"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py $ARGS > $O/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inner_loop_kernel -c 1 -f -o $O/prof_inner python bench.py $ARGS > $O/prof_bench.log 2>&1
tail -2 $O/prof_bench.log | cut -c1-200
ls -la $O
