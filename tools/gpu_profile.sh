#!/bin/bash
# ncu evidence for the fused kernel: launch list of one bench run + one --set full capture. Output -> gpurun_out/.
set -u
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --no-cpu-baseline ${BENCH_ARGS:-}"
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-300
echo "== full capture"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:inner_loop_kernel -c 1 -f -o gpurun_out/prof_inner \
    python bench.py $ARGS > gpurun_out/prof_bench.log 2>&1
tail -2 gpurun_out/prof_bench.log | cut -c1-300
ls -la gpurun_out
