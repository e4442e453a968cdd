#!/bin/bash
# round-2 GPU call 34: ncu --set full of the cluster lane kernel at population 16
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2ai
O=gpurun_out/r2ai
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inner_loop_mwc_kernel -c 1 -f -o $O/prof_mwc python bench.py --workload cartpole_se_pop16 --steps 1 --warmup 1 --no-cpu-baseline --extras none > $O/prof_bench.log 2>&1
tail -2 $O/prof_bench.log | cut -c1-160
