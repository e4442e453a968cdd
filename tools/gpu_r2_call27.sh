#!/bin/bash
# round-2 GPU call 27: ncu --set full of the U = 4 fused kernel (Acrobot SE + DDQN)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2ab
O=gpurun_out/r2ab
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inner_loop_kernel -c 1 -f -o $O/prof_inner python bench.py --workload acrobot_se --steps 1 --warmup 1 --no-cpu-baseline --extras none > $O/prof_bench.log 2>&1
tail -2 $O/prof_bench.log | cut -c1-200
