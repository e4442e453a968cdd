#!/bin/bash
# round-2 GPU call 15: TD3_discrete_vary lanes — bench workload + ncu --set full of the persistent TD3 kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
timeout 900 python bench.py --workload td3_discrete --steps 2 --warmup 1 --no-cpu-baseline --extras none > $O/bench_td3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:td3 -c 1 -f -o $O/prof_td3 python bench.py --workload td3_discrete --members-per-gpu 296 --steps 1 --warmup 0 --no-cpu-baseline --extras none > $O/prof_td3_bench.log 2>&1
tail -3 $O/bench_td3.log | cut -c1-1200
tail -3 $O/prof_td3_bench.log | cut -c1-300
ls -la $O
