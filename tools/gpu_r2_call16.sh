#!/bin/bash
# round-2 GPU call 16: ncu of the persistent TD3 kernel on a short shape (74 lanes, 2 train episodes), selected sections only
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
LE_TD3_TRAIN_EPISODES=2 timeout 420 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section WarpStateStats --section Occupancy --section MemoryWorkloadAnalysis --section LaunchStats \
  --clock-control none -k regex:td3 -c 1 -f -o $O/prof_td3 python bench.py --workload td3_discrete --members-per-gpu 74 --steps 1 --warmup 0 --no-cpu-baseline --extras none > $O/prof_td3_bench.log 2>&1
tail -3 $O/prof_td3_bench.log | cut -c1-300
ls -la $O
