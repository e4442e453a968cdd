#!/bin/bash
# ncu capture of the general (CTA-per-lane) kernel on a shortened dueling workload. Output -> gpurun_out/.
set -u
mkdir -p gpurun_out
W=${WORKLOAD:-acrobot_se_dueling}
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight \
    --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis \
    --clock-control none --import-source on -k regex:general_loop_kernel -c 1 -f -o gpurun_out/prof_general \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --workload $W --members-per-gpu ${MEMBERS:-98} \
    --lane-override max_steps=${MAX_STEPS:-120} --lane-override test_episodes=2 > gpurun_out/prof_general.log 2>&1
tail -1 gpurun_out/prof_general.log | cut -c1-300
