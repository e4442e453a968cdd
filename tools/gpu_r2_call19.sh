#!/bin/bash
# round-2 GPU call 19: where the general (CTA-per-lane) kernel spends its time now (FFMA instantiation, Acrobot DuelingDDQN and CartPole DuelingDDQN)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
ARGS="--steps 1 --warmup 1 --no-cpu-baseline --extras none"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:general_loop_kernel -c 1 -f -o $O/prof_general_ac \
    python bench.py --workload acrobot_se_dueling --members-per-gpu 99 $ARGS > $O/prof_general_ac.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:general_loop_kernel -c 1 -f -o $O/prof_general_cp \
    python bench.py --workload cartpole_se_dueling --members-per-gpu 99 $ARGS > $O/prof_general_cp.log 2>&1
ls -la $O
