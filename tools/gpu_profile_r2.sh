#!/bin/bash
# round-2 profile call: launch list + ncu --set full of the fused kernel (headline workload), tensor-pipe evidence of the
# tcgen05 general kernel, DRAM bytes of the full-size-ring regime, racecheck of the loop kernels at a 2-episode shape.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
ARGS="--steps 1 --warmup 1 --no-cpu-baseline --extras none"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py $ARGS > $O/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inner_loop_kernel -c 1 -f -o $O/prof_inner python bench.py $ARGS > $O/prof_bench.log 2>&1
# tcgen05 instantiation of the general kernel: tensor pipe + the usual pipes (short: 60 lanes, 1 training episode after the init one)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:general_loop_kernel -c 1 -f -o $O/prof_general_tc \
    python bench.py --workload acrobot_se_dueling_tc --members-per-gpu 50 $ARGS > $O/prof_general_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:general_loop_kernel -c 1 -f -o $O/prof_general_ffma \
    python bench.py --workload acrobot_se_dueling --members-per-gpu 99 $ARGS > $O/prof_general_ffma.log 2>&1
# full-size ring: DRAM bytes of the launch (two metrics: one or two replays of a 12 s kernel)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:inner_loop_kernel -c 1 \
    --csv --log-file $O/fullring_dram.csv python bench.py --workload cartpole_se_fullring $ARGS > $O/fullring_bench.log 2>&1
{
echo '# compute-sanitizer --tool racecheck python tools/race_loop.py <fused|general|tc>   (2-episode shapes of the persistent loop kernels)'
for w in fused general tc; do
  echo "## $w"
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/race_loop.py $w 2>&1 | grep -E "lanes|RACECHECK SUMMARY|hazard|Error|error" | tail -8
  echo "exit code: $?"
done
} > $O/sanitizer.txt 2>&1
cat $O/sanitizer.txt
tail -3 $O/prof_bench.log | cut -c1-300
tail -2 $O/prof_general_tc.log | cut -c1-300
tail -2 $O/fullring_bench.log | cut -c1-300
ls -la $O
