#!/bin/bash
# round-2 GPU call 8: row-split thin forward path (test rollouts) — parity + dueling / vary_hp workloads; shared-reciprocal tanh A/B;
# racecheck of the tcgen05 loop kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2h_pytest.log
for w in cartpole_se_dueling acrobot_se_dueling acrobot_se_dueling_tc vary_hp; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --extras none > gpurun_out/r2h_bench_$w.log 2>&1
done
for v in b200 srcp; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2h_bench_cp_$v.log 2>&1
done
LE_LIB_NAME=lible_srcp.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lockstep or td_update or many_lanes or edge" 2>&1 | tail -15 > gpurun_out/r2h_pytest_srcp.log
{ echo '## tc'; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/race_loop.py tc 2>&1 | grep -E "lanes|RACECHECK SUMMARY|hazard|Error|error" | tail -8; echo "exit code: $?"; } > gpurun_out/r2h_sanitizer_tc.txt 2>&1
tail -12 gpurun_out/r2h_pytest.log
tail -6 gpurun_out/r2h_pytest_srcp.log
cat gpurun_out/r2h_sanitizer_tc.txt
for f in gpurun_out/r2h_bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print(f, "%.3fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], d.get("seconds_per_evaluation"))
except Exception as e:
    print(f, "FAILED", e, open(f).read()[-600:])
PY
done
