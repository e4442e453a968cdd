#!/bin/bash
# round-2 GPU call 28: multi-warp lanes with worker-side prefetch of the next minibatch pass (leader no longer takes passes when W > 4)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2ad
O=gpurun_out/r2ad
timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest.log; tail -6 $O/pytest.log
for mp in 16 64; do
 for v in b200 mwold; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --workload cartpole_se_pop16 --members-per-gpu $mp --steps 5 --warmup 3 --no-cpu-baseline --extras none > $O/bench_pop${mp}_$v.log 2>&1
 done
done
{ echo '## mw + prefetch (racecheck)'; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/race_loop.py mw 2>&1 | grep -E "lanes|RACECHECK SUMMARY|hazard|Error|error" | tail -8; echo "exit code: ${PIPESTATUS[0]}";
  echo '## mw + prefetch (memcheck)'; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/race_loop.py mw 2>&1 | grep -E "lanes|ERROR SUMMARY|Invalid|Error|error" | tail -8; echo "exit code: ${PIPESTATUS[0]}"; } > $O/sanitizer_mw.txt 2>&1
cat $O/sanitizer_mw.txt
for f in $O/bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print(f, "%.3fM"%(d["value"]/1e6), "ms %.2f"%d["ms_per_step"], d.get("nes_generations_per_hour"))
except Exception as e:
    print(f, "FAILED", e, open(f).read()[-1500:])
PY
done
