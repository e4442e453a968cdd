#!/bin/bash
# round-2 GPU call 31: cluster lanes — full suite, racecheck / memcheck, pop-16 and pop-24 A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2ag
O=gpurun_out/r2ag
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest.log; tail -6 $O/pytest.log
{ echo '## cluster lanes (racecheck)'; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/race_loop.py mwc 2>&1 | grep -E "lanes|RACECHECK SUMMARY|hazard|Error|error" | tail -8; echo "exit code: ${PIPESTATUS[0]}";
  echo '## cluster lanes (memcheck)'; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/race_loop.py mwc 2>&1 | grep -E "lanes|ERROR SUMMARY|Invalid|Error|error" | tail -8; echo "exit code: ${PIPESTATUS[0]}"; } > $O/sanitizer_mwc.txt 2>&1
cat $O/sanitizer_mwc.txt
for mp in 16 24; do
 for m in 1 0; do
  LE_MWC=$m timeout 300 python bench.py --workload cartpole_se_pop16 --members-per-gpu $mp --steps 5 --warmup 3 --no-cpu-baseline --extras none > $O/bench_pop${mp}_mwc$m.log 2>&1
 done
done
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
