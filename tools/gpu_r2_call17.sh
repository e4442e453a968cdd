#!/bin/bash
# round-2 GPU call 17: racecheck + memcheck of the multi-warp lane kernel; full GPU suite; default bench.py run (wall time)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
{ echo '## mw (racecheck)'; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/race_loop.py mw 2>&1 | grep -E "lanes|RACECHECK SUMMARY|hazard|Error|error" | tail -8; echo "exit code: ${PIPESTATUS[0]}";
  echo '## mw (memcheck)'; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/race_loop.py mw 2>&1 | grep -E "lanes|ERROR SUMMARY|Invalid|Error|error" | tail -8; echo "exit code: ${PIPESTATUS[0]}"; } > $O/sanitizer_mw.txt 2>&1
cat $O/sanitizer_mw.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest.log
tail -4 $O/pytest.log
S=$(date +%s)
timeout 1500 python bench.py > $O/bench_default.log 2> $O/bench_default.err
echo "default bench wall seconds: $(( $(date +%s) - S ))" | tee $O/bench_default.time
tail -c 600 $O/bench_default.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2q/bench_default.log') if x.startswith('{')][-1]; d=json.loads(l)
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], 'e2e', d['e2e']['value'], d['cpu_baseline'].get('value'), d['cpu_baseline'].get('kind'))
for k,v in d['workloads'].items(): print(k, v.get('value'), v.get('frac'), v.get('ms_per_step'), v.get('error'))
print('strong', d['strong_scaling'].get('value'))
PY
