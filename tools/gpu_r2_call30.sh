#!/bin/bash
# round-2 GPU call 30: cluster lanes (one lane per thread-block cluster of two CTAs, DSMEM) — guarded first run, parity, pop-16 A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2af
O=gpurun_out/r2af
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lockstep" 2>&1 | tail -15 > $O/first.log; echo "first rc=$?"; tail -6 $O/first.log
if grep -q "passed" $O/first.log && ! grep -q "failed\|Error" $O/first.log; then
  timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
  timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest.log; tail -6 $O/pytest.log
  for m in 1 0; do
    LE_MWC=$m timeout 300 python bench.py --workload cartpole_se_pop16 --steps 5 --warmup 3 --no-cpu-baseline --extras none > $O/bench_pop16_mwc$m.log 2>&1
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
fi
