#!/bin/bash
# round-2 GPU call 13: multi-warp lanes (inner_loop_mw_kernel) — parity suite with LE_MW auto / 0, pop-16 latency A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=r2m
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${P}_pytest_auto.log
LE_MW=0 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${P}_pytest_mw0.log
for m in 0 auto; do
  if [ $m = auto ]; then unset LE_MW; else export LE_MW=$m; fi
  timeout 300 python bench.py --workload cartpole_se_pop16 --steps 5 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/${P}_bench_pop16_mw$m.log 2>&1
done
unset LE_MW
echo "== auto"; tail -30 gpurun_out/${P}_pytest_auto.log
echo "== LE_MW=0"; tail -5 gpurun_out/${P}_pytest_mw0.log
for f in gpurun_out/${P}_bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print(f, "%.3fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], "ms %.2f"%d["ms_per_step"], d.get("nes_generations_per_hour"))
except Exception as e:
    print(f, "FAILED", e, open(f).read()[-1500:])
PY
done
