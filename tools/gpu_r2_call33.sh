#!/bin/bash
# round-2 GPU call 33: SE pack split into a state-net section and two head sections (no zero layer-2 products, 3 instead of 4 float4 per CartPole record)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2ah
O=gpurun_out/r2ah
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest.log; tail -6 $O/pytest.log
timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
for w in cartpole_se sweep_h1024 acrobot_se cartpole_se_pop16; do
  timeout 300 python bench.py --workload $w --steps 4 --warmup 2 --no-cpu-baseline --extras none > $O/bench_$w.log 2>&1
done
for f in $O/bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
for l in [x for x in open(f) if x.startswith("{")]:
    d=json.loads(l); print(f, "%.3fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], "ms %.2f"%d["ms_per_step"])
PY
done
