#!/bin/bash
# round-2 GPU call 22: compact reduction for all layouts (AD = 2: one float4, AD = 3: two) + paired backward seeds — parity + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest.log; tail -5 $O/pytest.log
for w in cartpole_se acrobot_se cartpole_rn; do
 for v in b200 old; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline --extras none > $O/bench_${w}_$v.log 2>&1
 done
done
for f in $O/bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
for l in [x for x in open(f) if x.startswith("{")]:
    d=json.loads(l); print(f, "%.3fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], "ms %.1f"%d["ms_per_step"])
PY
done
