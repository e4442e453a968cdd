#!/bin/bash
# round-2 GPU call 23: prescaled stage (2 log2(e) applied once to the staged states / biases) — parity + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest.log; tail -8 $O/pytest.log
for v in b200 old b200 old; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras none >> $O/bench_cp_$v.log 2>&1
done
for f in $O/bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
for l in [x for x in open(f) if x.startswith("{")]:
    d=json.loads(l); print(f, "%.3fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], "ms %.1f"%d["ms_per_step"])
PY
done
