#!/bin/bash
# round-2 GPU call 9: row-owner TD update (LE_ROWOWN=1) — parity suite, A/B against the unit-owner chunk loop (lible_old.so),
# shared-reciprocal tanh and one-record-per-iteration variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=r2i
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/${P}_pytest.log
for v in b200 old srcp rq1; do
  [ -f learning_environments_b200/csrc/lible_$v.so ] || continue
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/${P}_bench_cp_$v.log 2>&1
done
for v in b200 old; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --workload cartpole_rn --steps 3 --warmup 2 --no-cpu-baseline --extras none > gpurun_out/${P}_bench_rn_$v.log 2>&1
done
tail -15 gpurun_out/${P}_pytest.log
for f in gpurun_out/${P}_bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print(f, "%.3fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], "ms %.1f"%d["ms_per_step"])
except Exception as e:
    print(f, "FAILED", e, open(f).read()[-600:])
PY
done
