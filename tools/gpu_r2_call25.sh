#!/bin/bash
# round-2 GPU call 25: state rows kept in registers for the backward pass (LE_KEEP_S=1) — A/B + parity of the variant
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
for v in b200 keeps b200 keeps; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras none >> $O/bench_cp_$v.log 2>&1
done
for v in b200 keeps; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --workload cartpole_rn --steps 3 --warmup 2 --no-cpu-baseline --extras none > $O/bench_rn_$v.log 2>&1
done
LE_LIB_NAME=lible_keeps.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -5 > $O/pytest_keeps.log; tail -3 $O/pytest_keeps.log
for f in $O/bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
for l in [x for x in open(f) if x.startswith("{")]:
    d=json.loads(l); print(f, "%.3fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], "ms %.1f"%d["ms_per_step"])
PY
done
