#!/bin/bash
# round-2 GPU call 11: software-pipelined row-owner loop; reciprocal sharing / fold variants: parity per variant + bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=r2k
for v in b200 v0 v1 v2; do
  [ -f learning_environments_b200/csrc/lible_$v.so ] || continue
  LE_LIB_NAME=lible_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -25 > gpurun_out/${P}_pytest_$v.log
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/${P}_bench_cp_$v.log 2>&1
done
LE_LIB_NAME=lible_b200.so timeout 300 python bench.py --workload cartpole_rn --steps 3 --warmup 2 --no-cpu-baseline --extras none > gpurun_out/${P}_bench_rn_b200.log 2>&1
for v in b200 v0 v1 v2; do echo "== $v"; tail -8 gpurun_out/${P}_pytest_$v.log; done
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
