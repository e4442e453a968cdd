#!/bin/bash
# round-2 GPU call 12: knock-out timing variants of the row-owner kernel (results are wrong by construction; timing only)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=r2l
for v in b200 ko1 ko2 ko3 ko4; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --extras none > gpurun_out/${P}_bench_cp_$v.log 2>&1
done
for f in gpurun_out/${P}_bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print(f, "%.3fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], "ms %.1f"%d["ms_per_step"], {k:v for k,v in d.items() if 'step' in k or 'learn' in k})
except Exception as e:
    print(f, "FAILED", e, open(f).read()[-600:])
PY
done
