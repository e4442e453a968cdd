#!/bin/bash
# round-2 GPU call 6: parity suite (near-tie attribution, qgap trace, TC variants), dueling workloads FFMA vs tcgen05
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/r2f_pytest.log
for w in cartpole_se_dueling acrobot_se_dueling acrobot_se_dueling_tc; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --extras none > gpurun_out/r2f_bench_$w.log 2>&1
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2f_bench_cp.log 2>&1
tail -60 gpurun_out/r2f_pytest.log
for f in gpurun_out/r2f_bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print(f, "%.3fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], d["config"].get("resident_warp_slots"))
except Exception as e:
    print(f, "FAILED", e, open(f).read()[-600:])
PY
done
