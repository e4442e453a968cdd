#!/bin/bash
# round-2 GPU call 20: cost of the batched test rollouts (thin forward path) in the general kernel: test_episodes 10 vs 1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --extras none --workload acrobot_se_dueling"
timeout 600 python bench.py $ARGS > $O/bench_te10.log 2>&1
timeout 600 python bench.py $ARGS --lane-override test_episodes=1 > $O/bench_te1.log 2>&1
timeout 600 python bench.py $ARGS --lane-override use_test_env=0 > $O/bench_notest.log 2>&1
for f in $O/bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print(f, "%.3fM"%(d["value"]/1e6), "ms %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"])
except Exception as e:
    print(f, "FAILED", e, open(f).read()[-800:])
PY
done
