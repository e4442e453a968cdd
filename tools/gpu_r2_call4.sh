#!/bin/bash
# round-2 GPU call 4: tcgen05 GEMM unit test + A/B of the non-pipelined fused-kernel variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -30 > gpurun_out/r2d_pytest_tc.log
for v in old b200 q0 q1 q2 q3; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2d_bench_$v.log 2>&1
done
for v in old b200; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --workload acrobot_se --steps 3 --warmup 2 --no-cpu-baseline --extras none > gpurun_out/r2d_bench_ac_$v.log 2>&1
done
tail -30 gpurun_out/r2d_pytest_tc.log
for f in gpurun_out/r2d_bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print(f, "%.2fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], d["config"].get("resident_warp_slots"))
except Exception as e:
    print(f, "FAILED", e, open(f).read()[-600:])
PY
done
