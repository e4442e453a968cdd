#!/bin/bash
# round-2 GPU call 5: tcgen05 GEMM unit tests (all operand forms), full parity suite with the tensor-core general kernel,
# dueling workloads with / without tensor cores, two more fused-kernel variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q 2>&1 | tail -30 > gpurun_out/r2e_pytest_tc.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2e_pytest.log
for v in b200 notc; do
  for w in cartpole_se_dueling acrobot_se_dueling; do
    LE_LIB_NAME=lible_$v.so timeout 600 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --extras none > gpurun_out/r2e_bench_${w}_$v.log 2>&1
  done
done
for v in b200 r0 r2 old; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2e_bench_cp_$v.log 2>&1
done
tail -15 gpurun_out/r2e_pytest_tc.log
tail -25 gpurun_out/r2e_pytest.log
for f in gpurun_out/r2e_bench_*.log; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print(f, "%.3fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], d["config"].get("resident_warp_slots"))
except Exception as e:
    print(f, "FAILED", e, open(f).read()[-600:])
PY
done
