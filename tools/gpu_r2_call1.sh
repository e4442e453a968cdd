#!/bin/bash
# round-2 GPU call 1: parity tests on the default build + A/B of the fused-kernel variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpu.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2_pytest1.log
for v in v0 v1 v2 v3; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$v.log 2>&1
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --workload acrobot_se --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ac_$v.log 2>&1
done
tail -3 gpurun_out/r2_pytest1.log
for v in v0 v1 v2 v3; do python - <<PY
import json
for f in ("gpurun_out/r2_bench_$v.log","gpurun_out/r2_bench_ac_$v.log"):
    try:
        l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
        print("$v", f.split("_")[-2] if "ac" in f else "cp", "%.2fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], d["config"].get("resident_warp_slots"))
    except Exception as e:
        print("$v", f, "FAILED", e, open(f).read()[-500:])
PY
done
