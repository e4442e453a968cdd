#!/bin/bash
# round-2 GPU call 2: old-kernel control, ordered-L1 build, ncu capture of the new fused kernel, MUFU interference ubench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 tools/ubench/ubench2 > gpurun_out/r2_ubench2.txt 2>&1
for v in old v1 b200 v2; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_$v.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inner_loop_kernel -c 1 -f -o gpurun_out/r2_prof_inner \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_prof_bench.log 2>&1
cat gpurun_out/r2_ubench2.txt
for v in old v1 b200 v2; do python - <<PY
import json
f="gpurun_out/r2b_bench_$v.log"
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
    print("$v", "%.2fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], d["config"].get("resident_warp_slots"), d["clocks"])
except Exception as e:
    print("$v", "FAILED", e, open(f).read()[-500:])
PY
done
ls -la gpurun_out | tail -5
