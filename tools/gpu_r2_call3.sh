#!/bin/bash
# round-2 GPU call 3: pipelined fused kernel — parity tests, A/B against the round-1 kernel, ncu capture, new bench.py smoke
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c_pytest.log
for v in old b200 p1; do
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2c_bench_$v.log 2>&1
  LE_LIB_NAME=lible_$v.so timeout 300 python bench.py --workload acrobot_se --steps 3 --warmup 2 --no-cpu-baseline --extras none > gpurun_out/r2c_bench_ac_$v.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inner_loop_kernel -c 1 -f -o gpurun_out/r2c_prof_inner \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --extras none > gpurun_out/r2c_prof_bench.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 2 > gpurun_out/r2c_bench_full.log 2>&1
tail -5 gpurun_out/r2c_pytest.log
for v in old b200 p1; do python - <<PY
import json
for f in ("gpurun_out/r2c_bench_$v.log","gpurun_out/r2c_bench_ac_$v.log"):
    try:
        l=[x for x in open(f) if x.startswith("{")][-1]; d=json.loads(l)
        print("$v", "ac" if "_ac_" in f else "cp", "%.2fM"%(d["value"]/1e6), "frac %.3f"%d["roofline"]["frac"], "upd %.2fM"%(d["with_update"]["value"]/1e6), "e2e %.2fM"%(d["e2e"]["value"]/1e6))
    except Exception as e:
        print("$v", f, "FAILED", e, open(f).read()[-800:])
PY
done
tail -c 3000 gpurun_out/r2c_bench_full.log
