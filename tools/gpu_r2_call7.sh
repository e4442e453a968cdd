#!/bin/bash
# round-2 GPU call 7: full GPU suite (3-hidden-layer goldens, TD3 loss parity) + the default bench line with every workload
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2g_pytest.log
timeout 1500 python bench.py > gpurun_out/r2g_bench_default.log 2> gpurun_out/r2g_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_bench_reference.log 2>&1
tail -40 gpurun_out/r2g_pytest.log
tail -5 gpurun_out/r2g_bench_default.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2g_bench_default.log') if x.startswith('{')][-1]
d=json.loads(l)
print("headline %.2fM frac %.3f upd %.2fM e2e %.2fM gph %.0f"%(d['value']/1e6,d['roofline']['frac'],d['with_update']['value']/1e6,d['e2e']['value']/1e6,d['nes_generations_per_hour']))
for k,v in d['workloads'].items():
    print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in('value','frac','ms_per_step','error','seconds_per_evaluation','nes_generations_per_hour','us_per_env_step_per_lane')})
print('strong', d['strong_scaling'].get('value'), d['strong_scaling'].get('ms_per_step'), d['strong_scaling'].get('error'))
print('cpu', {k:v for k,v in d['cpu_baseline'].items() if k!='sample'})
PY
tail -c 1500 gpurun_out/r2g_bench_reference.log
