#!/bin/bash
# round-2 final check of the shipped tree: smoke(), full GPU suite, default bench.py (all workloads), reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest.log; tail -3 $O/pytest.log
S=$(date +%s)
timeout 1500 python bench.py > $O/bench_default.log 2> $O/bench_default.err
echo "default bench wall seconds: $(( $(date +%s) - S ))" | tee $O/bench_default.time
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2z/bench_default.log') if x.startswith('{')][-1]; d=json.loads(l)
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'].get('value'), d['cpu_baseline'].get('kind'), 'gen/h', d.get('nes_generations_per_hour'))
for k,v in d['workloads'].items(): print(k, v.get('value'), v.get('frac'), v.get('ms_per_step'), v.get('error'))
print('strong', d['strong_scaling'].get('value'))
PY
