#!/bin/bash
# default bench.py line (all extra workloads) under torchrun on 2 GPUs, as the driver's scaling run launches it
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2ac
O=gpurun_out/r2ac
S=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench2_default.log 2> $O/bench2_default.err
echo "exit $? wall seconds: $(( $(date +%s) - S ))" | tee $O/bench2_default.time
tail -c 400 $O/bench2_default.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2ac/bench2_default.log') if x.startswith('{')][-1]; d=json.loads(l)
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['roofline']['frac'], 'e2e', d['e2e']['value'], 'with_update', d['with_update']['value'])
for k,v in d['workloads'].items(): print(k, v.get('value'), v.get('frac'), v.get('ms_per_step'), v.get('error'))
print('strong', d['strong_scaling'].get('value'))
PY
