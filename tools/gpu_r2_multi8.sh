#!/bin/bash
# 4-GPU check: bench.py weak / fixed-population under torchrun
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2_multi8_gpu.txt
{
echo "# gpurun --gpus 8 -- bash tools/gpu_r2_multi4.sh"
nvidia-smi -L | sed 's/UUID: GPU-[0-9a-f-]*/UUID: GPU-REDACTED/'
for sc in weak strong; do
  echo "## torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline --extras none --scaling $sc"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29699 bench.py --gpus 8 --steps 3 --warmup 2 \
      --no-cpu-baseline --extras none --scaling $sc 2>&1 | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','nes_population')}, 'with_update', d['with_update']['value'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])"
done
} > $O 2>&1
cat $O
