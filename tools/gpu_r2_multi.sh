#!/bin/bash
# round-2 multi-GPU call (gpurun --gpus 2): NCCL parity of GTN_Master (replicated / allreduce update) and bench.py on 2 ranks
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2_multi_gpu.txt
{
echo "# gpurun --gpus 2 -- bash tools/gpu_r2_multi.sh"
nvidia-smi -L
echo "## python -m pytest tests/test_gpu_multi.py -m gpu -q"
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6
for sc in weak strong; do
  echo "## torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --extras none --scaling $sc"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 2 \
      --no-cpu-baseline --extras none --scaling $sc 2>&1 | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','nes_population')}, 'with_update', d['with_update']['value'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])"
done
echo "## single GPU, same box: bench.py --steps 3 --warmup 2 --no-cpu-baseline --extras none (weak) / --scaling strong"
for sc in weak strong; do
  timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --extras none --scaling $sc 2>&1 | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','nes_population')}, 'with_update', d['with_update']['value'], 'e2e', d['e2e']['value'])"
done
} > $O 2>&1
cat $O
