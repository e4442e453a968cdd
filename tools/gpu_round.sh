#!/bin/bash
# One GPU session: parity tests, smoke, bench, launch list. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -${PYTEST_TAIL:-60} | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup ${BENCH_WARMUP:-3} ${BENCH_ARGS:-} 2>&1 | tail -5 | tee gpurun_out/bench.log
