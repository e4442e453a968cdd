#!/bin/bash
# round-2 GPU call 10: row-owner TD update — full parity suite + ncu --set full (source-level samples) of the fused kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest.log
ARGS="--steps 1 --warmup 1 --no-cpu-baseline --extras none"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inner_loop_kernel -c 1 -f -o $O/prof_inner python bench.py $ARGS > $O/prof_bench.log 2>&1
tail -12 $O/pytest.log
tail -3 $O/prof_bench.log | cut -c1-300
ls -la $O
