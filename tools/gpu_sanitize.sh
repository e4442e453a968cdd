#!/bin/bash
# compute-sanitizer over the GPU parity tests (memcheck on everything small, racecheck on the TD kernels). -> gpurun_out/sanitizer.txt
set -u
mkdir -p gpurun_out
{
echo '# compute-sanitizer on the GPU parity tests (B200)'
echo 'compute-sanitizer --tool memcheck python -m pytest tests -m gpu -k "se_forward or rn_reward or td_update or lockstep or nes_noise or general or h1024 or host_buffer or edge or vary"'
if [ -z "${SKIP_MEMCHECK:-}" ]; then timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x \
    -k "se_forward or rn_reward or td_update or lockstep or nes_noise or general or h1024 or host_buffer or edge or vary" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -8; fi
echo 'compute-sanitizer --tool racecheck python -m pytest tests -m gpu -k "td_update_vs_reference_golden"   (unit kernels only: whole inner loops take >25 min under racecheck)'
timeout ${RACE_TIMEOUT:-400} compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x \
    -k "td_update_vs_reference_golden" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -8
} | tee gpurun_out/sanitizer.txt
