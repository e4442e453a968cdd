#!/bin/bash
# round-2 GPU call 26: compute-sanitizer memcheck over the GPU parity tests with the shipped kernels (compact reduction, multi-warp lanes, staged pack)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2aa
O=gpurun_out/r2aa
{
echo '# compute-sanitizer --tool memcheck python -m pytest tests -m gpu -k "se_forward or rn_reward or td_update or lockstep or nes_noise or h1024 or host_buffer or edge or vary or multi_warp"   (second session of round 2)'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x \
    -k "se_forward or rn_reward or td_update or lockstep or nes_noise or h1024 or host_buffer or edge or vary or multi_warp" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|rror" | tail -8
echo "exit code: ${PIPESTATUS[0]}"
} > $O/sanitizer_memcheck.txt 2>&1
cat $O/sanitizer_memcheck.txt
