#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2ag
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 6 python tools/race_loop.py mwc 2>&1 | grep -v "^=========     Host Frame\|^=========         in \|^=========                in" | head -120 > gpurun_out/r2ag/race_detail.txt
head -100 gpurun_out/r2ag/race_detail.txt
