#!/bin/bash
# one full ncu capture (with source) of a kernel: $1 = kernel regex, $2 = launches to skip, $3 = region of scripts/profile_step.py, $4 = output stem
cd "$GRAFT_REPO_ROOT"
timeout 600 ncu --clock-control none --profile-from-start off --set full --import-source on -k regex:$1 --launch-skip $2 --launch-count 1 -o gpurun_out/$4 python scripts/profile_step.py --region $3 --tc 3 > gpurun_out/ncu_one.log 2>&1
tail -2 gpurun_out/ncu_one.log
