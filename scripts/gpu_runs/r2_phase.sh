#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python scripts/phase_times.py 2>&1 | tail -13 | tee gpurun_out/r02b_phase_times.txt
timeout 900 ncu --clock-control none --profile-from-start off --csv --metrics gpu__time_duration.sum --log-file gpurun_out/r02b_launches_frame_c2.csv python scripts/profile_step.py --region frame --tc 3 > gpurun_out/p1.log 2>&1
tail -2 gpurun_out/p1.log
