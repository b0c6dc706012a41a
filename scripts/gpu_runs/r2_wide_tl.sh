#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python scripts/fused_timeline.py 1 2>&1 | tee gpurun_out/r2_wide_timeline.txt | tail -110
