#!/bin/bash
# round 2: whole GPU test-suite + forward timing + bench on the fused ("wide") plan
cd "$GRAFT_REPO_ROOT"
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2_suite_tests.log
timeout 300 python scripts/fwd_time.py 2>&1 | tail -14 | tee gpurun_out/r2_suite_fwd_time.log
timeout 600 python bench.py 2>&1 | tail -3 | tee gpurun_out/r2_suite_bench.log
