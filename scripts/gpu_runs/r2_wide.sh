#!/bin/bash
# round 2: fused "wide" convolution (TMA for both operands): unit parity, whole-forward parity, timing
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2_wide_tests.log
timeout 300 python scripts/fwd_time.py 2>&1 | tail -14 | tee gpurun_out/r2_wide_fwd_time.log
