#!/bin/bash
# round 2, first GPU contact of the fused forward: unit parity of conv_fused.cu, then whole-forward timing fused vs unfused
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_fused_tests.log
cat gpurun_out/r2_fused_tests.log
timeout 300 python scripts/fwd_time.py 2>&1 | tail -20 | tee gpurun_out/r2_fwd_time.log
