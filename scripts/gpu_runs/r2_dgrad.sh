#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -k "fused_data_gradient" 2>&1 | tail -30 | tee gpurun_out/r2_dgrad_tests.log
