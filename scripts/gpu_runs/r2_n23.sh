#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests/test_dataprocess.py tests/test_gpu_internet.py tests/test_gpu_dropin_driver.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r2_n23_tests.log
