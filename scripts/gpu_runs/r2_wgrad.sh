#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -k "weight_gradient" 2>&1 | tail -25 | tee gpurun_out/r2_wgrad_tests.log
timeout 900 python -m pytest tests/test_gpu_hmr.py tests/test_gpu_adapt.py -x -q -m gpu 2>&1 | tail -8
for v in 1 0; do echo "DBOA_WGRAD_TMA=$v"; DBOA_WGRAD_TMA=$v timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; done
