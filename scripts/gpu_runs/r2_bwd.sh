#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -k "fused_data_gradient" 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_gpu_hmr.py tests/test_gpu_adapt.py tests/test_gpu_dp.py -x -q -m gpu 2>&1 | grep -v "Warning\|warn\|grad.sizes\|param.sizes\|return Variable\|^$\|Docs:" | tail -25 | tee gpurun_out/r2_bwd_tests.log
for v in 1 0; do echo "DBOA_FUSED_BWD=$v"; DBOA_FUSED_BWD=$v timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; done
