#!/bin/bash
# after a change to the backward kernels: kernel-level and HMR-backward parity tests, C2 bench with the fused data-gradient chain on / off
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused.py tests/test_gpu_hmr.py -x -q 2>&1 | tail -5
for fb in 1 0; do
DBOA_FUSED_BWD=$fb timeout 600 python bench.py --no-cpu-baseline --steps 60 --warmup 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2 fused_bwd=$fb', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'fwd ms', round(d['roofline']['ms_per_launch'],4))"
done
