#!/bin/bash
# after a change to the backward kernels: parity tests of the fused kernels and of the HMR backward, C2 bench with the fused backward off / on
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_hmr.py -x -q 2>&1 | tail -5
for fb in 0 1; do
DBOA_FUSED_BWD=$fb timeout 600 python bench.py --no-cpu-baseline --steps 40 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2 fused_bwd=$fb', d['value'], 'e2e', d['e2e']['value'], 'fwd ms', d['roofline']['ms_per_launch'])"
done
