#!/bin/bash
# largest cluster (split-K slices per tile) of the three wide kernels against C2 frames/s: big clusters need free SMs in one GPC
cd "$GRAFT_REPO_ROOT"
run() { timeout 600 python bench.py --no-cpu-baseline --steps 60 --warmup 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'fwd', round(d['roofline']['ms_per_launch'],4))"; }
run "default (16/16/16)"
DBOA_WGRAD_MAX_NZ=4 run "wgrad nz<=4"
DBOA_WGRAD_MAX_NZ=2 run "wgrad nz<=2"
DBOA_WGRAD_MAX_NZ=4 DBOA_DGRAD_MAX_NZ=4 run "wgrad,dgrad nz<=4"
DBOA_WGRAD_MAX_NZ=4 DBOA_DGRAD_MAX_NZ=4 DBOA_FUSED_MAX_NZ=4 run "all nz<=4"
DBOA_WGRAD_MAX_NZ=8 DBOA_DGRAD_MAX_NZ=8 DBOA_FUSED_MAX_NZ=8 run "all nz<=8"
DBOA_WGRAD_MAX_NZ=4 DBOA_WGRAD_MAX_CTAS=64 run "wgrad nz<=4 ctas 64"
