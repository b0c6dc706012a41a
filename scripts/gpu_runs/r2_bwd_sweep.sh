#!/bin/bash
cd "$GRAFT_REPO_ROOT"
run() { timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"; }
for dg in 32 64 96; do for wg in 16 48 128; do echo "DGRAD=$dg WGRAD=$wg"; DBOA_DGRAD_MAX_CTAS=$dg DBOA_WGRAD_MAX_CTAS=$wg run; done; done
echo "fused bwd, wgrad on CUDA cores"; DBOA_WGRAD_TMA=0 DBOA_DGRAD_MAX_CTAS=96 run
echo "fused bwd, async wgrad off"; DBOA_ASYNC_WGRAD=0 run
