#!/bin/bash
# K-slice / CTA-budget knobs of the fused kernels re-swept with the tensor-memory operand (the k-loop got cheaper: fewer, longer
# K-slices may now pay): forward time at batch 1 / 2 / 9, then C2 frames/s for the dgrad budget
cd "$GRAFT_REPO_ROOT"
fw() { echo "== $1"; FWD_FUSED_ONLY=1 timeout 300 python scripts/fwd_time.py 2>&1 | grep "l2_flushed=True" | sed 's/ fused=1 l2_flushed=True://; s/  launches.*//' | tr '\n' ' '; echo; }
fw "default (minkb 2, budget 96)"
DBOA_FUSED_MINKB=3 fw "minkb 3"
DBOA_FUSED_MINKB=4 fw "minkb 4"
DBOA_FUSED_MAX_CTAS=128 fw "budget 128"
DBOA_FUSED_MAX_CTAS=148 fw "budget 148"
DBOA_FUSED_MAX_CTAS=64 fw "budget 64"
DBOA_FUSED_MAX_NZ=4 fw "max nz 4"
DBOA_FUSED_MAX_NZ=8 fw "max nz 8"
run() { timeout 600 python bench.py --no-cpu-baseline --steps 60 --warmup 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; }
run "C2 default"
DBOA_DGRAD_MAX_CTAS=96 run "C2 dgrad=96"
DBOA_DGRAD_MAX_CTAS=48 run "C2 dgrad=48"
DBOA_FUSED_MAX_CTAS=128 run "C2 fwd budget 128"
run "C2 default again"
