#!/bin/bash
# fused backward: repeatability of value / e2e, and the CTA budget of the fused data gradient
cd "$GRAFT_REPO_ROOT"
run() { timeout 600 python bench.py --no-cpu-baseline --steps 60 --warmup 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; }
DBOA_FUSED_BWD=1 run "fb=1 a"
DBOA_FUSED_BWD=1 run "fb=1 b"
DBOA_FUSED_BWD=0 run "fb=0 a"
for c in 64 96; do DBOA_FUSED_BWD=1 DBOA_DGRAD_MAX_CTAS=$c run "fb=1 dgrad_ctas=$c"; done
DBOA_FUSED_BWD=1 DBOA_WGRAD_MAX_CTAS=64 run "fb=1 wgrad_ctas=64"
