#!/bin/bash
# CTA budgets of the data-gradient chain and of the weight gradients on the side streams (C2 frames/s)
cd "$GRAFT_REPO_ROOT"
run() { timeout 600 python bench.py --no-cpu-baseline --steps 60 --warmup 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; }
for w in 48 64 80 96; do DBOA_WGRAD_MAX_CTAS=$w run "dgrad=64 wgrad=$w"; done
DBOA_DGRAD_MAX_CTAS=80 DBOA_WGRAD_MAX_CTAS=64 run "dgrad=80 wgrad=64"
DBOA_WGRAD_STREAMS=1 run "dgrad=64 wgrad=128 one side stream"
DBOA_WGRAD_STREAMS=1 DBOA_WGRAD_MAX_CTAS=64 run "dgrad=64 wgrad=64 one side stream"
