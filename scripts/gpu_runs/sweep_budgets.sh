#!/bin/bash
# CTA budgets of the data-gradient chain and of the weight gradients on the side streams (C2 frames/s)
cd "$GRAFT_REPO_ROOT"
if [ "$1" != forward ]; then
run() { timeout 600 python bench.py --no-cpu-baseline --steps 60 --warmup 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; }
for w in 48 64 80 96; do DBOA_WGRAD_MAX_CTAS=$w run "dgrad=64 wgrad=$w"; done
DBOA_DGRAD_MAX_CTAS=80 DBOA_WGRAD_MAX_CTAS=64 run "dgrad=80 wgrad=64"
DBOA_WGRAD_STREAMS=1 run "dgrad=64 wgrad=128 one side stream"
DBOA_WGRAD_STREAMS=1 DBOA_WGRAD_MAX_CTAS=64 run "dgrad=64 wgrad=64 one side stream"
fi
# ---- K-slice / forward-budget knobs of the fused forward (re-swept with the tensor-memory operand: defaults unchanged): forward
# time at batch 1 / 2 / 9, then C2 frames/s; `sweep_budgets.sh forward` runs only this part
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
