#!/bin/bash
# the whole GPU suite, then C2 / C3 bench lines (no CPU baseline)
cd "$GRAFT_REPO_ROOT"
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
run() { timeout 600 python bench.py --no-cpu-baseline --steps 60 --warmup 8 $2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'fwd ms', round(d['roofline']['ms_per_launch'],4), 'launches', d['gpu_launches'])"; }
run c2
DBOA_DGRAD_MAX_CTAS=32 run "c2 dgrad32"
DBOA_DGRAD_MAX_CTAS=48 run "c2 dgrad48"
run c3 "--workload c3 --steps 20 --warmup 4"
