#!/bin/bash
cd "$GRAFT_REPO_ROOT"
run() { timeout 600 python bench.py --no-cpu-baseline --steps 60 --warmup 8 $2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; }
DBOA_TEACHER_PREFETCH=0 run "prefetch=0"
DBOA_TEACHER_PREFETCH=1 run "prefetch=1"
DBOA_TEACHER_PREFETCH=1 DBOA_FUSED_MAX_CTAS=128 run "prefetch=1 fwd_ctas=128"
DBOA_TEACHER_PREFETCH=1 DBOA_FUSED_MAX_CTAS=64 run "prefetch=1 fwd_ctas=64"
timeout 900 python -m pytest tests/test_gpu_adapt.py -x -q 2>&1 | tail -2
