#!/bin/bash
# after a change to the fused forward kernel: its parity tests, forward times at b = 1, 2, 9, a short C2 bench
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -5
timeout 300 python scripts/fwd_time.py 2>&1 | tail -8
timeout 600 python bench.py --no-cpu-baseline --steps 40 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2', d['value'], 'e2e', d['e2e']['value'], 'fwd ms', d['roofline']['ms_per_launch'])"
