#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for v in 0 96 74 64; do
  echo "=== DBOA_FUSED_MAX_CTAS=$v"; DBOA_FUSED_MAX_CTAS=$v timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'])"
done
echo "=== unfused"; DBOA_FUSED_FWD=0 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'])"
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -m gpu 2>&1 | tail -5
