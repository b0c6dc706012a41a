#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for mk in 2 3 4 6; do for mc in 64 96 128; do
  echo "=== MINKB=$mk MAX_CTAS=$mc"
  DBOA_FUSED_MINKB=$mk DBOA_FUSED_MAX_CTAS=$mc timeout 300 python scripts/fwd_time.py 2>&1 | grep "fused=1 l2_flushed=True"
done; done
for mk in 3 4; do
  echo "=== bench MINKB=$mk MAX_CTAS=96"; DBOA_FUSED_MINKB=$mk DBOA_FUSED_MAX_CTAS=96 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'])"
done
