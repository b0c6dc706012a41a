#!/bin/bash
# A/B of the counter-based dependency between fused forward launches (DBOA_CHAIN_FLAGS): network-level parity tests with it on,
# forward times and C2 bench lines with it off and on (same box, same run)
cd "$GRAFT_REPO_ROOT"
DBOA_CHAIN_FLAGS=1 timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_hmr.py tests/test_gpu_adapt.py -x -q 2>&1 | tail -${TAIL:-8}
for v in 0 1; do
  echo "== DBOA_CHAIN_FLAGS=$v"
  DBOA_CHAIN_FLAGS=$v FWD_FUSED_ONLY=1 timeout 300 python scripts/fwd_time.py 2>&1 | grep "l2_flushed=True"
done
for v in 0 1 0 1; do
  DBOA_CHAIN_FLAGS=$v timeout 600 python bench.py --no-cpu-baseline --steps 40 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chain=$v C2', d['value'], 'e2e', d['e2e']['value'], 'fwd ms', d['roofline']['ms_per_launch'])"
done
