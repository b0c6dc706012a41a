#!/bin/bash
# A/B of the tensor-memory activation operand (DBOA_OPERAND_TMEM): parity tests of the fused kernels with it on, forward times at
# batch 1 / 2 / 9 and C2 / C3 bench lines with it off and on (same box, same run)
cd "$GRAFT_REPO_ROOT"
DBOA_OPERAND_TMEM=1 timeout 900 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -${TAIL:-8}
for v in 0 1; do
  echo "== DBOA_OPERAND_TMEM=$v"
  DBOA_OPERAND_TMEM=$v FWD_FUSED_ONLY=1 timeout 300 python scripts/fwd_time.py 2>&1 | tail -6
done
for v in 0 1 0 1; do
  DBOA_OPERAND_TMEM=$v timeout 600 python bench.py --no-cpu-baseline --steps 40 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tmem=$v C2', d['value'], 'e2e', d['e2e']['value'], 'fwd ms', d['roofline']['ms_per_launch'])"
done
for v in 0 1; do
  DBOA_OPERAND_TMEM=$v timeout 600 python bench.py --no-cpu-baseline --workload c3 --steps 12 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tmem=$v C3', d['value'], 'e2e', d['e2e']['value'])"
done
