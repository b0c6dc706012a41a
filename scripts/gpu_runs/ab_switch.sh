#!/bin/bash
# A/B of one execution-model switch of the fused kernels, same box, same run:  ab_switch.sh DBOA_OPERAND_TMEM | DBOA_CHAIN_FLAGS
# (1) the network-level parity tests with the switch on, (2) forward times at batch 1 / 2 / 9 with it off and on, (3) C2 (and C3)
# bench lines off / on / off / on
cd "$GRAFT_REPO_ROOT"
V=${1:-DBOA_OPERAND_TMEM}
env $V=1 timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_hmr.py tests/test_gpu_adapt.py -x -q 2>&1 | tail -${TAIL:-3}
for v in 0 1; do
  echo "== $V=$v"
  env $V=$v FWD_FUSED_ONLY=1 timeout 300 python scripts/fwd_time.py 2>&1 | grep "l2_flushed=True"
done
for v in 0 1 0 1; do
  env $V=$v timeout 600 python bench.py --no-cpu-baseline --steps 40 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V=$v C2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'fwd ms', round(d['roofline']['ms_per_launch'],4))"
done
for v in 0 1; do
  env $V=$v timeout 600 python bench.py --no-cpu-baseline --workload c3 --steps 12 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V=$v C3', round(d['value'],2), 'e2e', round(d['e2e']['value'],2))"
done
