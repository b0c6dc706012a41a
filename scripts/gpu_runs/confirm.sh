#!/bin/bash
# after a host-side fix: the GPU suite once (complete log kept), smoke(), two default bench lines
cd "$GRAFT_REPO_ROOT"
timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/confirm_suite.log 2>&1; echo "suite rc=$?"; tail -1 gpurun_out/confirm_suite.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for i in 1 2; do timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/confirm_bench_$i.json; python -c "import json; d=json.load(open('gpurun_out/confirm_bench_$i.json')); print('C2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'fwd ms', round(d['roofline']['ms_per_launch'],4), 'launches', d['gpu_launches'], d['clocks'])"; done
