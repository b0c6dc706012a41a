#!/bin/bash
# end-of-round verification on one B200: whole GPU suite, smoke(), the driver's bench line (with the CPU baseline), the reference arm,
# C3 / C5 lines; `bash scripts/gpu_runs/final.sh evidence` also runs scripts/gpu_runs/evidence.sh
cd "$GRAFT_REPO_ROOT"
timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_tests_full.log 2>&1; echo "suite rc=$?" | tee gpurun_out/final_tests.log      # the complete log is kept: a crash must be attributable to a test
grep -n "passed\|failed\|Fatal\|Segmentation\|Abort\|core" gpurun_out/final_tests_full.log | tail -4 | cut -c1-200 | tee -a gpurun_out/final_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 1200 python bench.py 2> gpurun_out/final_bench_c2.err | tail -1 > gpurun_out/final_bench_c2.json
timeout 1200 python bench.py --impl reference --steps 4 --warmup 1 2> gpurun_out/final_bench_ref.err | tail -1 > gpurun_out/final_bench_c2_reference_arm.json
timeout 900 python bench.py --no-cpu-baseline --workload c3 2>/dev/null | tail -1 > gpurun_out/final_bench_c3.json
timeout 900 python bench.py --no-cpu-baseline --workload c5 2>/dev/null | tail -1 > gpurun_out/final_bench_c5.json
python - <<'PY'
import json
for f in ('c2', 'c2_reference_arm', 'c3', 'c5'):
    try:
        d = json.loads(open(f'gpurun_out/final_bench_{f}.json').read())
        r = d.get('roofline') or {}
        print(f, round(d['value'], 2), 'frames/s', 'e2e', round(d['e2e']['value'], 2), 'fwd ms', r.get('ms_per_launch'), 'frac', r.get('frac'),
              'cpu', (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('kind'), 'dyn', d['config'].get('dynamic_steps'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
if [ "$1" = evidence ]; then bash scripts/gpu_runs/evidence.sh > gpurun_out/final_evidence.log 2>&1; tail -14 gpurun_out/final_evidence.log; fi
