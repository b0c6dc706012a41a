#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_adapt.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r2_dp_tests.log
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'])"
timeout 900 python bench.py --no-cpu-baseline --workload c3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 900 python bench.py --no-cpu-baseline --workload c5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['dynamic_steps'])"
