#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests/test_gpu_fused.py -x -q -m gpu 2>&1 | grep -v "^E  .*tensor\|^ *$" | tail -30 | tee gpurun_out/r2_s2_tests.log
timeout 300 python scripts/fwd_time.py 2>&1 | grep "l2_flushed=True" | tee gpurun_out/r2_s2_fwd.log
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['ms_per_launch'])"
