#!/bin/bash
cd "$GRAFT_REPO_ROOT"
DBOA_LIB_PATH=$PWD/dynaboa_b200/build/libdboa_timeline.so timeout 600 python scripts/fused_timeline.py 1 > gpurun_out/r02_timeline_b1.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_hmr.py -x -q 2>&1 | tail -2
timeout 300 python scripts/fwd_time.py 2>&1 | grep "fused=1 l2_flushed=True"
timeout 600 python bench.py --no-cpu-baseline --steps 60 --warmup 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'fwd ms', round(d['roofline']['ms_per_launch'],4))"
