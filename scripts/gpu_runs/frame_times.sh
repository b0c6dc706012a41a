#!/bin/bash
# per-frame times of the timed region of the default bench run (K = 20, W = 3), with and without the NVML sampler thread
cd "$GRAFT_REPO_ROOT"
for a in "" "--no-clock-sampler" ""; do
  timeout 600 python bench.py --no-cpu-baseline --frame-times $a 2> gpurun_out/ft.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$a] C2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
  grep "frame times\|ms/frame" gpurun_out/ft.err
done
