#!/bin/bash
# the GPU suite N times in the form the driver runs it (complete logs kept): looks for intermittent failures
cd "$GRAFT_REPO_ROOT"
for i in $(seq 1 ${1:-3}); do
  timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/suite_$i.log 2>&1
  echo "run $i rc=$?"
  grep -n "passed\|failed\|Fatal\|Segmentation\|Abort\|core" gpurun_out/suite_$i.log | tail -5 | cut -c1-200
done
