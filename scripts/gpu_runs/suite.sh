#!/bin/bash
# the whole GPU suite with the complete log kept (gpurun_out/suite.log); optional pytest arguments
cd "$GRAFT_REPO_ROOT"
timeout 1800 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider "$@" > gpurun_out/suite.log 2>&1
echo "rc=$?"
grep -n "passed\|failed\|error\|Fatal\|Error" gpurun_out/suite.log | tail -20
tail -5 gpurun_out/suite.log | cut -c1-300
