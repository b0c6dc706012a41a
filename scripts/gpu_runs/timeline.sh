#!/bin/bash
# diagnostic build (in-kernel phase stamps) into a scratch library, timeline of one forward
cd "$GRAFT_REPO_ROOT"
B=${1:-1}
DBOA_LIB_PATH=$PWD/dynaboa_b200/build/libdboa_timeline.so timeout 600 python scripts/fused_timeline.py $B 2>&1 | tee gpurun_out/r02_timeline_b$B.txt | head -60
