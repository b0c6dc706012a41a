#!/bin/bash
# diagnostic library (scripts/build_timeline.sh): phase stamps of one forward, then forward times with parts of the fused kernel
# switched off (FWD_KNOBS: 1 no MMAs, 2 no transform body, 4 no cluster reduction; results are wrong, only the times are read)
cd "$GRAFT_REPO_ROOT"
export DBOA_LIB_PATH=$PWD/dynaboa_b200/build/libdboa_timeline.so
export DBOA_OPERAND_TMEM=${TMEM:-1}
timeout 600 python scripts/fused_timeline.py 1 > gpurun_out/timeline_tmem_b1.txt 2>&1
sed -n 1,12p gpurun_out/timeline_tmem_b1.txt | cut -c1-150
grep -A12 "^launch 1$" gpurun_out/timeline_tmem_b1.txt
for k in ${KNOBS:-0 1 2 3 4}; do
  echo "== knobs $k"
  FWD_KNOBS=$k FWD_FUSED_ONLY=1 timeout 300 python scripts/fwd_time.py 2>&1 | grep "l2_flushed=True"
done
