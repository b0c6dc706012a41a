#!/bin/bash
# round 2: where does a fused-forward launch spend its time?  (diagnostic DBOA_TIMELINE build; two register-budget variants)
cd "$GRAFT_REPO_ROOT"
echo "=== variant: 2 CTAs/SM register budget (96/108), smem <= 113 KB"
timeout 300 python scripts/fused_timeline.py 1 2>&1 | tail -60 | tee gpurun_out/r2_timeline_occ2.txt
timeout 300 python scripts/fwd_time.py 2>&1 | grep "B=1\|B=2" | tee gpurun_out/r2_fwd_time_occ2.txt
echo "=== variant: 1 CTA/SM registers (200)"
DBOA_LIB_PATH=$PWD/dynaboa_b200/libdboa_variant_1cta.so timeout 300 python scripts/fused_timeline.py 1 2>&1 | tail -60 | tee gpurun_out/r2_timeline_1cta.txt
DBOA_LIB_PATH=$PWD/dynaboa_b200/libdboa_variant_1cta.so timeout 300 python scripts/fwd_time.py 2>&1 | grep "B=1\|B=2" | tee gpurun_out/r2_fwd_time_1cta.txt
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_fused_tests2.log
