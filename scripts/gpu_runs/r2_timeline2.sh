#!/bin/bash
# round 2: fine-grained stamps of the fused conv + the 2-CTAs-per-SM plan
cd "$GRAFT_REPO_ROOT"
echo "=== 96/108 registers, 1 CTA per SM plan"
timeout 300 python scripts/fused_timeline.py 1 2>&1 | tee gpurun_out/r2_tl2_a.txt | tail -50
echo "=== 96/108 registers, 2 CTAs per SM plan"
DBOA_FUSED_CTAS_PER_SM=2 timeout 300 python scripts/fused_timeline.py 1 2>&1 | tee gpurun_out/r2_tl2_b.txt | head -52
DBOA_FUSED_CTAS_PER_SM=2 timeout 300 python scripts/fwd_time.py 2>&1 | grep "fused=1" | tee gpurun_out/r2_fwd_time_b.txt
DBOA_FUSED_CTAS_PER_SM=2 DBOA_FUSED_MINKB=1 timeout 300 python scripts/fwd_time.py 2>&1 | grep "fused=1" | tee gpurun_out/r2_fwd_time_c.txt
