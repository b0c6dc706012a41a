mkdir -p gpurun_out
DBOA_TIMELINE=1 python -m dynaboa_b200.build --force > gpurun_out/build_tl.log 2>&1
timeout 300 python scripts/kernel_timeline.py > gpurun_out/timeline.log 2>&1
cat gpurun_out/timeline.log | grep -v "median/max\|^   prologue" 
