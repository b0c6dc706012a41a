mkdir -p gpurun_out
cp dynaboa_b200/libdynaboa_b200.so /tmp/lib_keep.so
DBOA_TIMELINE=1 python -m dynaboa_b200.build --force > gpurun_out/build_tl.log 2>&1
timeout 300 python scripts/kernel_timeline.py > gpurun_out/timeline.log 2>&1
cat gpurun_out/timeline.log
