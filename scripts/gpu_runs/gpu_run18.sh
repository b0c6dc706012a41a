mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_hmr.py -m gpu -q -x > gpurun_out/t_tc.log 2>&1
tail -3 gpurun_out/t_tc.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/bench_a.log 2>&1
for f in a; do echo $f; grep -o '"value": [0-9.]*' gpurun_out/bench_$f.log | head -2; grep -o '"ms_per_launch": [0-9.]*' gpurun_out/bench_$f.log; done
DBOA_PDL=1 timeout 300 python scripts/conv_microbench.py > gpurun_out/conv_mb_pdl1.log 2>&1
cat gpurun_out/conv_mb_pdl1.log
DBOA_TIMELINE=1 python -m dynaboa_b200.build --force > gpurun_out/build_tl.log 2>&1
timeout 300 python scripts/kernel_timeline.py > gpurun_out/timeline.log 2>&1
cat gpurun_out/timeline.log
