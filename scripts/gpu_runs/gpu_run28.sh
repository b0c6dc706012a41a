mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_hmr.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/t_tc.log 2>&1
tail -2 gpurun_out/t_tc.log
timeout 300 python scripts/conv_microbench.py > gpurun_out/conv_mb.log 2>&1
tail -1 gpurun_out/conv_mb.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$B 2>/dev/null | grep -o '"value": [0-9.]*' | head -2
$B 2>/dev/null | grep -o '"value": [0-9.]*' | head -2
