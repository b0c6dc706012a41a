mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adapt.py -m gpu -q -x > gpurun_out/t_adapt.log 2>&1
tail -2 gpurun_out/t_adapt.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
DBOA_HIPRI=1 $B 2>/dev/null | grep -o '"value": [0-9.]*' | head -2
DBOA_HIPRI=0 $B 2>/dev/null | grep -o '"value": [0-9.]*' | head -2
DBOA_HIPRI=1 $B 2>/dev/null | grep -o '"value": [0-9.]*' | head -2
DBOA_HIPRI=0 $B 2>/dev/null | grep -o '"value": [0-9.]*' | head -2
