mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
tail -2 gpurun_out/t_all.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$B 2>/dev/null | grep -o '"value": [0-9.]*' | head -2 | tr '\n' ' '; echo
$B 2>/dev/null | grep -o '"value": [0-9.]*' | head -2 | tr '\n' ' '; echo
