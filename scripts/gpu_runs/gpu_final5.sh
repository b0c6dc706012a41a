mkdir -p gpurun_out
timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final_nocpu.log 2>/dev/null
tail -1 gpurun_out/bench_final_nocpu.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"ms_per_launch": [0-9.]*' | tr '\n' ' '
