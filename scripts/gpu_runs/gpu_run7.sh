mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -x > gpurun_out/t_tc.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --tc 0 > gpurun_out/bench_tc0.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --tc 1 > gpurun_out/bench_tc1.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --tc 2 > gpurun_out/bench_tc2.log 2>&1
grep -E "passed|failed|Error" gpurun_out/t_tc.log | tail -4
for f in gpurun_out/bench_tc0.log gpurun_out/bench_tc1.log gpurun_out/bench_tc2.log; do grep -o '"value": [0-9.]*' $f | head -1; grep -o '"ms_per_launch": [0-9.]*' $f; done
