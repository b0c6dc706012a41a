mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adapt.py -m gpu -q -x > gpurun_out/t_adapt.log 2>&1
tail -3 gpurun_out/t_adapt.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/bench_a.log 2>gpurun_out/bench_a.err
$B --serial-output > gpurun_out/bench_b.log 2>&1
for f in a b; do echo $f; grep -o '"value": [0-9.]*' gpurun_out/bench_$f.log | head -2; grep -o '"clocks": {[^}]*}' gpurun_out/bench_$f.log; done
tail -3 gpurun_out/bench_a.err
