mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1
tail -3 gpurun_out/t_all.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
DBOA_ASYNC_WGRAD=0 DBOA_TEACHER_STREAM=0 $B > gpurun_out/bench_s00.log 2>&1
DBOA_ASYNC_WGRAD=1 DBOA_TEACHER_STREAM=0 $B > gpurun_out/bench_s10.log 2>&1
DBOA_ASYNC_WGRAD=0 DBOA_TEACHER_STREAM=1 $B > gpurun_out/bench_s01.log 2>&1
DBOA_ASYNC_WGRAD=1 DBOA_TEACHER_STREAM=1 $B > gpurun_out/bench_s11.log 2>&1
for f in s00 s10 s01 s11; do echo $f; grep -o '"value": [0-9.]*' gpurun_out/bench_$f.log | head -2; done
