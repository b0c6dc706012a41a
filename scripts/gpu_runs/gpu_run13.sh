mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
tail -3 gpurun_out/t_all.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/bench_a.log 2>&1
DBOA_WGRAD_STREAMS=1 $B > gpurun_out/bench_b.log 2>&1
$B --tc 2 > gpurun_out/bench_c.log 2>&1
for f in a b c; do echo $f; grep -o '"value": [0-9.]*' gpurun_out/bench_$f.log | head -2; grep -o '"ms_per_launch": [0-9.]*' gpurun_out/bench_$f.log; done
DBOA_PDL=0 DBOA_ASYNC_WGRAD=0 timeout 300 python scripts/trace_step.py --tag r04s --region fwdbwd > gpurun_out/trace_fwdbwd_sync.log 2>&1
grep -v Warn gpurun_out/trace_fwdbwd_sync.log | head -14
