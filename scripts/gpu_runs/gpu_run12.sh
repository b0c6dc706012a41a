mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
tail -5 gpurun_out/t_all.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
DBOA_PDL=0 $B > gpurun_out/bench_pdl0.log 2>&1
DBOA_PDL=1 $B > gpurun_out/bench_pdl1.log 2>&1
for f in pdl0 pdl1; do echo $f; grep -o '"value": [0-9.]*' gpurun_out/bench_$f.log | head -2; grep -o '"ms_per_launch": [0-9.]*' gpurun_out/bench_$f.log; done
DBOA_ASYNC_WGRAD=0 timeout 300 python scripts/trace_step.py --tag r03s --region fwdbwd > gpurun_out/trace_fwdbwd_sync.log 2>&1
timeout 300 python scripts/trace_step.py --tag r03 --region frame > gpurun_out/trace_frame.log 2>&1
grep -v Warn gpurun_out/trace_fwdbwd_sync.log | head -14
grep -v Warn gpurun_out/trace_frame.log | head -12
