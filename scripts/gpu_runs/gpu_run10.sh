mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
tail -5 gpurun_out/t_all.log
timeout 300 python scripts/trace_step.py --tag r02 --region frame > gpurun_out/trace_frame.log 2>&1
timeout 300 python scripts/trace_step.py --tag r02 --region fwdbwd > gpurun_out/trace_fwdbwd.log 2>&1
DBOA_ASYNC_WGRAD=0 timeout 300 python scripts/trace_step.py --tag r02s --region fwdbwd > gpurun_out/trace_fwdbwd_sync.log 2>&1
head -50 gpurun_out/trace_frame.log
