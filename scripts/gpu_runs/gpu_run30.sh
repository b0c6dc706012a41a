mkdir -p gpurun_out
DBOA_PDL=0 DBOA_ASYNC_WGRAD=0 DBOA_TEACHER_STREAM=0 timeout 300 python scripts/trace_step.py --tag r07s --region frame > gpurun_out/trace_frame_serial.log 2>&1
grep -v Warn gpurun_out/trace_frame_serial.log | head -48
