mkdir -p gpurun_out
timeout 300 python scripts/phase_times.py > gpurun_out/phase_times.log 2>&1
cat gpurun_out/phase_times.log | tail -16
