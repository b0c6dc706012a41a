mkdir -p gpurun_out
nvcc -arch=sm_100a -O3 -o /tmp/load_latency scripts/experiments/load_latency.cu && /tmp/load_latency > gpurun_out/load_latency.log 2>&1
cat gpurun_out/load_latency.log
