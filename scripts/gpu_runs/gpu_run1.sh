mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_adapt.py 2>&1 | tail -60 > gpurun_out/t_kernels.log
timeout 900 python -m pytest tests/test_gpu_adapt.py -m gpu -q 2>&1 | tail -80 > gpurun_out/t_adapt.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
tail -5 gpurun_out/t_kernels.log; tail -5 gpurun_out/t_adapt.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log
