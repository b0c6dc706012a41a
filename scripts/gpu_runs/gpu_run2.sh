mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_adapt.py > gpurun_out/t_kernels.log 2>&1
timeout 900 python -m pytest tests/test_gpu_adapt.py -m gpu -q -s > gpurun_out/t_adapt.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nocpu.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -3 gpurun_out/t_kernels.log; tail -3 gpurun_out/t_adapt.log; tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/bench_nocpu.log; tail -2 gpurun_out/bench.log; tail -2 gpurun_out/bench_ref.log
