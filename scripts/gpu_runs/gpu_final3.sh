mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
tail -2 gpurun_out/t_all.log
timeout 300 python bench.py > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err
tail -1 gpurun_out/bench_full.log | cut -c1-260
