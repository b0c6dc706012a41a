mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
tail -2 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err
tail -1 gpurun_out/bench_full.log | cut -c1-260
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err
tail -1 gpurun_out/bench_ref.log | cut -c1-200
timeout 300 python scripts/phase_times.py > gpurun_out/phase_times.log 2>&1
tail -12 gpurun_out/phase_times.log
