mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
tail -3 gpurun_out/t_all.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err
tail -1 gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err
tail -1 gpurun_out/bench_ref.log
timeout 300 python scripts/trace_step.py --tag r05 --region frame > gpurun_out/trace_frame.log 2>&1
grep -v Warn gpurun_out/trace_frame.log | head -45
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
