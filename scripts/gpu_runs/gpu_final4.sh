mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1
tail -2 gpurun_out/t_all.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
