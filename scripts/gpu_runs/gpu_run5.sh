mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -x > gpurun_out/t_all.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nocpu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_frame.csv python scripts/profile_step.py --region frame > gpurun_out/prof_frame.log 2>&1
grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -5; tail -c 1200 gpurun_out/bench_nocpu.log
