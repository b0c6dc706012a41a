mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/t_all.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nocpu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_frame.csv python scripts/profile_step.py --region frame > gpurun_out/prof_frame.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_fwd_tc.csv python scripts/profile_step.py --region forward --tc 1 > gpurun_out/prof_fwd_tc.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_fwd.csv python scripts/profile_step.py --region forward --tc 0 > gpurun_out/prof_fwd.log 2>&1
grep -E "passed|failed" gpurun_out/t_all.log | tail -3; tail -c 1500 gpurun_out/bench_nocpu.log
