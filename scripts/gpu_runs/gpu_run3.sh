mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > gpurun_out/t_tc.log 2>&1
timeout 900 python -m pytest tests/test_gpu_hmr.py tests/test_gpu_adapt.py -m gpu -q -s > gpurun_out/t_adapt.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_frame.csv python scripts/profile_step.py --region frame > gpurun_out/prof_frame.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_fwdbwd.csv python scripts/profile_step.py --region fwdbwd > gpurun_out/prof_fwdbwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_fwd_kernel -c 6 -o gpurun_out/prof_conv_fwd python scripts/profile_step.py --region forward > gpurun_out/prof_full.log 2>&1
tail -4 gpurun_out/t_tc.log; tail -4 gpurun_out/t_adapt.log; tail -2 gpurun_out/prof_frame.log; ls -la gpurun_out | tail -8
