mkdir -p gpurun_out
S="--section SpeedOfLight --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis"
timeout 300 ncu $S --clock-control none --profile-from-start off -k regex:conv_fwd_kernel -s 20 -c 6 --csv --page raw --log-file gpurun_out/sec_convfwd.csv python scripts/profile_step.py --region fwdbwd > gpurun_out/p1.log 2>&1
timeout 300 ncu $S --clock-control none --profile-from-start off -k regex:gn_fwd_fused -s 20 -c 6 --csv --page raw --log-file gpurun_out/sec_gnfwd.csv python scripts/profile_step.py --region fwdbwd > gpurun_out/p2.log 2>&1
timeout 300 ncu $S --clock-control none --profile-from-start off -k regex:conv_dgrad -s 10 -c 6 --csv --page raw --log-file gpurun_out/sec_dgrad.csv python scripts/profile_step.py --region fwdbwd > gpurun_out/p3.log 2>&1
timeout 300 ncu $S --clock-control none --profile-from-start off -k regex:gn_bwd_fused -s 10 -c 6 --csv --page raw --log-file gpurun_out/sec_gnbwd.csv python scripts/profile_step.py --region fwdbwd > gpurun_out/p4.log 2>&1
ls -la gpurun_out/sec_*.csv
