#!/bin/bash
# round 2 evidence: ncu launch list of one frame, DRAM bytes of one forward (no cache flush between kernels), full captures of the
# fused forward convolution (two layer classes), of the tcgen05 weight gradient and of the fused data gradient
cd "$GRAFT_REPO_ROOT"
T=${TAG:-r02e}      # file tag of this capture set (profiles/<tag>_*)
N="ncu --clock-control none --profile-from-start off"
timeout 900 $N --csv --metrics gpu__time_duration.sum --log-file gpurun_out/${T}_launches_frame_c2.csv python scripts/profile_step.py --region frame --tc 3 > gpurun_out/p1.log 2>&1
timeout 600 $N --csv --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --log-file gpurun_out/${T}_launches_forward_b1_dram.csv python scripts/profile_step.py --region forward --tc 3 > gpurun_out/p2.log 2>&1
timeout 600 $N --set full --import-source on -k regex:conv_wide_kernel --launch-skip 4 --launch-count 2 -o gpurun_out/${T}_full_conv_wide_layer1 python scripts/profile_step.py --region forward --tc 3 > gpurun_out/p3.log 2>&1
timeout 600 $N --set full --import-source on -k regex:conv_wide_kernel --launch-skip 40 --launch-count 2 -o gpurun_out/${T}_full_conv_wide_layer4 python scripts/profile_step.py --region forward --tc 3 > gpurun_out/p4.log 2>&1
[ -n "$WGRAD" ] && timeout 600 $N --set full --import-source on -k regex:conv_wgrad_wide_kernel --launch-skip 3 --launch-count 2 -o gpurun_out/${T}_full_wgrad_wide python scripts/profile_step.py --region fwdbwd --tc 3 > gpurun_out/p5.log 2>&1
timeout 600 $N --set full --import-source on -k regex:dgrad_wide_kernel --launch-skip 6 --launch-count 2 -o gpurun_out/${T}_full_dgrad_wide python scripts/profile_step.py --region fwdbwd --tc 3 > gpurun_out/p6.log 2>&1
ls -la gpurun_out/${T}_*; tail -2 gpurun_out/p3.log
timeout 300 python scripts/phase_times.py 2>&1 | tail -12 | tee gpurun_out/${T}_phase_times.txt
