mkdir -p gpurun_out
N="ncu --clock-control none --profile-from-start off --csv"
timeout 900 $N --metrics gpu__time_duration.sum --log-file gpurun_out/r01b_launches_frame.csv python scripts/profile_step.py --region frame --tc 3 > gpurun_out/p1.log 2>&1
timeout 600 $N --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --log-file gpurun_out/r01b_launches_forward_dram.csv python scripts/profile_step.py --region forward --tc 3 > gpurun_out/p2.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:conv_tf32x3 -s 10 -c 3 -o gpurun_out/full_tc -f python scripts/profile_step.py --region forward --tc 3 > gpurun_out/p3.log 2>&1
ncu -i gpurun_out/full_tc.ncu-rep --page details --csv > gpurun_out/r01b_full_conv_tf32x3_details.csv 2>/dev/null
rm -f gpurun_out/full_tc.ncu-rep
wc -l gpurun_out/r01b_*.csv
