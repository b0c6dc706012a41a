mkdir -p gpurun_out
N="ncu --set full --import-source on --clock-control none --cache-control none --profile-from-start off"
for k in gn_bwd_fused_kernel gn_fwd_fused_kernel conv_wgrad_kernel conv_dgrad_kernel conv_tf32x3_kernel; do
  timeout 600 $N -k regex:$k -s 20 -c 3 -o gpurun_out/src_$k -f python scripts/profile_step.py --region fwdbwd --tc 1 > gpurun_out/ncu_$k.log 2>&1
  ls -la gpurun_out/src_$k.ncu-rep
done
