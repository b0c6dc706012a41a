mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_fwd_kernel|gn_bwd_fused|conv_dgrad" -c 120 -o gpurun_out/prof_k python scripts/profile_step.py --region fwdbwd > gpurun_out/prof_k.log 2>&1
ls -la gpurun_out/*.ncu-rep
