mkdir -p gpurun_out
DBOA_PDL=0 timeout 300 python scripts/conv_microbench.py > gpurun_out/conv_mb_pdl0.log 2>&1
DBOA_PDL=1 timeout 300 python scripts/conv_microbench.py > gpurun_out/conv_mb_pdl1.log 2>&1
cat gpurun_out/conv_mb_pdl0.log
cat gpurun_out/conv_mb_pdl1.log
