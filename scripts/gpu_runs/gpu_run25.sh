mkdir -p gpurun_out
for m in 2 4 8; do
  DBOA_TC_MINKB=$m timeout 300 python scripts/conv_microbench.py > gpurun_out/conv_mb_minkb$m.log 2>&1
  echo "minkb $m"; tail -1 gpurun_out/conv_mb_minkb$m.log
  DBOA_TC_MINKB=$m timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -o '"value": [0-9.]*' | head -1
done
paste <(awk '{print $1,$2,$3,$4,$5,$8,$11}' gpurun_out/conv_mb_minkb2.log) <(awk '{print $8,$11}' gpurun_out/conv_mb_minkb4.log) <(awk '{print $8,$11}' gpurun_out/conv_mb_minkb8.log) | head -30
