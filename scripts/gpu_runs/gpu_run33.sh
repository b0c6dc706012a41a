mkdir -p gpurun_out
R="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for i in 1 2 3; do
$R --master-port 2955$i bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null | tail -1 | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -4 | tr '\n' ' '; echo
done
