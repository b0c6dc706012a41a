mkdir -p gpurun_out
R="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$R --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --serial-output 2>/dev/null | tail -1 | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -4
$R --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null | tail -1 | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -4
NCCL_DEBUG=WARN $R --master-port 29543 bench.py --gpus 2 --steps 40 --warmup 5 2>/dev/null | tail -1 | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -4
