#!/bin/bash
cd "$GRAFT_REPO_ROOT"
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline "$@" 2>&1 | grep -v "^\[bench\]" | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['config'].get('dynamic_steps'))"; }
echo "buckets off"; DBOA_DP_BUCKETS=0 run
for ch in 2 4 8 16; do echo "buckets on, NCCL_MAX_NCHANNELS=$ch"; NCCL_MAX_NCHANNELS=$ch run; done
echo "buckets off, NCCL_MAX_NCHANNELS=8"; DBOA_DP_BUCKETS=0 NCCL_MAX_NCHANNELS=8 run
echo "c5 thr 2e-5"; run --workload c5 --cos-threshold 2e-5
