#!/bin/bash
# N x B200 (gpurun --gpus N): NCCL path of the data-parallel parity test, then C2 / C5 bench lines under torchrun next to N = 1
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/dp_nccl_test.log
for wl in ${WORKLOADS:-c2 c5}; do          # WORKLOADS=c2 NO_N1=1: the short form (one torchrun line)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline --workload $wl 2>&1 | grep -v "^\[bench\]" | tail -1 > gpurun_out/bench_n${N}_$wl.json
  [ -z "$NO_N1" ] && timeout 900 python bench.py --no-cpu-baseline --workload $wl 2>&1 | tail -1 > gpurun_out/bench_n1_$wl.json
done
python - <<PY
import json
for f in ('n1_c2', 'n${N}_c2', 'n1_c5', 'n${N}_c5'):
    try:
        d = json.loads(open(f'gpurun_out/bench_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'], 2), 'frames/s', round(d['ms_per_step'], 3), 'ms', 'e2e', round(d['e2e']['value'], 2), 'dynamic steps', d['config'].get('dynamic_steps'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
