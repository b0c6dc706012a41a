#!/bin/bash
# round 2, 2 x B200: NCCL path of the data-parallel test, then bench C2 / C5 under torchrun
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_dp2_nccl_test.log
for wl in c2 c5; do
  extra=""; [ $wl = c5 ] && extra="--cos-threshold 2e-6"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --workload $wl $extra 2>&1 | grep -v "^\[bench\]" | tail -2 | tee gpurun_out/r2_bench_n2_$wl.json
done
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_bench_n1_c2.json
timeout 900 python bench.py --no-cpu-baseline --workload c5 --cos-threshold 2e-6 2>&1 | tail -1 > gpurun_out/r2_bench_n1_c5.json
python - <<'PY'
import json
for f in ('n1_c2','n2_c2','n1_c5','n2_c5'):
    try:
        d=json.loads(open(f'gpurun_out/r2_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms', 'e2e', round(d['e2e']['value'],2), d['config'].get('dynamic_steps'), d['config'].get('one_minus_cos12_first_test'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
