mkdir -p gpurun_out
timeout 600 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>gpurun_out/bench_c3.err
tail -1 gpurun_out/bench_c3.log | cut -c1-330; tail -2 gpurun_out/bench_c3.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
tail -1 gpurun_out/bench_n2.log | cut -c1-330; tail -2 gpurun_out/bench_n2.err
