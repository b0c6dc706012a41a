# Round-end verification: full GPU test suite, smoke, default bench (+ CPU baseline), reference arm, C3 data point,
# ncu launch list of one frame and DRAM bytes of one forward.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
tail -2 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err
tail -1 gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err
tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>gpurun_out/bench_c3.err
tail -1 gpurun_out/bench_c3.log | cut -c1-400
timeout 300 python scripts/phase_times.py > gpurun_out/phase_times.log 2>&1
N="ncu --clock-control none --profile-from-start off --csv"
timeout 900 $N --metrics gpu__time_duration.sum --log-file gpurun_out/r01c_launches_frame.csv python scripts/profile_step.py --region frame --tc 3 > gpurun_out/p1.log 2>&1
timeout 600 $N --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --log-file gpurun_out/r01c_launches_forward_dram.csv python scripts/profile_step.py --region forward --tc 3 > gpurun_out/p2.log 2>&1
wc -l gpurun_out/r01c_*.csv
