python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 3000 gpurun_out/bench_r1.json; tail -3 gpurun_out/bench_r1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1.json 2>&1; tail -c 1500 gpurun_out/bench_ref_r1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 1 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log | cut -c1-300
ncu --set full --import-source on --clock-control none -k pd_tau_leap_kernel -c 1 -o gpurun_out/tau_stats_bench_r1 python bench.py --steps 1 --warmup 0 --cpu-seconds 1 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-300
