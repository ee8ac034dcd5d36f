python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for cfg in "final 256 4" "final 512 2" "stats 256 4" "stats_nohist 256 4" "full 256 4"; do set -- $cfg; python tools/run_one.py dts $1 32 $2 $3 1048576 200 2>&1 | tail -1; done > gpurun_out/tune9.log 2>&1
python tools/run_one.py cts stats 32 256 2 262144 200 2>&1 | tail -1 >> gpurun_out/tune9.log
python tools/run_one.py cts final 32 256 2 262144 200 2>&1 | tail -1 >> gpurun_out/tune9.log
cat gpurun_out/tune9.log
ncu --set full --import-source on --clock-control none -k pd_tau_leap_kernel -c 1 -o gpurun_out/tau_final_e python tools/run_one.py dts final 32 256 4 1048576 200 1 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
