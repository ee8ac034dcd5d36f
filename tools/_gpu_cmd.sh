python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1.json')); print('value %.4g e2e %.4g h2d %d d2h %d frac %.3f' % (d['value'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], d['roofline']['frac']))"; tail -3 gpurun_out/bench_r1.err
