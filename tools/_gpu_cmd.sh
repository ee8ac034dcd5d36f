python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_configs.py sir seir rmit 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], 'regs', d['regs'], 'ms %.2f'%d['kernel_ms'], 'steps/s %.4g'%d['steps_per_s'])
    else: print(l.rstrip())"
