"""Run ONE tau-leap / Gillespie kernel variant on cuda:0 (for ncu captures and quick timing).
usage: python tools/run_one.py METHOD OUT BITS BLOCK MINB M T [reps]
  METHOD: dts | cts      OUT: stats | stats_nohist | final | full"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from examples import models
from popdynamics_b200 import compiler, ensemble


def main():
    method, out, bits, block, minb, M, T = sys.argv[1:8]
    reps = int(sys.argv[8]) if len(sys.argv) > 8 else 3
    bits, block, minb, M, T = int(bits), int(block), int(minb), int(M), int(T)
    method = {'dts': 'discrete_time_stochastic', 'cts': 'continuous_time_stochastic'}[method]
    model = models.StrainsModel()
    model.make_times(0, T, 1)
    model.check_flow_transfers()
    cm = compiler.compile_model(model, method)
    y = model.get_init_list()
    dev = torch.device('cuda:0')
    C = cm.n_comp
    kw = dict(output='stats' if out.startswith('stats') else out, seed=1, device=0,
              count_bits=bits, block=block, min_blocks=minb)
    if out == 'stats_histmoments':
        kw.update(output='stats', hist_bins=2048, hist_max=2047, moments_from='hist',
                  moments=torch.zeros((T + 1, C, 2), dtype=torch.int64, device=dev),
                  hist=torch.zeros((T + 1, C, 2048), dtype=torch.int32, device=dev))
    elif out == 'stats':
        kw.update(hist_bins=2048, hist_max=2047,
                  moments=torch.zeros((T + 1, C, 2), dtype=torch.int64, device=dev),
                  hist=torch.zeros((T + 1, C, 2048), dtype=torch.int32, device=dev))
    elif out == 'stats_nohist':
        kw.update(moments=torch.zeros((T + 1, C, 2), dtype=torch.int64, device=dev))
    elif out == 'final':
        kw.update(out=torch.zeros((C, M), dtype=torch.float64, device=dev))
    else:
        kw.update(out=torch.zeros((T + 1, C, M), dtype=torch.float64, device=dev))
    best = 1e30
    for rep in range(reps):
        res = ensemble.run_compiled(model, cm, method, y, model.target_times, M, **kw)
        best = min(best, res.info['kernel_ms'])
    steps = res.info.get('n_steps', M * T)
    print('%s %s bits=%d block=%d minb=%d regs=%d blocks/sm=%d smem=%d  %.2f ms  %.3e steps/s'
          % (method, out, bits, block, minb, res.info['regs_per_thread'],
             res.info['blocks_per_sm'], res.info.get('smem_bytes', -1), best,
             steps / (best * 1e-3)), flush=True)


if __name__ == '__main__':
    main()
