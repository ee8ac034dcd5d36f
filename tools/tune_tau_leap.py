"""Kernel-variant sweep for the tau-leap kernel on the GPU box (block size, resident-block cap,
count width, output mode).  Prints one line per variant: kernel ms and member-steps/s."""
import itertools
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from examples import models
from popdynamics_b200 import compiler, ensemble


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    model = models.StrainsModel()
    model.make_times(0, T, 1)
    model.check_flow_transfers()
    cm = compiler.compile_model(model, 'discrete_time_stochastic')
    y = model.get_init_list()
    dev = torch.device('cuda:0')
    C = 5
    moments = torch.zeros((T + 1, C, 2), dtype=torch.int64, device=dev)
    hist = torch.zeros((T + 1, C, 2048), dtype=torch.int32, device=dev)
    final = torch.zeros((C, M), dtype=torch.float64, device=dev)
    variants = list(itertools.product(('stats', 'stats_nohist', 'final'), (32, 64), (128, 256, 512), (1, 2, 3, 4)))
    if os.environ.get('PD_TUNE_QUICK'):
        variants = [v for v in variants if v[1] == 32]
    for out, bits, block, mb in variants:
        kw = dict(output='stats' if out.startswith('stats') else 'final', seed=1, device=0,
                  count_bits=bits, block=block, min_blocks=max(1, mb * 256 // block))
        if out == 'stats':
            kw.update(hist_bins=2048, hist_max=2047, moments=moments, hist=hist)
        elif out == 'stats_nohist':
            kw.update(moments=moments)
        else:
            kw.update(out=final)
        best = 1e30
        try:
            for rep in range(3):
                res = ensemble.run_compiled(model, cm, 'discrete_time_stochastic', y,
                                            model.target_times, M, **kw)
                best = min(best, res.info['kernel_ms'])
        except Exception as exc:
            print('%-12s bits=%d block=%3d minb=%d FAILED: %s' % (out, bits, block, kw['min_blocks'], str(exc)[:120]), flush=True)
            continue
        print('%-12s bits=%d block=%3d minb=%d regs=%3d blocks/sm=%d  %8.2f ms  %.3e member-steps/s'
              % (out, bits, block, kw['min_blocks'], res.info['regs_per_thread'],
                 res.info['blocks_per_sm'], best, M * T / (best * 1e-3)), flush=True)


if __name__ == '__main__':
    main()
