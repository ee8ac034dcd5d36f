"""Time pd_gillespie_kernel on the TB model (config 4) over launch-geometry variants.
usage: python tools/tune_gillespie.py [members] [years] [block:min_blocks,...]
POPDYN_DEFINE=-DPD_GIL_OPT=<bits> selects the per-event arithmetic forms (kernels/pd_gillespie.cuh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from examples import configs
from popdynamics_b200 import compiler, ensemble


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
    years = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    model = configs.tb_gillespie()
    model.make_times(1950, 1950 + years, years)
    model.check_flow_transfers()
    method = 'continuous_time_stochastic'
    cm = compiler.compile_model(model, method)
    out = torch.zeros((cm.n_comp, M), dtype=torch.float64, device='cuda')
    geoms = ((256, 0), (256, 3), (256, 4), (128, 0), (128, 5), (128, 6), (128, 8), (64, 10), (64, 16))
    if len(sys.argv) > 3:
        geoms = tuple(tuple(int(x) for x in g.split(':')) for g in sys.argv[3].split(','))
    for block, minb in geoms:
        best = 1e30
        for _ in range(2):
            res = ensemble.run_compiled(model, cm, method, model.get_init_list(),
                                        model.target_times, M, output='final', seed=0, device=0,
                                        count_bits=32, out=out, block=block,
                                        min_blocks=minb if minb else 1)
            best = min(best, res.info['kernel_ms'])
        ev = int(res.info['n_steps'])
        print('define=%s block=%d minb=%d regs=%d blocks/sm=%d  %.1f ms  %.3e events/s'
              % (os.environ.get('POPDYN_DEFINE', '-'), block, minb, res.info['regs_per_thread'],
                 res.info['blocks_per_sm'], best, ev / (best * 1e-3)), flush=True)


if __name__ == '__main__':
    main()
