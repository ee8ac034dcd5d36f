"""Run ONE explicit-Euler or adaptive ('scipy') sweep launch on cuda:0 (for ncu captures).
usage: python tools/run_explicit_like.py METHOD CONFIG MEMBERS T_END   (CONFIG: seir | rmit)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from examples import configs
from popdynamics_b200 import compiler, ensemble


def main():
    method, config, M, t_end = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
    model = (configs.seir_sweep if config == 'seir' else configs.rmit_sweep)(0, M)
    model.make_times(0, int(t_end) if config == 'seir' else t_end, 1 if config == 'seir' else 1.)
    model.check_flow_transfers()
    cm = compiler.compile_model(model, method)
    rows = ensemble._member_param_rows(cm, M)
    mp = torch.stack([torch.from_numpy(r) for r in rows]).cuda()
    out = torch.empty((cm.n_comp, M), dtype=torch.float64, device='cuda')
    block = int(os.environ.get('PD_RUN_BLOCK', 256))
    minb = int(os.environ.get('PD_RUN_MIN_BLOCKS', 0))
    for _ in range(2):
        res = ensemble.run_compiled(model, cm, method, model.get_init_list(), model.target_times,
                                    M, output='final', device=0, member_params=mp, out=out,
                                    block=block, min_blocks=minb)
    steps = int(res.info['n_steps'])
    print('%s %s M=%d block=%d minb=%d regs=%d blocks/sm=%d  %.2f ms  %.3e steps/s'
          % (method, config, M, block, minb, res.info['regs_per_thread'],
             res.info['blocks_per_sm'], res.info['kernel_ms'],
             steps / (res.info['kernel_ms'] * 1e-3)))


if __name__ == '__main__':
    main()
