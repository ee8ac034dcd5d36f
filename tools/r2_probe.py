"""Round-2 probe: BASELINE configs 2 / 4 / 5 at (or near) their stated sizes with device-resident
buffers, plus the integer-multiply microbenchmarks.  One JSON line per item.
usage: python tools/r2_probe.py [micro] [seir] [rmit] [tb] [--members-scale F]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from examples import models
from popdynamics_b200 import capi, compiler, ensemble


def emit(**kw):
    print(json.dumps(kw), flush=True)


def micro():
    names = {0: 'fp64 DMUL+DADD (ops)', 1: 'fp64 DFMA (ops)', 2: 'IMAD.WIDE.U32+LOP3 (3 ops/elem)',
             3: 'IMAD.HI.U32+LOP3 (elems)', 4: 'IMAD lo+LOP3 (elems)', 5: 'SHF+SHF+LOP3 (elems)'}
    lane_clk = 148 * 128 * 1.965e9 / 1e9
    for kind, name in names.items():
        g = capi.microbench(kind, 0)
        emit(item='micro', kind=kind, name=name, gops=g, per_lane_clk=g / lane_clk)


def explicit_sweep(name, model, times, M, reps=2):
    model.make_times(*times)
    model.check_flow_transfers()
    cm = compiler.compile_model(model, 'explicit')
    dev = torch.device('cuda:0')
    mp = torch.from_numpy(ensemble._member_param_block(cm, M)).to(dev)
    out = torch.zeros((cm.n_comp, M), dtype=torch.float64, device=dev)
    best = 1e30
    for _ in range(reps):
        res = ensemble.run_compiled(model, cm, 'explicit', model.get_init_list(),
                                    model.target_times, M, output='final', device=0,
                                    member_params=mp, out=out)
        best = min(best, res.info['kernel_ms'])
        err = res.info['err_code']
    steps = int(res.info['n_steps'])
    emit(item=name, members=M, n_time=len(model.target_times), steps=steps, kernel_ms=best,
         steps_per_s=steps / (best * 1e-3), comp_steps_per_s=steps * cm.n_comp / (best * 1e-3),
         regs=res.info['regs_per_thread'], err=err)
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    scale = 1.0
    if '--members-scale' in sys.argv:
        scale = float(sys.argv[sys.argv.index('--members-scale') + 1])
    want = set(args) or {'micro', 'seir', 'rmit', 'tb'}
    rng = np.random.default_rng(0)
    if 'micro' in want:
        micro()
    if 'seir' in want:
        M = int(10**6 * scale)
        d = dict(models.SEIR_PARAMS['measles'])
        d['r0'] = rng.uniform(2., 18., M)
        d['duration_infectious'] = rng.uniform(2., 10., M)
        d['duration_preinfectious'] = rng.uniform(2., 10., M)
        explicit_sweep('seir_demog_50y', models.SeirDemographyModel(d), (0, 365 * 50, 1), M)
    if 'rmit' in want:
        M = int(10**8 * scale)
        t0 = time.time()
        p = dict(models.RMIT_FIXED_PARAMS)
        p.update(beta=rng.uniform(1., 500., M), rho=rng.uniform(0., 10., M),
                 sigma1=rng.uniform(0., .25, M), sigma2=rng.uniform(0., .25, M),
                 sigma3=rng.uniform(0., .25, M))
        emit(item='rmit_host_generation_s', seconds=time.time() - t0)
        explicit_sweep('rmit_1e8', models.RmitTbModel(p), (0, 500., 1.), M, reps=1)
    if 'tb' in want:
        M = int(125000 * scale)
        model = models.ScaleupTbModel([])
        model.make_times(1950, 2000, 1.)
        model.check_flow_transfers()
        method = 'continuous_time_stochastic'
        cm = compiler.compile_model(model, method)
        dev = torch.device('cuda:0')
        out = torch.zeros((cm.n_comp, M), dtype=torch.float64, device=dev)
        res = ensemble.run_compiled(model, cm, method, model.get_init_list(), model.target_times,
                                    M, output='final', seed=0, device=0, count_bits=32, out=out)
        ev = int(res.info['n_steps'])
        emit(item='tb_cts_50y', members=M, events=ev, kernel_ms=res.info['kernel_ms'],
             events_per_s=ev / (res.info['kernel_ms'] * 1e-3), events_per_member=ev / M,
             regs=res.info['regs_per_thread'], blocks_per_sm=res.info['blocks_per_sm'],
             err=res.info['err_code'])


if __name__ == '__main__':
    main()
