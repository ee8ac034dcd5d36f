"""Throughput of the BASELINE.json configurations other than the headline one (which bench.py
measures): kernel time from CUDA events (inputs resident in HBM), comp-steps/s, and the oracle
port on the host cores beside it.  Prints one JSON line per configuration.

    python tools/bench_configs.py [config ...]     configs: sir seir rmit strains_cts strains_dts_full
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from examples import models
from oracle import oracle
from popdynamics_b200 import compiler, ensemble


def seir_sweep(n, rng):
    d = dict(models.SEIR_PARAMS['measles'])
    d['r0'] = rng.uniform(2., 18., n)
    d['duration_infectious'] = rng.uniform(2., 10., n)
    d['duration_preinfectious'] = rng.uniform(2., 10., n)
    return models.SeirDemographyModel(d)


def rmit_sweep(n, rng):
    p = dict(models.RMIT_FIXED_PARAMS)
    # sigma in [0, .25] as in the reference's own figures (rmit_tb_model.py:150-178); with
    # sigma up to .5 and beta up to 500 explicit Euler at dt = .05 diverges for ~7 % of the
    # members and checks() raises, in the reference as here
    p.update(beta=rng.uniform(1., 500., n), rho=rng.uniform(0., 10., n),
             sigma1=rng.uniform(0., .25, n), sigma2=rng.uniform(0., .25, n),
             sigma3=rng.uniform(0., .25, n))
    return models.RmitTbModel(p)


CONFIGS = {
    # name: (builder(n, rng), method, make_times args, members, output, cpu sample members)
    'sir': (lambda n, rng: models.SirModel({}), 'explicit', (0, 50, 1), 1 << 22, 'final', 20000),
    'seir': (seir_sweep, 'explicit', (0, 3650, 1), 1 << 20, 'final', 256),
    'rmit': (rmit_sweep, 'explicit', (0, 500., 1.), 1 << 22, 'final', 2000),
    'strains_cts': (lambda n, rng: models.StrainsModel(), 'continuous_time_stochastic',
                    (0, 1000, 1), 1 << 20, 'final', 2000),
    'strains_cts_stats': (lambda n, rng: models.StrainsModel(), 'continuous_time_stochastic',
                          (0, 1000, 1), 1 << 20, 'stats', 2000),
    'strains_cts_stats_hist': (lambda n, rng: models.StrainsModel(), 'continuous_time_stochastic',
                               (0, 1000, 1), 1 << 20, 'stats_hist', 2000),
    'strains_dts_full': (lambda n, rng: models.StrainsModel(), 'discrete_time_stochastic',
                         (0, 200, 1), 1 << 20, 'full', 2000),
    # BASELINE config 4: TB model with its scale-ups evaluated inside the stochastic loops
    # (population 1e6: ~6e4 events per member-year; every tau-leap mean is in the PTRS branch)
    'tb_cts': (lambda n, rng: models.ScaleupTbModel([]), 'continuous_time_stochastic',
               (1950, 1951, 1.), 1 << 16, 'final', 64),
    'tb_dts': (lambda n, rng: models.ScaleupTbModel([]), 'discrete_time_stochastic',
               (1950, 2000, 1.), 1 << 20, 'final', 20000),
}


def run(name, reps=3):
    build, method, times, M, output, cpu_n = CONFIGS[name]
    M = int(os.environ.get('PD_BENCH_MEMBERS', M))
    rng = np.random.default_rng(0)
    model = build(M, rng)
    model.make_times(*times)
    model.check_flow_transfers()
    cm = compiler.compile_model(model, method)
    y = model.get_init_list()
    C, T = cm.n_comp, len(model.target_times)
    dev = torch.device('cuda:0')
    moments_from = 'hist' if output == 'stats_hist' else 'kernel'
    output = output.split('_')[0]
    kw = dict(output=output, seed=1, device=0)
    if output == 'stats':
        kw['moments_from'] = moments_from
    if method != 'explicit':
        kw.update(count_bits=32, block=int(os.environ.get('PD_BENCH_BLOCK', 256)))
        if os.environ.get('PD_BENCH_MIN_BLOCKS'):
            kw['min_blocks'] = int(os.environ['PD_BENCH_MIN_BLOCKS'])
    if output == 'final':
        kw['out'] = torch.zeros((C, M), dtype=torch.float64, device=dev)
    elif output == 'full':
        kw['out'] = torch.zeros((T, C, M), dtype=torch.float64, device=dev)
    else:
        kw.update(moments=torch.zeros((T, C, 2), dtype=torch.int64, device=dev),
                  hist=torch.zeros((T, C, 2048), dtype=torch.int32, device=dev),
                  hist_bins=2048, hist_max=2047)
    mp = ensemble._member_param_block(cm, M)
    if mp is not None:
        kw['member_params'] = torch.from_numpy(mp).to(dev)
    best, res = 1e30, None
    for _ in range(reps):
        res = ensemble.run_compiled(model, cm, method, y, model.target_times, M, **kw)
        ensemble.raise_device_error(res.info, method)
        best = min(best, res.info['kernel_ms'])
    steps = int(res.info['n_steps'])        # member-steps (explicit, DTS) or events (CTS)
    line = {'config': name, 'method': method, 'members': M, 'n_time': T, 'compartments': C,
            'output': output + ('/moments from hist' if moments_from == 'hist' else ''),
            'steps': steps, 'kernel_ms': best,
            'steps_per_s': steps / (best * 1e-3), 'comp_steps_per_s': steps * C / (best * 1e-3),
            'regs': res.info['regs_per_thread'], 'blocks_per_sm': res.info['blocks_per_sm'],
            'block': res.info['block']}
    # CPU: the oracle port on all host cores, bounded sample of the same workload
    model_c = build(cpu_n, np.random.default_rng(0))
    model_c.make_times(*times)
    model_c.check_flow_transfers()
    cm_c = compiler.compile_model(model_c, method)
    om = oracle.OracleModel(cm_c)
    n_thread = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    if method == 'explicit':
        out = om.explicit_batch(cpu_n, model_c.get_init_list(), model_c.target_times,
                                want_soln=False, n_thread=n_thread)
        cpu_steps = out['n_steps'] if 'n_steps' in out else steps / M * cpu_n
    else:
        out = om.stochastic_batch(method, cpu_n, model_c.get_init_list(), model_c.target_times,
                                  seed=1, want_soln=False, want_moments=False, n_thread=n_thread)
        cpu_steps = out['n_steps']
    dt = time.perf_counter() - t0
    line['cpu'] = {'steps_per_s': cpu_steps / dt, 'comp_steps_per_s': cpu_steps * C / dt,
                   'cores': n_thread, 'members': cpu_n, 'seconds': dt, 'kind': 'oracle port'}
    line['gpu_over_cpu'] = line['steps_per_s'] / line['cpu']['steps_per_s']
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    names = sys.argv[1:] or list(CONFIGS)
    for name in names:
        run(name)
