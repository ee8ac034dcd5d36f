# -*- coding: utf-8 -*-
"""
Generates tests/golden/reference_vectors.npz by running the UNMODIFIED reference
(/root/reference/basepop.py + the model classes of /root/reference/examples/*.py, loaded through
oracle/reference.py) in this container.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden

The reference has no tests or golden data of its own, so these vectors are
the pins: explicit trajectories of every example model, seeded stochastic runs (the reference
draws from the global ``random`` / ``numpy.random`` MT19937 generators, which the oracle restates),
injected-stream stochastic runs, ``pick_event`` known answers, diagnostics arrays and an
``is_continue`` chain.  Environment: CPython 3.12 (Neumaier ``sum``), numpy 2.3.
"""

import os
import sys

import numpy as np

from . import reference

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, os.pardir, 'tests', 'golden', 'reference_vectors.npz')

SEIR = {
    'measles': {'population': 1e6, 'start_infectious': 1., 'r0': 13.,
                'duration_preinfectious': 8., 'duration_infectious': 7.,
                'life_expectancy': 70. * 365.},
    'flu': {'population': 1e6, 'start_infectious': 1., 'r0': 2., 'duration_preinfectious': 2.,
            'duration_infectious': 2., 'life_expectancy': 70. * 365.},
}
RMIT = {'mu': 1. / 70., 'd': .1, 'phi': 12, 'f': .05, 'eta': 2e-4, 'tau': .5, 'alpha': 2. / 9.,
        'theta': 1., 'rho': 0.1, 'omega': 2e-5, 'sigma1': 0., 'sigma2': 0., 'sigma3': 0.}


def injected_stream(seed, n):
    """The uniform stream of the injected-stream tests (regenerated at test time)."""
    return np.random.default_rng(seed).random(n)


def main():
    if not reference.available():
        sys.exit('reference tree not found: %s' % reference.REFERENCE_DIR)
    out = {}
    sir = reference.example_classes('sir_model.py')['SirModel']
    seir = reference.example_classes('seir_model.py')
    strains = reference.example_classes('stochastic_strains_model.py')['StrainsModel']
    tb = reference.example_classes('tb_model.py')['SimpleTbModel']
    rmit = reference.example_classes('rmit_tb_model.py')['RmitTbModel']
    mix = reference.example_classes('mix_model.py')['SirModel']

    # ---- explicit ------------------------------------------------------------------------
    m = sir(); m.make_times(0, 50, 1); m.integrate()
    out['sir_explicit'] = m.soln_array
    out['sir_explicit_var_array'] = m.var_array
    out['sir_explicit_var_labels'] = np.array(m.var_labels)
    out['sir_explicit_flow_array'] = m.flow_array
    out['sir_explicit_fraction_immune'] = np.array(m.fraction_soln['immune'])
    for name in ('measles', 'flu'):
        m = seir['SeirModel'](SEIR[name]); m.make_times(0, 200, 1); m.integrate('explicit')
        out['seir_%s_explicit' % name] = m.soln_array
    m = seir['SeirDemographyModel'](SEIR['measles']); m.make_times(0, 2000, 1)
    m.integrate('explicit')
    out['seird_explicit_2000'] = m.soln_array
    # BASELINE config 2 at its stated length: 50 years, 376 696 derivative evaluations
    # (examples/seir_model.py:458-464); rows every 365 days plus the final state
    m = seir['SeirDemographyModel'](SEIR['measles']); m.make_times(0, 365 * 50, 1)
    m.integrate('explicit')
    out['seird_explicit_18250_every365'] = m.soln_array[::365]
    out['seird_explicit_18250_final'] = m.soln_array[-1]
    for tag, interventions in (('tb', []), ('tb_vacc', ['vaccination'])):
        m = tb(interventions); m.make_times(1900, 2050, .05); m.integrate(method='explicit')
        out[tag + '_explicit_every50'] = m.soln_array[::50]
        out[tag + '_explicit_final'] = m.soln_array[-1]
        out[tag + '_var_labels'] = np.array(m.var_labels)
        out[tag + '_var_array_every50'] = m.var_array[::50]
    m = rmit(RMIT); m.set_param('beta', 50.); m.make_times(0, 500., 1.)
    m.integrate(method='explicit')
    out['rmit_explicit'] = m.soln_array
    out['rmit_proportion'] = np.array([m.vars['proportion']])
    betas = [1., 7., 20., 99., 250., 500.]
    props, vals = [], []
    for beta in betas:
        m = rmit(dict(RMIT, sigma1=0.25, sigma2=0.125, sigma3=0.125, tau=2.))
        m.set_param('beta', beta); m.set_param('rho', 0.5); m.make_times(0, 500., 1.)
        m.integrate(method='explicit')
        props.append(m.soln_array[-1])
        vals.append(m.vars['proportion'])
    out['rmit_sweep_betas'] = np.array(betas)
    out['rmit_sweep_final'] = np.array(props)
    out['rmit_sweep_proportion'] = np.array(vals)   # model.vars['proportion'], rmit_tb_model.py:143
    # figure 2 of the RMIT example (rmit_tb_model.py:133-143): sigma = 0, beta swept, the value
    # read is model.vars['proportion'] of the final state -- six of its beta points
    finals, vals = [], []
    for beta in betas:
        m = rmit(RMIT); m.set_param('beta', float(beta)); m.make_times(0, 500., 1.)
        m.integrate(method='explicit')
        finals.append(m.soln_array[-1]); vals.append(m.vars['proportion'])
    out['rmit_fig2_betas'] = np.array(betas)
    out['rmit_fig2_final'] = np.array(finals)
    out['rmit_fig2_proportion'] = np.array(vals)
    m = strains(); m.make_times(0, 200, 1); m.integrate('explicit')
    out['strains_explicit'] = m.soln_array

    # ---- integrate('scipy'): scipy.integrate.odeint (LSODA) on the derivative closure --------
    # (basepop.py:882-886).  Not reproducible bit for bit by another solver: these pin the
    # adaptive device integrator within the tolerance stated in DESIGN.md.
    import scipy.integrate  # noqa: F401  (the reference checks sys.modules for 'scipy')
    m = sir(); m.make_times(0, 50, 1); m.integrate('scipy')
    out['sir_odeint'] = m.soln_array
    m = seir['SeirModel'](SEIR['measles']); m.make_times(0, 200, 1); m.integrate('scipy')
    out['seir_measles_odeint'] = m.soln_array
    # SEIR with demography: only the first 300 days.  Over longer horizons (2000 days, the 50
    # years of config 2) LSODA evaluates the derivative on a slightly negative internal state
    # between epidemic waves and the reference's own checks() assert (basepop.py:623 -> :1025)
    # aborts integrate('scipy') with AssertionError.
    m = seir['SeirDemographyModel'](SEIR['measles']); m.make_times(0, 300, 1)
    m.integrate('scipy')
    out['seird_odeint_300'] = m.soln_array
    m = tb([]); m.make_times(1900, 2050, .05); m.integrate(method='scipy')
    out['tb_odeint_every50'] = m.soln_array[::50]
    m = rmit(RMIT); m.set_param('beta', 50.); m.make_times(0, 100., 1.); m.integrate('scipy')
    out['rmit_odeint_100'] = m.soln_array
    m = strains(); m.make_times(0, 200, 1); m.integrate('scipy')
    out['strains_odeint'] = m.soln_array

    # ---- is_continue chain (numpy.float64 state -> naive sum path) -----------------------
    m = mix({'start_population': 100}); m.make_times(0, 5, 0.2); m.integrate('explicit')
    m.make_times(5, 10, 0.2); m.integrate('explicit', is_continue=True)
    out['mix_continue_explicit'] = m.soln_array
    out['mix_continue_times'] = np.array(m.times, dtype=float)

    # ---- seeded stochastic runs (global MT19937 state) -----------------------------------
    for seed in (1, 2, 7):
        for method, tag in (('discrete_time_stochastic', 'dts'),
                            ('continuous_time_stochastic', 'cts')):
            reference.seed_all(seed)
            m = strains(); m.make_times(0, 300, 1); m.integrate(method)
            out['strains_%s_seed%d' % (tag, seed)] = m.soln_array
            reference.seed_all(seed)
            m = sir(); m.make_times(0, 50, 1); m.integrate(method)
            out['sir_%s_seed%d' % (tag, seed)] = m.soln_array
    # a population large enough to reach the PTRS branch (lam >= 10)
    reference.seed_all(11)
    m = sir({'population': 5000, 'start_infectious': 20.}); m.make_times(0, 60, 1)
    m.integrate('discrete_time_stochastic')
    out['sir_big_dts_seed11'] = m.soln_array
    reference.seed_all(5)
    m = seir['SeirDemographyModel'](dict(SEIR['measles'], population=20000., start_infectious=10.))
    m.make_times(0, 120, 1); m.integrate('discrete_time_stochastic')
    out['seird_dts_seed5'] = m.soln_array

    # ---- scale-up functions inside the stochastic loops (declared extension):
    # the reference class with calculate_vars calling calculate_scaleup_vars() first
    class ScaleupTb(tb):
        def calculate_vars(self):
            self.calculate_scaleup_vars()
            tb.calculate_vars(self)
    for method, tag, times in (('discrete_time_stochastic', 'dts', (1985, 2000, 1.)),
                               ('continuous_time_stochastic', 'cts', (1988, 1992, 1.))):
        reference.seed_all(9)
        m = ScaleupTb([]); m.make_times(*times); m.integrate(method)
        out['tb_scaleup_%s_seed9' % tag] = m.soln_array

    # ---- injected-stream runs ------------------------------------------------------------
    for k in range(4):
        u = injected_stream(100 + k, 20000)
        with reference.injected_uniforms(u) as inj:
            m = strains(); m.make_times(0, 150, 1); m.integrate('discrete_time_stochastic')
        out['strains_dts_inject%d' % k] = m.soln_array
        out['strains_dts_inject%d_consumed' % k] = np.array([inj.rng.consumed])
        u = injected_stream(200 + k, 20000)
        with reference.injected_uniforms(u) as inj:
            m = strains(); m.make_times(0, 150, 1); m.integrate('continuous_time_stochastic')
        out['strains_cts_inject%d' % k] = m.soln_array
        out['strains_cts_inject%d_consumed' % k] = np.array([inj.rng.consumed])

    # ---- pick_event known answers --------------------------------------------------------
    ref = reference.basepop()
    rng = np.random.default_rng(42)
    intervals, draws, picks = [], [], []
    saved = ref.random.random
    try:
        for _ in range(200):
            n = int(rng.integers(1, 12))
            iv = rng.random(n) * (rng.random(n) > 0.25)
            u = float(rng.random())
            ref.random.random = lambda u=u: u
            picks.append(ref.pick_event(list(iv)))
            row = np.full(12, -1.0); row[:n] = iv
            intervals.append(row); draws.append(u)
    finally:
        ref.random.random = saved
    out['pick_intervals'] = np.array(intervals)
    out['pick_uniform'] = np.array(draws)
    out['pick_event'] = np.array(picks)

    np.savez_compressed(OUT, **out)
    print('wrote %s: %d arrays, %.1f KiB' % (os.path.normpath(OUT), len(out),
                                             os.path.getsize(OUT) / 1024.))


if __name__ == '__main__':
    main()
