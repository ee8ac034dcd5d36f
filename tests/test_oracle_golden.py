"""Pins the CPU oracle (oracle/popdyn_oracle.c + the flow compiler's bytecode) to golden vectors
produced by the UNMODIFIED reference (oracle/make_golden.py).  Everything is bit-exact:
explicit trajectories, seeded stochastic runs reproduced through the MT19937 restatements,
injected-stream runs, pick_event, and the is_continue chain."""
import numpy as np
import pytest

from oracle import make_golden, oracle

import helpers


def assert_bitwise(a, b):
    assert a.shape == b.shape
    assert np.array_equal(a, b), 'max abs diff %g' % np.abs(a - b).max()


@pytest.mark.parametrize('name,key,times', [
    ('sir', 'sir_explicit', (0, 50, 1)),
    ('seir_measles', 'seir_measles_explicit', (0, 200, 1)),
    ('seir_flu', 'seir_flu_explicit', (0, 200, 1)),
    ('seird', 'seird_explicit_2000', (0, 2000, 1)),
    ('rmit', 'rmit_explicit', (0, 500., 1.)),
    ('strains', 'strains_explicit', (0, 200, 1)),
])
def test_explicit_matches_reference(golden, name, key, times):
    m = helpers.make_model(name)
    m.make_times(*times)
    soln, n_steps, err = helpers.oracle_explicit(m)
    assert err == 0
    assert_bitwise(soln, golden[key])


@pytest.mark.parametrize('name', ['tb', 'tb_vacc'])
def test_explicit_tb_with_scaleups(golden, name):
    m = helpers.make_model(name)
    m.make_times(1900, 2050, .05)
    assert m.target_times[-1] == 2049.9999999998727      # accumulating grid, basepop.py:373-376
    soln, n_steps, err = helpers.oracle_explicit(m)
    assert (n_steps, err) == (3000, 0)
    assert_bitwise(soln[::50], golden[name + '_explicit_every50'])
    assert_bitwise(soln[-1], golden[name + '_explicit_final'])


def test_explicit_step_counts_of_the_survey():
    # fp64 time accumulation gives 1024 / 4102 derivative calls, not 20*(T-1)
    for times, want in (((0, 50, 1), 1024), ((0, 200, 1), 4102)):
        m = helpers.make_model('sir')
        m.make_times(*times)
        t_eval, dt, interval = oracle.explicit_schedule(m.target_times)
        assert len(dt) == want
        assert dt.min() < 1e-12 and dt.max() == 0.05


def test_rmit_sweep(golden):
    for beta, want in zip(golden['rmit_sweep_betas'], golden['rmit_sweep_final']):
        m = helpers.make_model('rmit', params=dict(sigma1=0.25, sigma2=0.125, sigma3=0.125,
                                                   tau=2., beta=float(beta), rho=0.5))
        m.make_times(0, 500., 1.)
        soln, _, err = helpers.oracle_explicit(m)
        assert err == 0
        assert_bitwise(soln[-1], want)


def test_is_continue_chain_uses_naive_sum(golden):
    # after is_continue the state is numpy.float64 and Python's sum() is a plain left fold
    m = helpers.make_model('mix', params={'start_population': 100})
    m.make_times(0, 5, 0.2)
    first, _, _ = helpers.oracle_explicit(m)
    cm = helpers.compiled(m, 'explicit', is_continue=True)
    assert cm.naive_sum
    m.make_times(5, 10, 0.2)
    second, _, _ = oracle.OracleModel(cm).explicit(first[-1], m.target_times)
    assert_bitwise(np.concatenate((first, second[1:])), golden['mix_continue_explicit'])


@pytest.mark.parametrize('seed', [1, 2, 7])
@pytest.mark.parametrize('model,times', [('strains', (0, 300, 1)), ('sir', (0, 50, 1))])
def test_seeded_discrete_time_stochastic(golden, model, times, seed):
    m = helpers.make_model(model)
    m.make_times(*times)
    soln, n, err = helpers.oracle_stochastic(m, 'discrete_time_stochastic',
                                             oracle.Rng('numpy', seed))
    assert err == 0 and n == len(m.target_times) - 1
    assert_bitwise(soln, golden['%s_dts_seed%d' % (model, seed)])


@pytest.mark.parametrize('seed', [1, 2, 7])
@pytest.mark.parametrize('model,times', [('strains', (0, 300, 1)), ('sir', (0, 50, 1))])
def test_seeded_continuous_time_stochastic(golden, model, times, seed):
    m = helpers.make_model(model)
    m.make_times(*times)
    soln, n, err = helpers.oracle_stochastic(m, 'continuous_time_stochastic',
                                             oracle.Rng('python', seed))
    assert err == 0
    assert_bitwise(soln, golden['%s_cts_seed%d' % (model, seed)])


def test_seeded_runs_reaching_the_ptrs_branch(golden):
    m = helpers.make_model('sir', params={'population': 5000, 'start_infectious': 20.})
    m.make_times(0, 60, 1)
    soln, _, err = helpers.oracle_stochastic(m, 'discrete_time_stochastic', oracle.Rng('numpy', 11))
    assert err == 0
    assert_bitwise(soln, golden['sir_big_dts_seed11'])
    params = dict(helpers.models.SEIR_PARAMS['measles'], population=20000., start_infectious=10.)
    m = helpers.make_model('seird', params=params)
    m.make_times(0, 120, 1)
    soln, _, err = helpers.oracle_stochastic(m, 'discrete_time_stochastic', oracle.Rng('numpy', 5))
    assert err == 0
    assert_bitwise(soln, golden['seird_dts_seed5'])


@pytest.mark.parametrize('method,tag,kind,times', [
    ('discrete_time_stochastic', 'dts', 'numpy', (1985, 2000, 1.)),
    ('continuous_time_stochastic', 'cts', 'python', (1988, 1992, 1.)),
])
def test_scaleup_functions_inside_the_stochastic_loops(golden, method, tag, kind, times):
    # BASELINE config 4: the reference's TB class with calculate_vars calling
    # calculate_scaleup_vars() first; its sigmoidal closures branch on time and call math.exp,
    # the tracer explores the branches and the oracle evaluates them at every member's own time
    m = helpers.make_model('tb_scaleup')
    m.make_times(*times)
    soln, _, err = helpers.oracle_stochastic(m, method, oracle.Rng(kind, 9))
    assert err == 0
    assert_bitwise(soln, golden['tb_scaleup_%s_seed9' % tag])


@pytest.mark.parametrize('k', range(4))
def test_injected_stream_runs(golden, k):
    for tag, method, seed in (('dts', 'discrete_time_stochastic', 100 + k),
                              ('cts', 'continuous_time_stochastic', 200 + k)):
        m = helpers.make_model('strains')
        m.make_times(0, 150, 1)
        rng = oracle.Rng('inject', uniforms=make_golden.injected_stream(seed, 20000))
        soln, _, err = helpers.oracle_stochastic(m, method, rng)
        assert err == 0
        assert_bitwise(soln, golden['strains_%s_inject%d' % (tag, k)])
        assert rng.consumed == int(golden['strains_%s_inject%d_consumed' % (tag, k)][0])


def test_pick_event_known_answers(golden):
    for row, u, want in zip(golden['pick_intervals'], golden['pick_uniform'], golden['pick_event']):
        intervals = row[row >= 0]
        assert oracle.pick_event(intervals, u) == want


def test_stream_exhaustion_is_reported():
    m = helpers.make_model('strains')
    m.make_times(0, 50, 1)
    rng = oracle.Rng('inject', uniforms=np.full(10, 0.5))
    _, _, err = helpers.oracle_stochastic(m, 'discrete_time_stochastic', rng)
    assert err == 4
