"""GPU parity (run on the B200 box): the CUDA kernels, called through the C-ABI, against the CPU
oracle on the same seeded inputs and against the reference's golden vectors.

Bars: explicit Euler -- bit-exact fp64 (the north-star tolerance is rel 1e-10; --fmad=false and
preserved program order make it exact); stochastic integrators -- bit-exact integer counts, both
under an injected uniform stream and with the native Philox streams."""
import os

import numpy as np
import pytest

from oracle import make_golden, oracle
from popdynamics_b200 import capi, ensemble

import helpers

pytestmark = pytest.mark.gpu

EXPLICIT_RTOL = 1e-10      # north-star tolerance; the assertions below demand exact equality


def test_extension_is_loaded_from_the_tree():
    import os
    assert os.path.dirname(capi.LIB_PATH).endswith('popdynamics_b200')
    assert capi.lib().pd_version() == 100 and capi.device_count() >= 1


# ------------------------------------------------------------------------------- explicit
@pytest.mark.parametrize('name,key,times', [
    ('sir', 'sir_explicit', (0, 50, 1)),
    ('seir_measles', 'seir_measles_explicit', (0, 200, 1)),
    ('seir_flu', 'seir_flu_explicit', (0, 200, 1)),
    ('seird', 'seird_explicit_2000', (0, 2000, 1)),
    ('rmit', 'rmit_explicit', (0, 500., 1.)),
    ('strains', 'strains_explicit', (0, 200, 1)),
])
def test_dropin_integrate_explicit_matches_reference_golden(golden, name, key, times):
    m = helpers.make_model(name)
    m.make_times(*times)
    m.integrate('explicit')
    assert m.soln_array.shape == golden[key].shape
    np.testing.assert_allclose(m.soln_array, golden[key], rtol=EXPLICIT_RTOL, atol=0)
    assert np.array_equal(m.soln_array, golden[key])
    assert m.times == m.target_times


@pytest.mark.parametrize('name', ['tb', 'tb_vacc'])
def test_dropin_tb_with_scaleups_and_diagnostics(golden, name):
    m = helpers.make_model(name)
    m.make_times(1900, 2050, .05)
    m.integrate(method='explicit')
    assert np.array_equal(m.soln_array[::50], golden[name + '_explicit_every50'])
    assert np.array_equal(m.soln_array[-1], golden[name + '_explicit_final'])
    # diagnostics pass incl. the stale scale-up quirk
    assert m.var_labels == list(golden[name + '_var_labels'])
    assert np.array_equal(m.var_array[::50], golden[name + '_var_array_every50'])


def test_dropin_sir_diagnostics_and_vars(golden):
    m = helpers.make_model('sir')
    m.make_times(0, 50, 1)
    m.integrate()
    assert m.var_labels == list(golden['sir_explicit_var_labels'])
    assert np.array_equal(m.var_array, golden['sir_explicit_var_array'])
    assert np.array_equal(m.flow_array, golden['sir_explicit_flow_array'])
    assert np.array_equal(np.array(m.fraction_soln['immune']),
                          golden['sir_explicit_fraction_immune'])
    r = helpers.make_model('rmit')
    r.make_times(0, 500., 1.)
    r.integrate(method='explicit')
    assert r.vars['proportion'] == golden['rmit_proportion'][0]   # rmit_tb_model.py:143


def test_dropin_is_continue_chain(golden):
    m = helpers.make_model('mix', params={'start_population': 100})
    m.make_times(0, 5, 0.2)
    m.integrate('explicit')
    m.make_times(5, 10, 0.2)
    m.integrate('explicit', is_continue=True)
    assert np.array_equal(m.soln_array, golden['mix_continue_explicit'])
    assert np.array_equal(np.array(m.times, dtype=float), golden['mix_continue_times'])


def test_explicit_parameter_sweep_matches_oracle_per_member(golden):
    n = 300
    rng = np.random.default_rng(0)
    params = dict(sigma1=rng.uniform(0, .5, n), sigma2=rng.uniform(0, .5, n),
                  sigma3=rng.uniform(0, .5, n), beta=rng.uniform(1., 500., n),
                  rho=rng.uniform(0., 10., n), tau=2.)
    m = helpers.make_model('rmit', params=params)
    m.make_times(0, 60., 1.)
    res = m.integrate_ensemble('explicit', output='full')
    assert res.soln.shape == (61, 6, n)
    om = oracle.OracleModel(helpers.compiled(m, 'explicit'))
    want = om.explicit_batch(n, m.get_init_list(), m.target_times)
    assert want['err'] == 0
    assert np.array_equal(np.transpose(res.soln, (2, 0, 1)), want['soln'])
    final = m.integrate_ensemble('explicit', output='final').final
    assert np.array_equal(final.T, want['final'])
    # the six golden (beta) points of the reference's own figure-8 sweep
    betas = golden['rmit_sweep_betas']
    m = helpers.make_model('rmit', params=dict(sigma1=0.25, sigma2=0.125, sigma3=0.125, tau=2.,
                                               beta=betas.copy(), rho=0.5))
    m.make_times(0, 500., 1.)
    final = m.integrate_ensemble('explicit', output='final').final
    assert np.array_equal(final.T, golden['rmit_sweep_final'])


def test_explicit_per_member_initial_state_and_ragged_sizes():
    for n in (1, 31, 257):
        start = np.linspace(1., 40., n)
        m = helpers.make_model('sir', params={'population': 500.0, 'start_infectious': start})
        m.make_times(0, 12, 1)
        res = m.integrate_ensemble('explicit', output='full')
        om = oracle.OracleModel(helpers.compiled(m, 'explicit'))
        y0 = np.stack([500.0 - start, start, np.zeros(n)], axis=1)
        want = om.explicit_batch(n, y0, m.target_times)
        assert np.array_equal(np.transpose(res.soln, (2, 0, 1)), want['soln'])


def test_single_time_point_and_empty_interval():
    m = helpers.make_model('sir')
    m.make_times(3, 3, 1)
    m.integrate()
    assert m.soln_array.shape == (1, 3) and list(m.soln_array[0]) == [49., 1., 0.]


def test_device_checks_raise_assertionerror():
    class Exploding(helpers.models.SirModel):
        def calculate_vars(self):
            helpers.models.SirModel.calculate_vars(self)
            self.vars['rate_force'] = self.vars['rate_force'] / (self.compartments['immune'] * 0.)

    m = Exploding()
    m.make_times(0, 3, 1)
    with pytest.raises(AssertionError):          # NaN compartments fail `>= 0` (basepop.py:1025)
        m.integrate()


def test_default_checks_report_the_reference_failure_index():
    # explicit Euler at dt = .05 diverges for these RMIT parameters: the compartments overflow to
    # NaN and the default checks() (basepop.py:1023-1025) fails; the target index the device error
    # flag reports must be the one at which the reference's per-sub-step assert fires.
    cases = [dict(rho=0.4675096681435942, sigma1=0.4864215594357471, sigma2=0.40346161064467484,
                  sigma3=0.40408850758699627, beta=462.931574029995),
             dict(rho=6.022910254870821, sigma1=0.4835608014116187, sigma2=0.2066750375597674,
                  sigma3=0.2580841604166728, beta=498.5081362431734)]
    for params in cases:
        m = helpers.make_model('rmit', params=params)
        m.make_times(0, 500., 1.)
        om = oracle.OracleModel(helpers.compiled(m, 'explicit'))
        want = om.explicit_first_failure(m.get_init_list(), m.target_times)
        assert want > 0
        res = m.integrate_ensemble('explicit', n_members=5, output='final', check=False)
        assert res.info['err_code'] == 1 and res.info['err_where'] == want
        with pytest.raises(AssertionError, match='target index %d' % want):
            m.integrate_ensemble('explicit', n_members=5, output='final')


# ------------------------------------------------------------------------------- stochastic
def _inject(method, model_name, n_member, n_uniform, times, seed, params=None):
    m = helpers.make_model(model_name, **({'params': params} if params else {}))
    m.make_times(*times)
    u = np.random.default_rng(seed).random((n_uniform, n_member))
    res = m.integrate_ensemble(method, n_members=n_member, output='full', uniforms=u)
    om = oracle.OracleModel(helpers.compiled(m, method))
    fn = om.tau_leap if method == 'discrete_time_stochastic' else om.gillespie
    for i in range(n_member):
        rng = oracle.Rng('inject', uniforms=u[:, i].copy())
        want, _, err = fn(m.get_init_list(), m.target_times, rng)
        assert err == 0
        got = res.soln[:, :, i]
        assert np.array_equal(got, want), 'member %d differs' % i
    return res


@pytest.mark.parametrize('method', ['discrete_time_stochastic', 'continuous_time_stochastic'])
def test_injected_stream_bit_exact_strains(method):
    _inject(method, 'strains', 200, 12000, (0, 100, 1), seed=5)


def test_injected_stream_bit_exact_reaches_ptrs_branch():
    # lam >= 10: PTRS with log / sqrt / loggam on the device
    _inject('discrete_time_stochastic', 'sir', 96, 6000, (0, 40, 1), seed=6,
            params={'population': 5000, 'start_infectious': 20.})
    params = dict(helpers.models.SEIR_PARAMS['measles'], population=20000., start_infectious=10.)
    _inject('discrete_time_stochastic', 'seird', 64, 20000, (0, 60, 1), seed=7, params=params)


@pytest.mark.parametrize('method,times,n_uniform', [
    ('discrete_time_stochastic', (1985, 2000, 1.), 4000),
    ('continuous_time_stochastic', (1988, 1988.02, .01), 8000),
])
def test_injected_stream_bit_exact_tb_with_device_scaleups(method, times, n_uniform):
    # BASELINE config 4: sigmoidal scale-up closures traced (branches explored, math.exp lowered)
    # and evaluated on the device at every member's own time; population 1e6 -> every Poisson
    # mean is in the PTRS branch
    _inject(method, 'tb_scaleup', 40, n_uniform, times, seed=8)


def test_native_philox_tb_with_device_scaleups_matches_oracle():
    m = helpers.make_model('tb_scaleup')
    m.make_times(1985, 1995, 1.)
    n, seed = 70, 4
    res = m.integrate_ensemble('discrete_time_stochastic', n_members=n, output='full', seed=seed)
    om = oracle.OracleModel(helpers.compiled(m, 'discrete_time_stochastic'))
    for i in (0, 1, 37, 69):
        want, _, err = om.tau_leap(m.get_init_list(), m.target_times,
                                   oracle.Rng('philox', seed=seed, member=i))
        assert err == 0 and np.array_equal(res.soln[:, :, i], want)


def test_injected_stream_golden_runs_of_the_reference(golden):
    for tag, method, base in (('dts', 'discrete_time_stochastic', 100),
                              ('cts', 'continuous_time_stochastic', 200)):
        u = np.stack([make_golden.injected_stream(base + k, 20000) for k in range(4)], axis=1)
        m = helpers.make_model('strains')
        m.make_times(0, 150, 1)
        res = m.integrate_ensemble(method, n_members=4, output='full',
                                   uniforms=np.ascontiguousarray(u))
        for k in range(4):
            assert np.array_equal(res.soln[:, :, k], golden['strains_%s_inject%d' % (tag, k)])


@pytest.mark.parametrize('output', ['full', 'stats', 'final'])
def test_gillespie_failed_member_does_not_stall_its_warp(output):
    """One member hits log(0) (u2 == 0, basepop.py:776 raises there) in the middle of its run.  The
    recording window must let its warp finish, the error must name the member, and with full
    output every other member still equals the oracle."""
    from popdynamics_b200 import ensemble
    method = 'continuous_time_stochastic'
    m = helpers.make_model('strains')
    m.make_times(0, 30, 1)
    n = 70
    u = np.random.default_rng(21).random((4000, n))
    u[41, 3] = 0.0                                  # the second uniform of member 3's 21st event
    cm = helpers.compiled(m, method)
    kw = dict(output=output, uniforms=u)
    if output == 'stats':
        kw.update(hist_bins=2048, hist_max=2047)
    res = ensemble.run_compiled(m, cm, method, m.get_init_list(), m.target_times, n, **kw)
    assert res.info['err_code'] == 5 and res.info['err_member'] == 3
    with pytest.raises(ValueError, match='math domain error'):
        ensemble.raise_device_error(res.info, method)
    if output == 'full':
        om = oracle.OracleModel(cm)
        for i in (0, 2, 4, 31, 32, 69):
            want, _, err = om.gillespie(m.get_init_list(), m.target_times,
                                        oracle.Rng('inject', uniforms=u[:, i].copy()))
            assert err == 0 and np.array_equal(res.soln[:, :, i], want)


@pytest.mark.parametrize('method', ['discrete_time_stochastic', 'continuous_time_stochastic'])
def test_single_target_time_records_the_initial_state(method):
    m = helpers.make_model('strains')
    m.target_times = [0.]
    y0 = np.array(m.get_init_list(), dtype=np.int64)
    stats = m.integrate_ensemble(method, n_members=45, output='stats', seed=1, hist_bins=2048,
                                 hist_max=2047)
    assert np.array_equal(stats.moments[0, :, 0].astype(np.int64), 45 * y0)
    assert [int(stats.hist[0, c, y0[c]]) for c in range(5)] == [45] * 5
    full = m.integrate_ensemble(method, n_members=45, output='full', seed=1)
    assert full.soln.shape == (1, 5, 45) and np.array_equal(full.soln[0, :, 7], y0)


def test_exhausted_stream_is_reported():
    m = helpers.make_model('strains')
    m.make_times(0, 50, 1)
    with pytest.raises(RuntimeError, match='exhausted'):
        m.integrate_ensemble('discrete_time_stochastic', n_members=8, uniforms=np.full((10, 8), .5))


def test_device_philox_equals_oracle_philox():
    out = np.empty((11, 70))
    capi.philox_fill(out, 11, 70, 1000, 0xDEADBEEFCAFE, device=0)
    for j in (0, 1, 33, 69):
        rng = oracle.Rng('philox', seed=0xDEADBEEFCAFE, member=1000 + j)
        assert [rng.random() for _ in range(11)] == list(out[:, j])


@pytest.mark.parametrize('method', ['discrete_time_stochastic', 'continuous_time_stochastic'])
@pytest.mark.parametrize('count_bits', [64, 32])
def test_native_philox_bit_exact_against_oracle(method, count_bits):
    m = helpers.make_model('strains')
    m.make_times(0, 120, 1)
    n, seed, offset = 333, 99, 5000
    res = m.integrate_ensemble(method, n_members=n, output='full', seed=seed,
                               member_offset=offset, count_bits=count_bits)
    om = oracle.OracleModel(helpers.compiled(m, method))
    want = om.stochastic_batch(method, n, m.get_init_list(), m.target_times, seed, member0=offset)
    assert want['err'] == 0
    assert np.array_equal(np.transpose(res.soln, (2, 0, 1)), want['soln'])
    assert res.info['n_steps'] == want['n_steps']
    final = m.integrate_ensemble(method, n_members=n, output='final', seed=seed,
                                 member_offset=offset, count_bits=count_bits).final
    assert np.array_equal(final.T, want['soln'][:, -1, :])


def test_tau_leap_on_an_irregular_grid_with_repeated_and_decreasing_times():
    # the code tables are built for the first dt only: other steps compute their codes; a
    # repeated target time (dt = 0) draws nothing; a decreasing one gives negative Poisson means,
    # which numpy rejects with ValueError
    m = helpers.make_model('strains')
    m.target_times = [0., 1., 2., 2.5, 2.5, 4., 4.25, 9., 10., 11.]
    n, seed = 200, 21
    for count_bits in (32, 64):
        res = m.integrate_ensemble('discrete_time_stochastic', n_members=n, output='full',
                                   seed=seed, count_bits=count_bits)
        om = oracle.OracleModel(helpers.compiled(m, 'discrete_time_stochastic'))
        want = om.stochastic_batch('discrete_time_stochastic', n, m.get_init_list(),
                                   m.target_times, seed)
        assert want['err'] == 0
        assert np.array_equal(np.transpose(res.soln, (2, 0, 1)), want['soln'])
        assert np.array_equal(res.soln[3], res.soln[4])              # dt = 0: nothing happens
    m.target_times = [0., 1., 0.5, 2.]
    om = oracle.OracleModel(helpers.compiled(m, 'discrete_time_stochastic'))
    assert om.stochastic_batch('discrete_time_stochastic', 4, m.get_init_list(), m.target_times,
                               seed)['err'] == 2
    with pytest.raises(ValueError, match='Poisson mean'):
        m.integrate_ensemble('discrete_time_stochastic', n_members=4, output='final', seed=seed)
    # without the tables the trajectories are the same (POPDYN_CODE_TABLE=0 is read per launch)
    import os
    m.target_times = [float(t) for t in range(40)]
    with_tables = m.integrate_ensemble('discrete_time_stochastic', n_members=n, output='final',
                                       seed=seed).final
    os.environ['POPDYN_CODE_TABLE'] = '0'
    try:
        without = m.integrate_ensemble('discrete_time_stochastic', n_members=n, output='final',
                                       seed=seed).final
    finally:
        del os.environ['POPDYN_CODE_TABLE']
    assert np.array_equal(with_tables, without)
    # the pair tables of the two force-of-infection rows (count x count x beta): switched off,
    # and so small that members leave them and come back (the rows then compute their codes):
    # the same trajectories, and the oracle's
    m.make_times(0, 300, 1)
    om = oracle.OracleModel(helpers.compiled(m, 'discrete_time_stochastic'))
    want = om.stochastic_batch('discrete_time_stochastic', n, m.get_init_list(), m.target_times,
                               seed)
    assert want['err'] == 0
    finals = []
    for setting in (None, '0,0', '1100,6', '8,4'):
        if setting is not None:
            os.environ['POPDYN_CODE_TABLE2'] = setting
        try:
            finals.append(m.integrate_ensemble('discrete_time_stochastic', n_members=n,
                                               output='final', seed=seed).final)
        finally:
            os.environ.pop('POPDYN_CODE_TABLE2', None)
        assert np.array_equal(finals[-1].T, want['soln'][:, -1, :]), setting


def test_shard_invariance_by_global_member_id():
    m = helpers.make_model('strains')
    m.make_times(0, 60, 1)
    whole = m.integrate_ensemble('discrete_time_stochastic', n_members=1000, output='final', seed=7)
    a = m.integrate_ensemble('discrete_time_stochastic', n_members=400, output='final', seed=7)
    b = m.integrate_ensemble('discrete_time_stochastic', n_members=600, output='final', seed=7,
                             member_offset=400)
    assert np.array_equal(whole.final, np.concatenate([a.final, b.final], axis=1))


def test_stochastic_parameter_sweep_per_member():
    n = 128
    r0 = np.linspace(2.5, 4.0, n)
    m = helpers.make_model('strains', params=[('r0_invader', r0)])
    m.make_times(0, 80, 1)
    res = m.integrate_ensemble('discrete_time_stochastic', output='full', seed=3)
    om = oracle.OracleModel(helpers.compiled(m, 'discrete_time_stochastic'))
    want = om.stochastic_batch('discrete_time_stochastic', n, m.get_init_list(), m.target_times, 3)
    assert np.array_equal(np.transpose(res.soln, (2, 0, 1)), want['soln'])


def test_dropin_integrate_stochastic_single_member():
    for method in ('discrete_time_stochastic', 'continuous_time_stochastic'):
        m = helpers.make_model('sir')
        m.seed = 1234
        m.make_times(0, 50, 1)
        m.integrate(method)
        om = oracle.OracleModel(helpers.compiled(m, method))
        want = om.stochastic_batch(method, 1, m.get_init_list(), m.target_times, 1234)
        assert np.array_equal(m.soln_array, want['soln'][0])
        assert m.soln_array.sum(axis=1).tolist() == [50.] * 51          # closed population
        assert 'prevalence' in m.var_labels


def test_continued_stochastic_segments_do_not_replay_the_same_stream():
    # with a fixed model.seed, integrate(is_continue=True) must not restart the Philox stream:
    # two consecutive segments from the SAME state with the same grid would otherwise repeat
    def two_segments():
        m = helpers.make_model('sir', params={'population': 400, 'start_infectious': 40.})
        m.seed = 77
        m.make_times(0, 5, 1)
        m.integrate('continuous_time_stochastic')
        first = m.soln_array.copy()
        m.make_times(5, 10, 1)
        m.integrate('continuous_time_stochastic', is_continue=True)
        return first, m.soln_array[5:], m.last_run_info['stream']
    a_first, a_second, stream = two_segments()
    b_first, b_second, _ = two_segments()
    assert stream == 1
    assert np.array_equal(a_first, b_first) and np.array_equal(a_second, b_second)   # reproducible
    # the continued segment is what the oracle produces from stream 1 of the same seed
    m = helpers.make_model('sir', params={'population': 400, 'start_infectious': 40.})
    m.make_times(5, 10, 1)
    om = oracle.OracleModel(helpers.compiled(m, 'continuous_time_stochastic'))
    want, _, err = om.gillespie([float(v) for v in a_first[-1]], m.target_times,
                                oracle.Rng('philox', seed=77, member=1))
    assert err == 0 and np.array_equal(want, a_second)


def test_ensemble_continue_from_is_is_continue_for_every_member():
    n, seed = 120, 9
    m = helpers.make_model('strains')
    for method in ('discrete_time_stochastic', 'explicit'):
        m.make_times(0, 30, 1)
        first = m.integrate_ensemble(method, n_members=n, output='final', seed=seed)
        m.make_times(30, 50, 1)
        second = m.integrate_ensemble(method, output='full', seed=seed, continue_from=first)
        assert second.segment == 1 and second.n_members == n
        om = oracle.OracleModel(helpers.compiled(m, method, is_continue=True))
        for i in (0, 60, n - 1):
            start = [float(v) for v in first.final[:, i]]
            if method == 'explicit':
                want, _, err = om.explicit(start, m.target_times)
            else:
                want, _, err = om.tau_leap(start, m.target_times,
                                           oracle.Rng('philox', seed=seed, member=(1 << 40) + i))
            assert err == 0 and np.array_equal(second.soln[:, :, i], want)


def test_mix_model_hybrid_switching_runs():
    # examples/mix_model.py:107-136: per-day switching between CTS and explicit via is_continue
    m = helpers.make_model('mix', params={'start_population': 100})
    m.seed = 5
    m.make_times(0, 1, 0.2)
    m.integrate('continuous_time_stochastic')
    day, expected_rows = 1, len(m.target_times)
    while day < 12:
        method = 'explicit' if min(m.compartments.values()) > 3 else 'continuous_time_stochastic'
        m.make_times(day, day + 1, 0.2)      # accumulating grid: 5 or 6 points per day
        expected_rows += len(m.target_times) - 1
        m.integrate(method=method, is_continue=True)
        day += 1
    assert m.soln_array.shape[0] == len(m.times) == expected_rows
    assert np.all(m.soln_array.sum(axis=1) <= 100.0 + 1e-9)
    assert np.all(m.soln_array >= 0)


def test_hybrid_switching_ensemble_matches_the_oracle_loop_per_member():
    # the mix_model.py loop batched -- both integrators launched over the same
    # state with complementary member masks, stream id = segment << 40 | member
    n, seed, cutoff = 300, 17, 3
    segments = [(d, d + 1, 0.2) for d in range(14)]
    m = helpers.make_model('mix', params={'start_population': 100})
    res = ensemble.integrate_hybrid(m, segments, n, cutoff, seed=seed, output='full')
    soln = res.soln.cpu().numpy()
    modes = res.modes.cpu().numpy()
    assert 0 < modes[1:].sum() < modes[1:].size            # both integrators were exercised
    for i in (0, 1, 150, n - 1):
        rows, want_modes = helpers.oracle_hybrid(
            lambda: helpers.make_model('mix', params={'start_population': 100}), segments, cutoff,
            lambda k: oracle.Rng('philox', seed=seed, member=(k << 40) + i))
        assert [int(v) for v in modes[:, i]] == [int(w == 'explicit') for w in want_modes]
        assert np.array_equal(soln[:, :, i], rows), 'member %d' % i
    final = ensemble.integrate_hybrid(m, segments, n, cutoff, seed=seed, output='final').final
    assert np.array_equal(final.cpu().numpy(), soln[-1])


# ------------------------------------------------------------------------------- statistics
@pytest.mark.parametrize('method', ['discrete_time_stochastic', 'continuous_time_stochastic'])
def test_streaming_statistics_equal_numpy_over_full_trajectories(method):
    m = helpers.make_model('strains')
    m.make_times(0, 70, 1)
    n, seed, bins = 3000, 11, 2048
    full = m.integrate_ensemble(method, n_members=n, output='full', seed=seed).soln
    stats = m.integrate_ensemble(method, n_members=n, output='stats', seed=seed, hist_bins=bins,
                                 hist_max=2047)
    counts = full.astype(np.int64)                                   # [T][C][M]
    assert np.array_equal(stats.moments[..., 0].astype(np.int64), counts.sum(axis=2))
    assert np.array_equal(stats.moments[..., 1].astype(np.int64), (counts ** 2).sum(axis=2))
    np.testing.assert_allclose(stats.mean(), full.mean(axis=2), rtol=1e-12)
    np.testing.assert_allclose(stats.var(), full.var(axis=2), rtol=1e-9, atol=1e-9)
    assert list(stats.hist_shift) == [0] * 5
    for t in (0, 1, 35, 70):
        for c in range(5):
            want = np.bincount(np.minimum(counts[t, c], bins - 1), minlength=bins)
            assert np.array_equal(stats.hist[t, c], want)
    q = [0.025, 0.25, 0.5, 0.75, 0.975]
    quant = stats.quantiles(q)
    for t in (1, 35, 70):
        for c in range(5):
            assert list(quant[t, c]) == list(helpers.inverted_cdf_quantiles(counts[t, c], q))


@pytest.mark.parametrize('method', ['discrete_time_stochastic', 'continuous_time_stochastic'])
def test_moments_derived_from_lossless_histograms(method):
    m = helpers.make_model('strains')
    m.make_times(0, 50, 1)
    kw = dict(n_members=2500, output='stats', seed=4, hist_bins=2048, hist_max=2047)
    ref = m.integrate_ensemble(method, **kw)
    got = m.integrate_ensemble(method, moments_from='hist', **kw)
    assert np.array_equal(got.hist, ref.hist) and np.array_equal(got.moments, ref.moments)
    dev = m.integrate_ensemble(method, moments_from='hist', statistics='device', **kw)
    assert np.array_equal(dev.moments.cpu().numpy().view(np.uint64), ref.moments)
    # a count that reaches the last bin cannot be represented: loud error, not wrong moments
    with pytest.raises(RuntimeError, match='range'):
        m.integrate_ensemble(method, n_members=500, output='stats', seed=4, hist_bins=512,
                             hist_max=511, moments_from='hist')
    with pytest.raises(ValueError, match='width-1'):
        m.integrate_ensemble(method, n_members=500, output='stats', seed=4, hist_bins=512,
                             hist_max=4095, moments_from='hist')


def test_hist_moments_specialisation_must_match_the_run(monkeypatch):
    """moments_from_hist is a kernel specialisation (PD_FLAG_HIST_MOMENTS): a model built without
    the flag refuses such a run instead of silently skipping the moments; n_steps and ragged last
    blocks (members leave the kernel early there) are unaffected."""
    from popdynamics_b200 import capi
    m = helpers.make_model('strains')
    m.make_times(0, 20, 1)
    kw = dict(n_members=333, output='stats', seed=8, hist_bins=2048, hist_max=2047)
    ok = m.integrate_ensemble('discrete_time_stochastic', moments_from='hist', **kw)
    ref = m.integrate_ensemble('discrete_time_stochastic', **kw)
    assert np.array_equal(ok.moments, ref.moments) and np.array_equal(ok.hist, ref.hist)
    assert ok.info['n_steps'] == ref.info['n_steps'] == 333 * 20
    assert int(ok.hist[5].sum()) == 333 * 5
    monkeypatch.setattr(capi, 'FLAG_HIST_MOMENTS', 0)
    with pytest.raises(capi.PopdynError, match='PD_FLAG_HIST_MOMENTS'):
        m.integrate_ensemble('discrete_time_stochastic', moments_from='hist', **kw)


def test_statistics_large_counts_binned_and_subset_of_times():
    params = dict(helpers.models.SEIR_PARAMS['measles'], population=200000., start_infectious=50.)
    m = helpers.make_model('seird', params=params)
    m.make_times(0, 30, 1)
    n = 700
    full = m.integrate_ensemble('discrete_time_stochastic', n_members=n, output='full', seed=2).soln
    stats = m.integrate_ensemble('discrete_time_stochastic', n_members=n, output='stats', seed=2,
                                 hist_bins=256, hist_max=200100., hist_times=[0, 10, 30])
    counts = full.astype(np.int64)
    assert np.array_equal(stats.moments[..., 0].astype(np.int64), counts.sum(axis=2))
    assert np.array_equal(stats.moments[..., 1].astype(np.int64), (counts ** 2).sum(axis=2))
    assert stats.hist.shape == (3, 4, 256) and list(stats.hist_times) == [0, 10, 30]
    shift = stats.hist_shift
    assert shift[0] == 10                    # (200100 + 1) / 256 -> 2**10 per bin
    for r, t in enumerate([0, 10, 30]):
        for c in range(4):
            want = np.bincount(np.minimum(counts[t, c] >> shift[c], 255), minlength=256)
            assert np.array_equal(stats.hist[r, c], want)


def test_statistics_kept_on_the_device_equal_host_statistics():
    q = [0.025, 0.5, 0.975]
    kw = dict(n_members=3000, output='stats', seed=12, hist_bins=2048, hist_max=2047)
    m = helpers.make_model('strains')
    m.make_times(0, 60, 1)
    host = m.integrate_ensemble('discrete_time_stochastic', **kw)
    dev = m.integrate_ensemble('discrete_time_stochastic', statistics='device', **kw)
    assert hasattr(dev.hist, 'data_ptr') and dev.hist.is_cuda and dev.moments.is_cuda
    assert dev.info['d2h_bytes'] < 4096          # no histogram crossed PCIe
    assert np.array_equal(dev.moments.cpu().numpy().view(np.uint64), host.moments)
    assert np.array_equal(dev.hist.cpu().numpy().view(np.uint32), host.hist)
    assert np.array_equal(dev.quantiles(q).cpu().numpy(), host.quantiles(q))
    assert np.array_equal(dev.mean(), host.mean()) and np.array_equal(dev.var(), host.var())


def test_resident_torch_buffers_are_used_in_place():
    import torch
    m = helpers.make_model('strains')
    m.make_times(0, 40, 1)
    n = 512
    soln = torch.zeros((41, 5, n), dtype=torch.float64, device='cuda:0')
    res = m.integrate_ensemble('discrete_time_stochastic', n_members=n, output='full', seed=4,
                               out=soln)
    host = m.integrate_ensemble('discrete_time_stochastic', n_members=n, output='full', seed=4)
    assert res.info['h2d_bytes'] < 4096 and res.info['d2h_bytes'] == 0
    assert host.info['d2h_bytes'] == 41 * 5 * n * 8
    assert np.array_equal(soln.cpu().numpy(), host.soln)
    moments = torch.zeros((41, 5, 2), dtype=torch.int64, device='cuda:0')
    m.integrate_ensemble('discrete_time_stochastic', n_members=n, output='stats', seed=4,
                         moments=moments)
    assert np.array_equal(moments[..., 0].cpu().numpy(), host.soln.sum(axis=2).astype(np.int64))


# ------------------------------------------------------------------------------- diagnostics
def _host_diagnostics(make_member_model, method, times, trajectory, info=None):
    """The drop-in's own calculate_diagnostics (the reference's algorithm, basepop.py:911-960,
    golden-tested against the reference) on one member's stored trajectory."""
    m = make_member_model()
    m.make_times(*times)
    m.check_flow_transfers()
    m.soln_array = np.array(trajectory)
    m.times = list(m.target_times)
    if info is not None and m.scaleup_fns:
        # the reference leaves the scale-up vars of the LAST derivative call in self.vars
        m.time = info['last_eval_time']
        m.calculate_scaleup_vars()
    m.calculate_diagnostics()
    return m


@pytest.mark.parametrize('method', ['explicit', 'discrete_time_stochastic',
                                    'continuous_time_stochastic'])
def test_batched_diagnostics_equal_the_host_pass_per_member(method):
    n = 150
    r0 = np.linspace(2.5, 4.0, n)
    times = (0, 40, 1)
    m = helpers.make_model('strains', params=[('r0_invader', r0)])
    m.make_times(*times)
    res = m.integrate_ensemble(method, output='full', seed=3)
    diag = ensemble.diagnostics(m, res, want_flows=True)
    assert diag.vars.shape == (41, len(diag.var_labels), n)
    for i in (0, 77, n - 1):
        host = _host_diagnostics(
            lambda: helpers.make_model('strains', params=[('r0_invader', float(r0[i]))]),
            method, times, res.soln[:, :, i])
        assert diag.var_labels == host.var_labels
        assert np.array_equal(diag.vars[:, :, i], host.var_array)
        assert np.array_equal(diag.flows[:, :, i], host.flow_array)
        frac = diag.fraction('susceptible', res.soln)[:, i]
        assert np.array_equal(frac, np.array(host.fraction_soln['susceptible']))


def test_batched_diagnostics_final_state_sweep_and_stale_scaleups(golden):
    # BASELINE config 5 reads model.vars['proportion'] of the final state (rmit_tb_model.py:143)
    betas = golden['rmit_sweep_betas']
    m = helpers.make_model('rmit', params=dict(sigma1=0.25, sigma2=0.125, sigma3=0.125, tau=2.,
                                               beta=betas.copy(), rho=0.5))
    m.make_times(0, 500., 1.)
    res = m.integrate_ensemble('explicit', output='final')
    diag = ensemble.diagnostics(m, res)
    final = golden['rmit_sweep_final']                               # [n][C] from the reference
    # numpy.float64 compartments: Python's sum() is the plain left fold there
    population = np.array([sum(np.float64(v) for v in row) for row in final])
    want = final[:, 4] / population
    assert np.array_equal(diag.var('proportion'), want)
    assert diag.vars.shape == (len(diag.var_labels), len(betas))
    # TB model: the scale-up var keeps the value of the last derivative call
    m = helpers.make_model('tb')
    m.make_times(1900, 2050, .05)
    res = m.integrate_ensemble('explicit', n_members=3, output='full')
    diag = ensemble.diagnostics(m, res)
    assert diag.var_labels == list(golden['tb_var_labels'])
    assert np.array_equal(diag.vars[::50, :, 1], golden['tb_var_array_every50'])


@pytest.mark.parametrize('method', ['discrete_time_stochastic', 'continuous_time_stochastic'])
def test_user_defined_checks_run_inside_the_stochastic_kernels(method):
    # basepop.py:789 / :846: checks() after the update, lowered through the traced-assert
    # mechanism: updated counts (y2), vars of the step's start; the oracle fires identically
    class Limited(helpers.models.StrainsModel):
        limit = 10 ** 6

        def checks(self, error_margin=0.1):
            assert self.compartments['infectious_invader'] <= self.limit
            assert self.vars['population'] + 50 >= sum(self.compartments.values())
            assert self.vars['rate_death'] < 10.

    n, seed = 512, 21
    m = Limited()
    m.make_times(0, 60, 1)
    res = m.integrate_ensemble(method, n_members=n, output='full', seed=seed)
    plain = helpers.make_model('strains')
    plain.make_times(0, 60, 1)
    want = plain.integrate_ensemble(method, n_members=n, output='full', seed=seed)
    assert np.array_equal(res.soln, want.soln)            # passing checks change nothing
    Limited.limit = 3
    m = Limited()
    m.make_times(0, 60, 1)
    res = m.integrate_ensemble(method, n_members=n, output='full', seed=seed, check=False)
    assert res.info['err_code'] == 1
    bad = int(res.info['err_member'])
    om = oracle.OracleModel(helpers.compiled(m, method))
    fn = om.tau_leap if method.startswith('discrete') else om.gillespie
    _, _, err = fn(m.get_init_list(), m.target_times, oracle.Rng('philox', seed=seed, member=bad))
    assert err == 1
    over = want.soln[:, 2, :].max(axis=0) > 3             # members whose invader passes 3
    assert over.any()
    if method.startswith('discrete'):                     # checked exactly at the stored times
        assert over[bad]
    with pytest.raises(AssertionError):
        m.integrate_ensemble(method, n_members=n, output='final', seed=seed)


def test_ptrs_squeeze_never_disagrees_with_the_exact_acceptance_test(monkeypatch):
    # pd_common.cuh: the slow PTRS acceptance test is decided by a certified squeeze; a build with
    # PD_PTRS_CHECK evaluates BOTH for every slow test and raises error 9 on any disagreement.
    # Means from 10 to ~1e7 (population swept over six decades), ~1e8 slow tests.
    n = 1 << 17
    pop = np.round(10 ** np.random.default_rng(8).uniform(2.5, 8.5, n))
    def build():
        m = helpers.make_model('sir', params={'population': pop, 'start_infectious': 20.})
        m.make_times(0, 120, 1)
        return m
    plain = build().integrate_ensemble('discrete_time_stochastic', n_members=n, output='final',
                                       seed=77)
    monkeypatch.setenv('POPDYN_DEFINE', '-DPD_PTRS_CHECK=1')
    ensemble.clear_kernel_cache()
    checked = build().integrate_ensemble('discrete_time_stochastic', n_members=n, output='final',
                                         seed=77, check=False)
    monkeypatch.setenv('POPDYN_DEFINE', '-DPD_PTRS_SQUEEZE=0')
    ensemble.clear_kernel_cache()
    exact = build().integrate_ensemble('discrete_time_stochastic', n_members=n, output='final',
                                       seed=77)
    monkeypatch.delenv('POPDYN_DEFINE')
    ensemble.clear_kernel_cache()
    assert checked.info['err_code'] == 0
    assert np.array_equal(plain.final, exact.final) and np.array_equal(plain.final, checked.final)
    m = build()
    om = oracle.OracleModel(helpers.compiled(m, 'discrete_time_stochastic'))
    for i in (0, 4097, n - 1):
        want, _, err = om.tau_leap([v[i] if isinstance(v, np.ndarray) else v
                                    for v in m.get_init_list()], m.target_times,
                                   oracle.Rng('philox', seed=77, member=i),
                                   params=om.member_params(i))
        assert err == 0 and np.array_equal(plain.final[:, i], want[-1]), i


# ------------------------------------------------------------------------------- pd_math.cuh
@pytest.mark.gpu
def test_kernel_forms_of_exp_log_and_division_equal_the_math_library_bit_for_bit():
    """csrc/kernels/pd_math.cuh restates the library's fast paths with constant-bank coefficients
    and guards the divider against zero numerators: 1e8 arguments per form, no differing bit."""
    from popdynamics_b200 import capi
    bad, first = capi.math_selfcheck(100_000_000, seed=2024)
    assert bad == [0, 0, 0, 0], (bad, first)


@pytest.mark.gpu
def test_gillespie_arithmetic_forms_do_not_change_results(tmp_path):
    """PD_GIL_OPT=0 (plain divisions, library exp / log, ternary dead rows) against the later
    builds of pd_gillespie_kernel up to the default (127: + row masks, counting search, fast
    first pass): same events, same waiting times, member for member."""
    import subprocess
    import sys
    script = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from examples import configs
m = configs.tb_gillespie()
m.make_times(1950, 1951, 0.25)
res = m.integrate_ensemble('continuous_time_stochastic', n_members=4096, output='full', seed=11)
np.save(sys.argv[1], np.asarray(res.soln.cpu() if hasattr(res.soln, 'cpu') else res.soln))
print(int(res.info['n_steps']))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for tag, define in (('plain', '-DPD_GIL_OPT=0'), ('tuned', '-DPD_GIL_OPT=15'),
                        ('counting', '-DPD_GIL_OPT=63'), ('fast_pass', '-DPD_GIL_OPT=127')):
        env = dict(os.environ, POPDYN_DEFINE=define)
        path = str(tmp_path / (tag + '.npy'))
        done = subprocess.run([sys.executable, '-c', script, path], env=env, capture_output=True,
                              text=True, timeout=600)
        assert done.returncode == 0, done.stderr[-2000:]
        outs.append((np.load(path), int(done.stdout.split()[-1])))
    assert outs[0][1] > 0
    for soln, n_events in outs[1:]:
        assert n_events == outs[0][1]
        assert np.array_equal(soln, outs[0][0])


@pytest.mark.gpu
@pytest.mark.parametrize('recover', [-0.1, float('nan'), -0.0, 0.0])
def test_gillespie_fast_pass_hands_odd_multipliers_to_the_exact_pass(recover):
    """The fast first pass of pd_gillespie_kernel takes every rate as it is when no row multiplier
    has a sign bit, is infinite or NaN; a negative, NaN or -0.0 multiplier makes rates that
    `val > 0` (basepop.py:718) must filter, so those events take the exact pass.  SIR with such a
    recovery rate (the row is never live), injected stream, against the oracle member by member."""
    method = 'continuous_time_stochastic'
    m = helpers.make_model('sir', params={'population': 60, 'start_infectious': 4.})
    m.set_param('infection_rate_recover', recover)
    m.make_times(0, 20, 2)
    n_member, n_uniform = 48, 4000
    u = np.random.default_rng(77).random((n_uniform, n_member))
    res = m.integrate_ensemble(method, n_members=n_member, output='full', uniforms=u)
    om = oracle.OracleModel(helpers.compiled(m, method))
    for i in range(n_member):
        rng = oracle.Rng('inject', uniforms=u[:, i].copy())
        want, _, err = om.gillespie(m.get_init_list(), m.target_times, rng)
        assert err == 0
        assert np.array_equal(res.soln[:, :, i], want), 'member %d differs' % i
    assert (np.asarray(res.soln)[-1, 2] == 0).all()          # nobody ever recovers


@pytest.mark.gpu
@pytest.mark.parametrize('model_name,params,times', [
    ('sir', {'population': 40, 'start_infectious': 3.}, (0, 60, 2)),   # row 0 (infection) dies out
    ('strains', None, (0, 30, 2)),                                     # row 0 is a birth row
])
def test_gillespie_pick_at_the_ends_of_the_interval_table(model_name, params, times):
    """pick_event (basepop.py:161-178) at its two edges, which the counting search of
    pd_gillespie_kernel treats apart: u1 == 0 (point 0: the first LIVE row, also when the rows
    before it are dead) and u1 == 1 - 2^-53 (the product may round up to the total: the last live
    row).  Every third reaction of every member draws one of the two."""
    method = 'continuous_time_stochastic'
    m = helpers.make_model(model_name, **({'params': params} if params else {}))
    m.make_times(*times)
    n_member, n_uniform = 64, 10000
    rng = np.random.default_rng(31)
    u = rng.random((n_uniform, n_member))
    edge = rng.random((n_uniform // 2, n_member))
    u1 = u[0::2]
    u1[edge < 1 / 6] = 0.0
    u1[edge > 5 / 6] = 1.0 - 2.0 ** -53
    res = m.integrate_ensemble(method, n_members=n_member, output='full', uniforms=u)
    om = oracle.OracleModel(helpers.compiled(m, method))
    for i in range(n_member):
        want, _, err = om.gillespie(m.get_init_list(), m.target_times,
                                    oracle.Rng('inject', uniforms=u[:, i].copy()))
        assert err == 0 and np.array_equal(res.soln[:, :, i], want), i
