"""Seeded random compartmental models (tests/random_models.py): every flow kind, vars reading
vars, divisions, up to 48 flow-table rows.

* CPU, where the reference tree exists: the oracle against the UNMODIFIED reference running the
  same model class, all three integrators, bit for bit (stochastic: both sides seeded MT19937).
* GPU: the kernels through the C-ABI against the oracle on the same models."""
import numpy as np
import pytest

from oracle import oracle, reference
from popdynamics_b200 import basepop

import helpers
import random_models

Ours = random_models.random_model_class(basepop, basepop.old_div)
CASE_IDS = ['seed%d' % seed for seed, _ in random_models.CASES]


@pytest.mark.skipif(not reference.available(), reason='reference tree not present')
@pytest.mark.parametrize('seed,kw', random_models.CASES, ids=CASE_IDS)
def test_oracle_equals_the_reference_on_random_models(seed, kw):
    ref_module = reference.basepop()
    from past.utils import old_div
    Theirs = random_models.random_model_class(ref_module, old_div)
    for method, kind in (('explicit', None), ('discrete_time_stochastic', 'numpy'),
                         ('continuous_time_stochastic', 'python')):
        theirs = Theirs(seed, **kw)
        theirs.make_times(0, 30, 1)
        reference.seed_all(seed)
        theirs.integrate(method)
        ours = Ours(seed, **kw)
        ours.make_times(0, 30, 1)
        if method == 'explicit':
            soln, _, err = helpers.oracle_explicit(ours)
        else:
            soln, _, err = helpers.oracle_stochastic(ours, method, oracle.Rng(kind, seed))
        assert err == 0
        assert np.array_equal(soln, theirs.soln_array), method


@pytest.mark.gpu
@pytest.mark.parametrize('seed,kw', random_models.CASES, ids=CASE_IDS)
def test_kernels_equal_the_oracle_on_random_models(seed, kw):
    n = 97
    m = Ours(seed, **kw)
    m.make_times(0, 25, 1)
    res = m.integrate_ensemble('explicit', n_members=n, output='full')
    want, _, err = helpers.oracle_explicit(m)
    assert err == 0
    for i in (0, n - 1):
        assert np.array_equal(res.soln[:, :, i], want)
    for method in ('discrete_time_stochastic', 'continuous_time_stochastic'):
        for count_bits in (32, 64):
            res = m.integrate_ensemble(method, n_members=n, output='full', seed=seed,
                                       count_bits=count_bits)
            om = oracle.OracleModel(helpers.compiled(m, method))
            want = om.stochastic_batch(method, n, m.get_init_list(), m.target_times, seed)
            assert want['err'] == 0
            assert np.array_equal(np.transpose(res.soln, (2, 0, 1)), want['soln']), method
        stats = m.integrate_ensemble(method, n_members=n, output='stats', seed=seed,
                                     hist_bins=64, hist_max=4095.)
        counts = np.transpose(want['soln'], (1, 2, 0)).astype(np.int64)      # [T][C][M]
        assert np.array_equal(stats.moments[..., 0].astype(np.int64), counts.sum(axis=2))
        assert np.array_equal(stats.moments[..., 1].astype(np.int64), (counts ** 2).sum(axis=2))
        shift = stats.hist_shift
        for c in range(counts.shape[1]):
            bins = np.minimum(counts[-1, c] >> shift[c], 63)
            assert np.array_equal(stats.hist[-1, c], np.bincount(bins, minlength=64))


@pytest.mark.gpu
def test_injected_stream_on_the_largest_random_model():
    seed, kw = random_models.CASES[-1]                 # 48 rows: 64-bit live mask
    m = Ours(seed, **kw)
    m.make_times(0, 12, 1)
    n, n_uniform = 40, 4000
    u = np.random.default_rng(3).random((n_uniform, n))
    for method in ('discrete_time_stochastic', 'continuous_time_stochastic'):
        res = m.integrate_ensemble(method, n_members=n, output='full', uniforms=u)
        om = oracle.OracleModel(helpers.compiled(m, method))
        fn = om.tau_leap if method == 'discrete_time_stochastic' else om.gillespie
        for i in range(n):
            want, _, err = fn(m.get_init_list(), m.target_times,
                              oracle.Rng('inject', uniforms=u[:, i].copy()))
            assert err == 0 and np.array_equal(res.soln[:, :, i], want)
