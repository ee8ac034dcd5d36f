"""Statistical parity of the native Philox ensembles (run on the B200 box): two-sample KS and
moment tests against ensembles produced with the reference's own generators -- the oracle driven by
the MT19937 restatements is bit-identical to the unmodified reference (tests/test_oracle_golden.py)
-- plus the reference's analytic pin: P(invader extinction) -> r0_resident / r0_invader
(/root/reference/examples/stochastic_strains_model.py:180-196) and size-independent invariants at
large ensemble sizes."""
import numpy as np
import pytest
from scipy import stats as sps

from oracle import oracle

import helpers

pytestmark = pytest.mark.gpu


def reference_ensemble(method, n, times):
    m = helpers.make_model('strains')
    m.make_times(*times)
    om = oracle.OracleModel(helpers.compiled(m, method))
    fn, kind = (om.tau_leap, 'numpy') if method == 'discrete_time_stochastic' \
        else (om.gillespie, 'python')
    finals = []
    for seed in range(n):
        soln, _, err = fn(m.get_init_list(), m.target_times, oracle.Rng(kind, 1000 + seed))
        assert err == 0
        finals.append(soln[-1])
    return np.array(finals)                                       # [n][C]


@pytest.mark.parametrize('method,times', [('discrete_time_stochastic', (0, 200, 1)),
                                          ('continuous_time_stochastic', (0, 200, 1))])
def test_ks_and_moments_against_reference_generators(method, times):
    n_ref, n_gpu = 1500, 20000
    ref = reference_ensemble(method, n_ref, times)
    m = helpers.make_model('strains')
    m.make_times(*times)
    gpu = m.integrate_ensemble(method, n_members=n_gpu, output='final', seed=31).final.T
    for c in range(5):
        ks = sps.ks_2samp(ref[:, c], gpu[:, c])
        assert ks.pvalue > 1e-3, 'compartment %d: KS p=%g' % (c, ks.pvalue)
        se = np.sqrt(ref[:, c].var() / n_ref + gpu[:, c].var() / n_gpu) + 1e-12
        assert abs(ref[:, c].mean() - gpu[:, c].mean()) < 5 * se + 1e-9


def test_invader_extinction_probability_is_r0_ratio():
    # resident r0 = 2, invader r0 = 3 -> P(extinction of the invader) -> 2/3
    m = helpers.make_model('strains')
    m.make_times(0, 1000, 1)
    n = 200000
    res = m.integrate_ensemble('discrete_time_stochastic', n_members=n, output='final', seed=1)
    extinct = float(np.mean(res.final[2] == 0))
    # the discrete-time scheme with dt = 1 is an approximation of the branching process; the
    # continuous-time run below is the exact one
    assert abs(extinct - 2. / 3.) < 0.02
    res = m.integrate_ensemble('continuous_time_stochastic', n_members=40000, output='final',
                               seed=1)
    extinct = float(np.mean(res.final[2] == 0))
    assert abs(extinct - 2. / 3.) < 4 * np.sqrt(2. / 9. / 40000) + 0.005


def test_closed_population_is_conserved_at_scale():
    m = helpers.make_model('sir', params={'population': 1000, 'start_infectious': 5.})
    m.make_times(0, 60, 1)
    n = 1 << 20
    for method in ('discrete_time_stochastic', 'continuous_time_stochastic'):
        final = m.integrate_ensemble(method, n_members=n, output='final', seed=9).final
        assert np.all(final >= 0)
        assert np.all(final.sum(axis=0) == 1000.)
    stats = m.integrate_ensemble('discrete_time_stochastic', n_members=n, output='stats', seed=9)
    totals = stats.moments[..., 0].astype(np.int64).sum(axis=1)
    assert np.all(totals == 1000 * n)                               # a checksum over all members


def test_explicit_ensemble_of_identical_members_is_identical(golden):
    m = helpers.make_model('seir_measles')
    m.make_times(0, 200, 1)
    n = 100000
    final = m.integrate_ensemble('explicit', n_members=n, output='final').final
    assert np.all(final == golden['seir_measles_explicit'][-1][:, None])


def test_headline_configuration_at_full_size_through_size_independent_properties():
    # BASELINE config 3 as bench.py runs it: 10^7 members x 1000 tau-leap steps, statistics on
    # the device.  The oracle cannot run this size; what must hold at any size:
    #   every member is recorded exactly once at every (t, c);
    #   the first moment equals the histogram's own first moment (bin width 1, no overflow);
    #   the second moment likewise;
    #   two half-ensembles keyed by global member id add up to the whole, bit for bit.
    import torch
    m = helpers.make_model('strains')
    m.make_times(0, 1000, 1)
    n, bins = 10 ** 7, 2048
    kw = dict(output='stats', seed=5, hist_bins=bins, hist_max=bins - 1, statistics='device',
              count_bits=32)
    whole = m.integrate_ensemble('discrete_time_stochastic', n_members=n, **kw)
    hist = whole.hist.to(torch.int64)                               # [T][C][bins]
    assert bool((hist.sum(dim=2) == n).all())
    assert int(hist[..., bins - 1].sum()) == 0                      # nothing in the overflow bin
    values = torch.arange(bins, device=hist.device, dtype=torch.int64)
    assert torch.equal((hist * values).sum(dim=2), whole.moments[..., 0])
    assert torch.equal((hist * values * values).sum(dim=2), whole.moments[..., 1])
    half = n // 2
    a = m.integrate_ensemble('discrete_time_stochastic', n_members=half, **kw)
    b = m.integrate_ensemble('discrete_time_stochastic', n_members=n - half, member_offset=half,
                             **kw)
    assert torch.equal(a.moments + b.moments, whole.moments)
    assert torch.equal(a.hist + b.hist, whole.hist)
    # the reference's analytic pin at this size: P(invader extinct) -> 2/3 (dt = 1 bias < 0.02)
    extinct = float(hist[-1, 2, 0]) / n
    assert abs(extinct - 2. / 3.) < 0.02
    # the mean total population obeys E[N'] = E[N] + 2 - 0.001 E[N] exactly (Poisson means are
    # linear in the counts, clamps never bind for the deaths): N_t = 2000 - 909 * 0.999^t
    for t in (1, 10, 100, 1000):
        population = float(whole.moments[t, :, 0].sum()) / n
        expected = 2000. - (2000. - 1091.) * 0.999 ** t
        sigma = np.sqrt(2. * t + 1091.) / np.sqrt(n)               # generous bound on the s.e.
        assert abs(population - expected) < 6 * sigma + 1e-6, (t, population, expected)
