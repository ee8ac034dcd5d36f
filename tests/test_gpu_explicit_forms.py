# -*- coding: utf-8 -*-
"""The explicit Euler kernel's integer forms (PD_EXPL_OPT, csrc/kernels/pd_explicit.cuh): the sign-bit
clamp fused with the default checks() and the bit tests of the compensated sum must give the
results of the plain fp64 comparisons -- /root/reference/basepop.py:684-687 (`if y[i] < 0`),
:1016-1025 (`assert compartment >= 0`), CPython's `if c and isfinite(c)` -- on ordinary runs, on
diverging ones (infinities, NaN) and on the one state the sign bit misjudges (-0.0)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _same(a, b):
    """Equal as doubles incl. the sign of zeros, NaN where NaN (payloads are not compared: x86
    and the GPU produce different default NaNs)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nan = np.isnan(a)
    return (a.shape == b.shape and np.array_equal(nan, np.isnan(b)) and
            np.array_equal(a[~nan], b[~nan]) and
            np.array_equal(np.signbit(a[~nan]), np.signbit(b[~nan])))


_FORMS_SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, ROOT)
from examples import configs, models
out = {}
m = configs.seir_sweep(0, 4096)
m.make_times(0, 400, 1)
res = m.integrate_ensemble('explicit', n_members=4096, output='full')
out['seir'] = np.asarray(res.soln)
# sigma up to .5, beta up to 500: explicit Euler at dt = .05 diverges for some members
# (DESIGN.md 3.4; members 0 and 1 are two such parameter sets), their compartments pass through
# inf to NaN and checks() fails
p = dict(models.RMIT_FIXED_PARAMS)
p['beta'] = configs.sweep_uniform(3, 0, 1., 500., 0, 2048)
p['rho'] = configs.sweep_uniform(3, 1, 0., 10., 0, 2048)
for k in (1, 2, 3):
    p['sigma%d' % k] = configs.sweep_uniform(3, 1 + k, 0., .5, 0, 2048)
known = [dict(rho=0.4675096681435942, sigma1=0.4864215594357471, sigma2=0.40346161064467484,
              sigma3=0.40408850758699627, beta=462.931574029995),
         dict(rho=6.022910254870821, sigma1=0.4835608014116187, sigma2=0.2066750375597674,
              sigma3=0.2580841604166728, beta=498.5081362431734)]
for i, d in enumerate(known):
    for key, value in d.items():
        p[key][i] = value
m = models.RmitTbModel(p)
m.make_times(0, 500., 1.)
res = m.integrate_ensemble('explicit', n_members=2048, output='final', check=False)
out['rmit'] = np.asarray(res.final)
out['rmit_err'] = np.array([res.info['err_code']])
np.savez(sys.argv[1], **out)
'''.replace('ROOT', repr(ROOT))


@pytest.mark.gpu
def test_explicit_integer_forms_do_not_change_results(tmp_path):
    outs = []
    for tag, define in (('plain', '-DPD_EXPL_OPT=0'), ('sum', '-DPD_EXPL_OPT=1'),
                        ('clamp', '-DPD_EXPL_OPT=2'), ('both', '-DPD_EXPL_OPT=3'),
                        ('rolled', '-DPD_EXPL_UNROLL=1'), ('twosum', 'POPDYN_SUM_FORM')):
        env = dict(os.environ, POPDYN_DEFINE=define)
        if define == 'POPDYN_SUM_FORM':
            env = dict(os.environ, POPDYN_SUM_FORM='twosum')
        path = str(tmp_path / (tag + '.npz'))
        done = subprocess.run([sys.executable, '-c', _FORMS_SCRIPT, path], env=env,
                              capture_output=True, text=True, timeout=900)
        assert done.returncode == 0, done.stderr[-2000:]
        outs.append(np.load(path))
    ref = outs[0]
    assert np.isnan(ref['rmit']).any() and ref['rmit_err'][0] == 1    # the sweep does diverge
    assert np.isnan(ref['rmit'][:, :2]).any(axis=0).all()
    assert np.isfinite(ref['rmit']).all(axis=0).sum() > 1000
    for got in outs[1:]:
        assert _same(got['seir'], ref['seir'])
        assert _same(got['rmit'], ref['rmit'])
        assert got['rmit_err'][0] == ref['rmit_err'][0]


def _one_member(model):
    """One member through the kernel (default build) and through the oracle: trajectory, error
    flag, failing target index."""
    model.check_flow_transfers()
    res = model.integrate_ensemble('explicit', n_members=1, output='full', check=False)
    om = oracle.OracleModel(helpers.compiled(model, 'explicit'))
    want, _, err = om.explicit(model.get_init_list(), model.target_times)
    where = om.explicit_first_failure(model.get_init_list(), model.target_times)
    return np.asarray(res.soln)[:, :, 0], res.info, want, err, where


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['minus_zero_at_rest', 'minus_zero_filling', 'nan_param',
                                  'inf_param', 'overflow', 'nan_in_last_substep',
                                  'tiny_denormals'])
def test_explicit_fused_checks_on_the_edges(case):
    """States the fast path must hand to the exact comparisons, one member each, against the
    oracle: the trajectory (signs of zeros included), whether checks() fails and at which target
    index (the interval of the first derivative call that is handed a NaN, basepop.py:623)."""
    times = (0, 6, 1)
    m = helpers.make_model('sir')
    if case == 'minus_zero_at_rest':
        # nobody infectious: every flow is (+-)0, the immune class stays -0.0 (-0.0 + -0.0 only
        # if dt * f is -0.0 too; here f = +0.0, so it turns +0.0 at once -- as in the reference)
        m.set_compartment('infectious', 0.)
        m.set_compartment('immune', -0.0)
    elif case == 'minus_zero_filling':
        m.set_compartment('immune', -0.0)
    elif case == 'nan_param':
        m.set_param('infection_rate_recover', float('nan'))
    elif case == 'inf_param':
        m.set_param('infection_beta', float('inf'))
    elif case == 'overflow':
        m.set_param('infection_beta', 1e300)
        m.set_compartment('susceptible', 1e300)
        m.set_compartment('infectious', 1e300)
    elif case == 'nan_in_last_substep':
        # min_dt .05, targets 0.05 apart: one sub-step per interval, so a NaN made in interval i
        # is first seen by the derivative call of interval i + 1 (and never, if i was the last)
        m.set_param('infection_rate_recover', float('nan'))
        times = (0, 0.2, 0.05)
    elif case == 'tiny_denormals':
        m.set_compartment('susceptible', 5e-324)
        m.set_compartment('infectious', 3e-310)
        m.set_param('infection_rate_recover', 30.)        # dt * rate > 1: the update goes negative
    m.make_times(*times)
    got, info, want, err, where = _one_member(m)
    assert _same(got, want), (got, want)
    assert int(info['err_code'] != 0) == err
    assert (info['err_where'] if err else 0) == where
    if case in ('nan_param', 'inf_param', 'overflow', 'nan_in_last_substep'):
        assert err == 1 and where >= 1


@pytest.mark.gpu
def test_explicit_fused_checks_in_the_very_last_substep_raise_nothing():
    """A NaN produced by the last sub-step of the run is never handed to a derivative call: the
    reference's loop ends without an assertion (basepop.py:672-688), and so does the kernel."""
    m = helpers.make_model('sir')
    m.set_param('infection_rate_recover', float('nan'))
    m.make_times(0, 0.05, 0.05)                       # exactly one sub-step
    got, info, want, err, where = _one_member(m)
    assert err == 0 and where == 0 and info['err_code'] == 0
    assert np.isnan(got[-1]).any() and _same(got, want)
