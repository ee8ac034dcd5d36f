"""Shared helpers of the test-suite: model builders, oracle drivers, numpy statistics."""
import numpy as np

from examples import models
from oracle import oracle
from popdynamics_b200 import compiler

RMIT = dict(models.RMIT_FIXED_PARAMS)


def make_model(name, **kw):
    if name == 'sir':
        m = models.SirModel(kw.get('params', {}))
    elif name == 'mix':
        m = models.MixSirModel(kw.get('params', {}))
    elif name == 'seir_measles':
        m = models.SeirModel(models.SEIR_PARAMS['measles'])
    elif name == 'seir_flu':
        m = models.SeirModel(models.SEIR_PARAMS['flu'])
    elif name == 'seird':
        m = models.SeirDemographyModel(kw.get('params', models.SEIR_PARAMS['measles']))
    elif name == 'strains':
        m = models.StrainsModel(kw.get('params', []))
    elif name == 'tb':
        m = models.SimpleTbModel([])
    elif name == 'tb_vacc':
        m = models.SimpleTbModel(['vaccination'])
    elif name == 'tb_scaleup':
        m = models.ScaleupTbModel(kw.get('interventions', []))
    elif name == 'rmit':
        m = models.RmitTbModel(dict(RMIT, **kw.get('params', {'beta': 50.})))
    else:
        raise KeyError(name)
    return m


def compiled(model, method, **kw):
    model.check_flow_transfers()
    return compiler.compile_model(model, method, **kw)


def scaleup_table(model, cm, target_times, min_dt=0.05):
    if not cm.su_labels:
        return None
    t_eval, dt, n_sub = compiler.explicit_schedule(target_times, min_dt)
    return np.array([[model.scaleup_fns[l](t) for l in cm.su_labels] for t in t_eval.tolist()])


def oracle_explicit(model, **kw):
    cm = compiled(model, 'explicit', **kw)
    om = oracle.OracleModel(cm)
    soln, n, err = om.explicit(model.get_init_list(), model.target_times,
                               su_table=scaleup_table(model, cm, model.target_times))
    return soln, n, err


def oracle_stochastic(model, method, rng):
    cm = compiled(model, method)
    om = oracle.OracleModel(cm)
    fn = om.tau_leap if method == 'discrete_time_stochastic' else om.gillespie
    return fn(model.get_init_list(), model.target_times, rng)


def inverted_cdf_quantiles(samples, q):
    """numpy restatement of the library's quantile definition: smallest value v with
    count(x <= v) >= ceil(q * n), at least rank 1."""
    s = np.sort(np.asarray(samples))
    n = len(s)
    out = []
    for qq in q:
        rank = max(1, int(np.ceil(qq * n)))
        out.append(s[rank - 1])
    return np.array(out)


def oracle_hybrid(make_model_fn, segments, cutoff, rng_for_segment,
                  stochastic_method='continuous_time_stochastic'):
    """The loop of examples/mix_model.py:107-136 over the oracle: first segment stochastic, then
    per segment explicit when min(compartments) > cutoff; every segment is a continued run
    starting from the last row.  ``rng_for_segment(k)`` gives the uniform source of segment k.
    Returns (soln rows of the concatenated grid, list of modes)."""
    m = make_model_fn()
    rows, modes, state = None, [], None
    for k, seg in enumerate(segments):
        m.make_times(*seg)
        if k == 0:
            state = [float(v) for v in m.get_init_list()]
            method = stochastic_method
        else:
            method = 'explicit' if min(state) > cutoff else stochastic_method
        modes.append(method)
        cm = compiled(m, method, is_continue=(k > 0))
        om = oracle.OracleModel(cm)
        if method == 'explicit':
            soln, _, err = om.explicit(state, m.target_times)
        else:
            fn = om.gillespie if method == 'continuous_time_stochastic' else om.tau_leap
            soln, _, err = fn(state, m.target_times, rng_for_segment(k))
        assert err == 0
        rows = soln if rows is None else np.concatenate([rows, soln[1:]], 0)
        state = [float(v) for v in soln[-1]]
    return rows, modes
