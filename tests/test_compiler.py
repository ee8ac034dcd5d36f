"""Flow compiler + calculate_vars tracer: flow-table order, dead-code elimination, parameter slots,
Python's sum() semantics, loud rejection of constructs that cannot be lowered, and NVRTC
compilation of the generated CUDA for sm_100a (needs no GPU)."""
import math

import numpy as np
import pytest

from oracle import oracle
from popdynamics_b200 import capi, compiler
from popdynamics_b200.basepop import BaseModel
from popdynamics_b200.compiler import TraceError

import helpers


def test_strains_flow_table_is_in_canonical_order():
    # birth; S->I_inv; S->I_res; I_res->R_res; I_inv->R_inv;
    # 5 background deaths in label order; 2 infection deaths
    cm = helpers.compiled(helpers.make_model('strains'), 'discrete_time_stochastic')
    t = cm.flow_table()
    assert list(t['kind']) == [0, 1, 1, 2, 2, 3, 3, 3, 3, 3, 4, 4]
    assert list(t['frm']) == [-1, 0, 0, 1, 2, 0, 1, 2, 3, 4, 2, 1]
    assert list(t['to']) == [0, 2, 1, 3, 4, -1, -1, -1, -1, -1, -1, -1]
    assert list(t['is_int']) == [1] + [0] * 11          # rate_birth is the Python int 2
    assert cm.integer_state and not cm.check_nodes


def test_dead_code_elimination_drops_unused_population():
    cm = helpers.compiled(helpers.make_model('sir'), 'explicit')
    ops = [cm.graph.nodes[n][0] for n in cm.order]
    assert ops.count('mul') == 1 and 'add' not in ops    # only beta * infectious survives
    cm = helpers.compiled(helpers.make_model('seird'), 'explicit')
    assert 'select' in [cm.graph.nodes[n][0] for n in cm.order]   # Neumaier sum feeds rate_birth


def test_param_slots_and_member_sweeps():
    betas = np.linspace(1., 500., 16)
    m = helpers.make_model('rmit', params={'beta': betas, 'rho': np.linspace(0., 10., 16)})
    cm = helpers.compiled(m, 'explicit')
    assert cm.n_members_implied() == 16
    swept = [cm.param_labels[k] for k in cm.member_param_slots()]
    assert sorted(swept) == ['beta', 'rho']
    assert 'pm[' in cm.cuda_source() and np.isnan(cm.uniform_params()).sum() == 2
    m = helpers.make_model('rmit', params={'beta': betas, 'rho': np.zeros(3)})
    with pytest.raises(ValueError):
        helpers.compiled(m, 'explicit').n_members_implied()


def test_later_set_param_is_picked_up_at_integrate_time():
    # rmit_tb_model.py:139-142 sets beta after construction; set_flows re-runs per integrate
    m = helpers.make_model('rmit', params={'beta': 50.})
    m.set_param('rho', 0.7)
    cm = helpers.compiled(m, 'explicit')
    assert cm.param_values[cm.param_labels.index('rho')] == 0.7


class _Weird(BaseModel):
    def __init__(self, body):
        BaseModel.__init__(self)
        self.body = body
        self.set_compartment('a', 10.)
        self.set_compartment('b', 0.)
        self.set_param('k', 0.5)

    def set_flows(self):
        self.set_var_transfer_rate_flow('a', 'b', 'rate')

    def calculate_vars(self):
        self.vars['rate'] = self.body(self)


@pytest.mark.parametrize('body', [
    lambda s: s.params['k'] if s.compartments['a'] > 1 else 0.,       # data-dependent branch
    lambda s: math.exp(s.compartments['a']),                          # math.* on a traced value
    lambda s: s.compartments['a'] ** 2,                               # pow
    lambda s: float(s.compartments['a']),                             # scalar conversion
    lambda s: max(s.compartments['a'], 1.),                           # builtin max -> bool()
    lambda s: np.sqrt(s.compartments['a']),                           # numpy ufunc
    lambda s: s.compartments['a'] % 3,                                # modulo
])
def test_unsupported_constructs_are_rejected_loudly(body):
    with pytest.raises(TypeError):
        helpers.compiled(_Weird(body), 'explicit')


def test_supported_arithmetic_matches_python():
    body = lambda s: abs(-s.compartments['a']) * s.params['k'] / (1. + s.compartments['b']) \
        - 0.25 + s.time * 0.
    m = _Weird(body)
    cm = helpers.compiled(m, 'explicit')
    assert cm.uses_time
    om = oracle.OracleModel(cm)
    values = om.eval_values([10., 3.], time=2.)
    assert values[om.slot_of[cm.flow_node[0]]] == abs(-10.) * 0.5 / (1. + 3.) - 0.25 + 2. * 0.


def test_missing_var_raises_keyerror_like_the_reference():
    # the stochastic loops never evaluate scale-ups -> KeyError in tb_model
    with pytest.raises(KeyError, match='program_rate_detect'):
        helpers.compiled(helpers.make_model('tb'), 'discrete_time_stochastic')
    helpers.compiled(helpers.make_model('tb'), 'explicit')


def test_sum_semantics():
    vals = [1e16, 1.0, -1e16, 3.0, 1e-3]

    class S(_Weird):
        def __init__(self):
            _Weird.__init__(self, None)
            for i, v in enumerate(vals):
                self.set_param('p%d' % i, v)

        def calculate_vars(self):
            self.vars['rate'] = sum(self.params['p%d' % i] for i in range(len(vals)))

    for naive, want in ((False, sum(vals)), (True, sum(np.float64(v) for v in vals))):
        cm = helpers.compiled(S(), 'explicit', naive_sum=naive)
        om = oracle.OracleModel(cm)
        got = om.eval_values([10., 0.])[om.slot_of[cm.flow_node[0]]]
        assert got == want
    assert sum(vals) != sum(np.float64(v) for v in vals)   # the two paths really differ


def test_custom_checks_are_traced_into_device_assertions():
    class Checked(_Weird):
        def checks(self, error_margin=0.1):
            assert self.compartments['a'] + self.compartments['b'] <= 10. + error_margin
            assert self.vars['rate'] >= 0.

    cm = helpers.compiled(Checked(lambda s: s.params['k']), 'explicit')
    assert len(cm.check_nodes) == 2
    om = oracle.OracleModel(cm)
    m = Checked(lambda s: s.params['k'])
    m.make_times(0, 2, 1)
    _, _, err = om.explicit([10., 0.], m.target_times)
    assert err == 0
    _, _, err = om.explicit([10., 5.], m.target_times)
    assert err == 1


def test_custom_checks_in_the_stochastic_loops_see_the_updated_counts_and_the_old_vars():
    # basepop.py:789 / :846: checks() runs after the update; self.vars is what calculate_vars and
    # calculate_events left at the START of the step (incl. the positive-only death totals)
    class Draining(_Weird):
        def __init__(self, floor):
            _Weird.__init__(self, lambda s: s.params['k'])
            self.floor = floor
            self.set_param('mu', 0.01)

        def set_flows(self):
            _Weird.set_flows(self)
            self.set_background_death_rate('mu')

        def calculate_vars(self):
            _Weird.calculate_vars(self)
            self.vars['a_before'] = self.compartments['a']

        def checks(self, error_margin=0.1):
            assert self.compartments['a'] >= self.floor          # updated count
            assert self.vars['a_before'] >= self.compartments['a']   # old var vs new count
            assert self.vars['rate_death'] <= 0.2                # calculate_events side output

    for method in ('discrete_time_stochastic', 'continuous_time_stochastic'):
        cm = helpers.compiled(Draining(0.), method)
        assert len(cm.post_check_nodes) == 3 and cm.check_nodes == []
        assert '#define PD_HAS_STEP_CHECKS 1' in cm.cuda_source()
        assert 'y2[0]' in cm.cuda_source()
        m = Draining(0.)
        m.make_times(0, 6, 1)
        fn = oracle.OracleModel(cm).tau_leap if method.startswith('discrete') else \
            oracle.OracleModel(cm).gillespie
        soln, _, err = fn([10., 0.], m.target_times, oracle.Rng('philox', seed=3, member=1))
        assert err == 0 and soln[-1, 0] < 10
        # a floor the draining compartment must cross: the assertion fires
        cm = helpers.compiled(Draining(8.), method)
        _, _, err = oracle.OracleModel(cm).tau_leap(
            [10., 0.], m.target_times, oracle.Rng('philox', seed=3, member=1)) \
            if method.startswith('discrete') else oracle.OracleModel(cm).gillespie(
                [10., 0.], m.target_times, oracle.Rng('philox', seed=3, member=1))
        assert err == 1
    # the default checks() stays free for integer counts
    cm = helpers.compiled(_Weird(lambda s: s.params['k']), 'discrete_time_stochastic')
    assert cm.post_check_nodes == [] and '#define PD_HAS_STEP_CHECKS 0' in cm.cuda_source()


def test_explicit_schedule_host_logic():
    tt = list(range(0, 51))
    t_eval, dt, n_sub = compiler.explicit_schedule(tt)
    o_eval, o_dt, o_interval = oracle.explicit_schedule(tt)
    assert np.array_equal(t_eval, o_eval) and np.array_equal(dt, o_dt)
    assert n_sub.sum() == len(dt) == 1024 and set(n_sub) <= {20, 21}
    assert np.array_equal(np.repeat(np.arange(1, 51), n_sub), o_interval)


@pytest.mark.parametrize('name,method', [
    ('sir', 'explicit'), ('seird', 'explicit'), ('tb_vacc', 'explicit'), ('rmit', 'explicit'),
    ('strains', 'discrete_time_stochastic'), ('strains', 'continuous_time_stochastic'),
    ('seird', 'discrete_time_stochastic'), ('rmit', 'continuous_time_stochastic'),
])
def test_generated_cuda_compiles_for_sm_100a(built_library, name, method):
    cm = helpers.compiled(helpers.make_model(name), method)
    outs = ('full', 'final') if method == 'explicit' else ('full', 'stats')
    for out in outs:
        for rng in ((0,) if method == 'explicit' else (0, 1)):
            k = capi.Model(cm.cuda_source(), capi.METHOD_CODES[method], capi.OUTPUT_CODES[out],
                           rng, 0, cm.n_comp, cm.n_param, len(cm.member_param_slots()), cm.n_su,
                           cm.n_flow)
            cubin = k.cubin()
            assert cubin[:4] == b'\x7fELF' and len(cubin) > 4096
            k.close()


def test_nvrtc_errors_surface_with_the_log(built_library):
    with pytest.raises(capi.PopdynError, match='NVRTC'):
        capi.Model('#define PD_C 1\nthis is not CUDA\n', 0, 0, 0, 0, 1, 0, 0, 0, 0)
    cm = helpers.compiled(helpers.make_model('sir'), 'explicit')
    with pytest.raises(capi.PopdynError, match='stochastic'):
        capi.Model(cm.cuda_source(), capi.EXPLICIT, capi.STATS, 0, 0, 3, 2, 0, 0, 2)


def test_old_div_on_an_int_initial_compartment_is_rejected_for_the_ode_method():
    # the reference's first derivative call sees the Python ints of set_compartment and floors
    # old_div(int, int) there, then divides truly: one compiled expression cannot do both
    class Halving(_Weird):
        def __init__(self, init):
            _Weird.__init__(self, lambda s: old_div(s.compartments['a'], 4) * s.params['k'])
            self.set_compartment('a', init)

    from popdynamics_b200.basepop import old_div
    with pytest.raises(TraceError, match='first sub-step'):
        helpers.compiled(Halving(10), 'explicit')
    helpers.compiled(Halving(10.), 'explicit')                   # floats: true division
    cm = helpers.compiled(Halving(10), 'discrete_time_stochastic')  # integer state: floors always
    assert 'floor(' in cm.cuda_source()
    # the shipped examples give int initial values but never divide them
    helpers.compiled(helpers.make_model('strains'), 'explicit')


def test_both_forms_of_the_compensated_sum_equal_cpython_sum():
    """traced_sum emits the correction term of CPython's (>= 3.12) compensated float sum either as
    CPython's own magnitude branch with the operands selected before the arithmetic ('ordered',
    the default) or as Knuth's TwoSum ('twosum', POPDYN_SUM_FORM): the same number whenever
    nothing overflows, and a non-finite compensation -- which the end test discards -- in the
    same cases otherwise."""
    import math
    import sys
    if sys.version_info < (3, 12):
        pytest.skip('builtin sum() compensates from CPython 3.12 on')

    def emitted(xs, form):
        total, comp = xs[0], None
        for x in xs[1:]:
            t = total + x
            if form == 'twosum':
                bb = t - total
                term = (total - (t - bb)) + (x - bb)
            else:                                      # 'ordered': operands selected first
                big_first = abs(total) >= abs(x)
                big, small = (total, x) if big_first else (x, total)
                term = (big - t) + small
            comp = term if comp is None else comp + term
            total = t
        if comp is not None and comp != 0.0 and math.isfinite(comp):
            total = total + comp
        return total

    rng = np.random.default_rng(12)
    specials = [0.0, -0.0, 1e308, -1e308, 1.7976931348623157e308, 5e-324, float('inf'),
                float('-inf'), float('nan')]
    for trial in range(60000):
        n = int(rng.integers(2, 13))
        spread = (1, 8, 40, 300)[trial % 4]
        xs = [float(v) for v in rng.standard_normal(n) * 10.0 ** rng.integers(-spread, spread + 1, n)]
        if trial % 7 == 0:
            xs = [abs(v) for v in xs]                     # rates are non-negative
        if trial % 50 == 0:
            xs[int(rng.integers(0, n))] = specials[int(rng.integers(0, len(specials)))]
        want = sum(xs)
        for form in ('ordered', 'twosum'):
            got = emitted(xs, form)
            assert (want == got and math.copysign(1, want) == math.copysign(1, got)) or \
                (want != want and got != got), (form, xs, want, got)


def test_code_table_rows_are_recognised_in_the_generated_header():
    """The tau-leap kernel looks up the Poisson codes of two kinds of rows instead of computing
    them: count * uniform parameter (PD_ROW_TAB_<row>: one table per parameter slot) and
    count * (count' * uniform parameter) -- a force of infection -- (PD_ROW_TAB2_<row>: one pair
    table per (compartment, parameter) inside the rate, PD_ROW_TAB2B_<row> that compartment).
    A swept (per-member) parameter makes the code member dependent: no table for such a row."""
    import re

    def defines(model, method='discrete_time_stochastic'):
        src = helpers.compiled(model, method).cuda_source()
        return {k: v for k, v in re.findall(r'#define (PD_\w+) (.+)', src)}

    d = defines(helpers.make_model('strains'))
    labels = helpers.compiled(helpers.make_model('strains'), 'discrete_time_stochastic').labels
    assert d['PD_NTAB2'] == '2' and d['PD_LIVE_NROW'] == '10'
    pair_rows = [r for r in range(10) if d['PD_ROW_TAB2_%d' % r] != '-1']
    assert pair_rows == [1, 2]                       # the two infection rows of the flow table
    inside = {labels[int(d['PD_ROW_TAB2B_%d' % r])] for r in pair_rows}
    assert inside == {'infectious_invader', 'infectious_resident'}
    for r in pair_rows:
        assert d['PD_ROW_TAB_%d' % r] == '-1'        # never both kinds
    count_rows = [r for r in range(10) if d['PD_ROW_TAB_%d' % r] != '-1']
    assert count_rows == [3, 4, 5, 6, 7, 8, 9] and d['PD_ROW_UNI_0'] == '0'
    # the SIR model's force of infection is infection_beta * infectious: one pair table
    d = defines(helpers.make_model('sir'))
    assert d['PD_NTAB2'] == '1'
    # ... unless beta is swept: then neither that row nor any other gets a table keyed on it
    swept = helpers.make_model('sir')
    swept.set_param('infection_beta', np.linspace(0.01, 0.05, 7))
    d = defines(swept)
    assert d['PD_NTAB2'] == '0'
    # the ODE methods have no integer counts to key a table on
    d = defines(helpers.make_model('strains'), 'explicit')
    assert d['PD_NTAB2'] == '0'
