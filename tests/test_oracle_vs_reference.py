"""Live comparison with the UNMODIFIED reference (only where /root/reference exists, i.e. in the
authoring container; the GPU box relies on tests/golden/).  Also loads the reference's own example
classes, unmodified, on top of popdynamics_b200.basepop -- the drop-in claim -- and checks that the
flow compiler lowers them to the same trajectories."""
import numpy as np
import pytest

from oracle import oracle, reference
from popdynamics_b200 import basepop, compiler

import helpers

pytestmark = pytest.mark.skipif(not reference.available(), reason='reference tree not present')

SEIR = helpers.models.SEIR_PARAMS


def dropin(script, cls):
    return reference.example_classes(script, against=basepop)[cls]


def run_dropin_explicit(model):
    model.check_flow_transfers()
    cm = compiler.compile_model(model, 'explicit')
    om = oracle.OracleModel(cm)
    return om.explicit(model.get_init_list(), model.target_times,
                       su_table=helpers.scaleup_table(model, cm, model.target_times))[0]


@pytest.mark.parametrize('script,cls,args,times', [
    ('sir_model.py', 'SirModel', (), (0, 50, 1)),
    ('seir_model.py', 'SeirModel', (SEIR['flu'],), (0, 200, 1)),
    ('seir_model.py', 'SeirDemographyModel', (SEIR['measles'],), (0, 730, 1)),
    ('stochastic_strains_model.py', 'StrainsModel', (), (0, 100, 1)),
    ('tb_model.py', 'SimpleTbModel', (['vaccination'],), (1940, 1990, .05)),
    ('rmit_tb_model.py', 'RmitTbModel', (dict(helpers.RMIT, beta=120.),), (0, 200., 1.)),
    ('mix_model.py', 'SirModel', ({'start_population': 100},), (0, 20, 0.2)),
])
def test_reference_example_classes_run_unmodified_on_the_dropin(script, cls, args, times):
    theirs = reference.example_classes(script)[cls](*args)
    theirs.make_times(*times)
    theirs.integrate('explicit')
    ours = dropin(script, cls)(*args)
    ours.make_times(*times)
    assert ours.target_times == theirs.target_times
    soln = run_dropin_explicit(ours)
    assert np.array_equal(soln, theirs.soln_array)


@pytest.mark.parametrize('method,kind', [('discrete_time_stochastic', 'numpy'),
                                         ('continuous_time_stochastic', 'python')])
def test_reference_strains_class_stochastic_on_the_dropin(method, kind):
    reference.seed_all(3)
    theirs = reference.example_classes('stochastic_strains_model.py')['StrainsModel']()
    theirs.make_times(0, 120, 1)
    theirs.integrate(method)
    ours = dropin('stochastic_strains_model.py', 'StrainsModel')()
    ours.make_times(0, 120, 1)
    ours.check_flow_transfers()
    om = oracle.OracleModel(compiler.compile_model(ours, method))
    fn = om.tau_leap if kind == 'numpy' else om.gillespie
    soln, _, err = fn(ours.get_init_list(), ours.target_times, oracle.Rng(kind, 3))
    assert err == 0 and np.array_equal(soln, theirs.soln_array)


def test_tb_model_fails_the_same_way_under_stochastic_methods():
    # KeyError 'program_rate_detect' in the reference, and in the tracer
    theirs = reference.example_classes('tb_model.py')['SimpleTbModel']()
    theirs.make_times(1950, 1951, 1.)
    with pytest.raises(KeyError, match='program_rate_detect'):
        theirs.integrate('discrete_time_stochastic')
    ours = dropin('tb_model.py', 'SimpleTbModel')()
    ours.make_times(1950, 1951, 1.)
    ours.check_flow_transfers()
    with pytest.raises(KeyError, match='program_rate_detect'):
        compiler.compile_model(ours, 'discrete_time_stochastic')


@pytest.mark.parametrize('method,kind,times', [
    ('discrete_time_stochastic', 'numpy', (1960, 1975, 1.)),
    ('continuous_time_stochastic', 'python', (1969, 1971, 1.)),
])
def test_tb_with_scaleups_in_the_stochastic_loops_matches_the_reference(method, kind, times):
    # the declared extension for config 4: same subclass on both sides
    base = reference.example_classes('tb_model.py')['SimpleTbModel']

    class Theirs(base):
        def calculate_vars(self):
            self.calculate_scaleup_vars()
            base.calculate_vars(self)

    reference.seed_all(21)
    theirs = Theirs(['vaccination'])
    theirs.make_times(*times)
    theirs.integrate(method)
    ours = helpers.make_model('tb_scaleup', interventions=['vaccination'])
    ours.make_times(*times)
    soln, _, err = helpers.oracle_stochastic(ours, method, oracle.Rng(kind, 21))
    assert err == 0 and np.array_equal(soln, theirs.soln_array)


def test_hybrid_switching_loop_of_mix_model_matches_the_reference():
    # examples/mix_model.py:107-136, run on the unmodified reference with both global generators
    # seeded; the oracle walks the same loop with ONE MT19937 stream shared by all segments
    Sir = reference.example_classes('mix_model.py')['SirModel']
    reference.seed_all(13)
    theirs = Sir({'start_population': 100})
    theirs.make_times(0, 1, 0.2)
    theirs.integrate(method='continuous_time_stochastic')
    segments, n_explicit = [(0, 1, 0.2)], 0
    day = 1
    while day < 30:
        method = 'explicit' if min(theirs.compartments.values()) > 3 \
            else 'continuous_time_stochastic'
        n_explicit += method == 'explicit'
        theirs.make_times(day, day + 1, 0.2)
        theirs.integrate(method=method, is_continue=True)
        segments.append((day, day + 1, 0.2))
        day += 1
    shared = oracle.Rng('python', 13)
    rows, modes = helpers.oracle_hybrid(
        lambda: helpers.make_model('mix', params={'start_population': 100}), segments, 3,
        lambda k: shared)
    assert 0 < n_explicit < 29 and modes.count('explicit') == n_explicit
    assert np.array_equal(rows, theirs.soln_array)


def test_golden_file_is_current(golden):
    # the committed fixtures are what the reference produces today
    m = reference.example_classes('sir_model.py')['SirModel']()
    m.make_times(0, 50, 1)
    m.integrate()
    assert np.array_equal(m.soln_array, golden['sir_explicit'])
    assert list(golden['sir_explicit_var_labels']) == m.var_labels


def test_injected_uniform_protocol_against_the_reference():
    from oracle import make_golden
    u = make_golden.injected_stream(77, 30000)
    with reference.injected_uniforms(u) as inj:
        theirs = reference.example_classes('sir_model.py')['SirModel'](
            {'population': 3000, 'start_infectious': 30.})
        theirs.make_times(0, 40, 1)
        theirs.integrate('discrete_time_stochastic')
    m = helpers.make_model('sir', params={'population': 3000, 'start_infectious': 30.})
    m.make_times(0, 40, 1)
    rng = oracle.Rng('inject', uniforms=u)
    soln, _, err = helpers.oracle_stochastic(m, 'discrete_time_stochastic', rng)
    assert err == 0 and np.array_equal(soln, theirs.soln_array)
    assert rng.consumed == inj.rng.consumed


@pytest.mark.parametrize('method,kind', [('discrete_time_stochastic', 'numpy'),
                                         ('continuous_time_stochastic', 'python')])
def test_user_defined_checks_in_the_stochastic_loops_fire_like_the_reference(method, kind):
    # basepop.py:789 / :846: checks() after every event / step, on the updated compartments with
    # the vars of the step's start.  Same subclass body on both sides, same MT19937 seed.
    def body(self, error_margin=0.1):
        assert self.compartments['infectious_invader'] <= self.limit
        assert self.vars['population'] + 50 >= sum(self.compartments.values())
        assert self.vars['rate_death'] < 10.

    base = reference.example_classes('stochastic_strains_model.py')['StrainsModel']
    ours_base = dropin('stochastic_strains_model.py', 'StrainsModel')
    for limit, expect_failure in ((10 ** 6, False), (0, True)):
        Theirs = type('Theirs', (base,), {'checks': body, 'limit': limit})
        Ours = type('Ours', (ours_base,), {'checks': body, 'limit': limit})
        reference.seed_all(4)
        theirs = Theirs()
        theirs.make_times(0, 150, 1)
        failed = False
        try:
            theirs.integrate(method)
        except AssertionError:
            failed = True
        assert failed == expect_failure
        ours = Ours()
        ours.make_times(0, 150, 1)
        ours.check_flow_transfers()
        cm = compiler.compile_model(ours, method)
        assert len(cm.post_check_nodes) == 3
        om = oracle.OracleModel(cm)
        fn = om.tau_leap if kind == 'numpy' else om.gillespie
        soln, _, err = fn(ours.get_init_list(), ours.target_times, oracle.Rng(kind, 4))
        assert (err == 1) == expect_failure
        if not expect_failure:
            assert np.array_equal(soln, theirs.soln_array)


def test_host_derivative_closure_and_event_list_equal_the_references():
    # make_derivate_fn (basepop.py:613-626) and calculate_events (:690-735) exist on the drop-in as
    # host methods; the reference's unmodified example classes run on both base classes
    for script, cls, args, y, t in (
            ('tb_model.py', 'SimpleTbModel', ([],), [9e5, 100., 2000., 50., 20., 10.], 1985.3),
            ('seir_model.py', 'SeirDemographyModel', (SEIR['measles'],),
             [9e5, 20., 30., 9e4], 12.5),
            ('rmit_tb_model.py', 'RmitTbModel', (dict(helpers.RMIT, beta=80.),),
             [.6, .01, .05, .3, .001, .03], 3.0)):
        theirs = reference.example_classes(script)[cls](*args)
        ours = dropin(script, cls)(*args)
        theirs.check_flow_transfers()
        ours.check_flow_transfers()
        f_theirs, f_ours = theirs.make_derivate_fn(), ours.make_derivate_fn()
        assert f_ours(list(y), t) == f_theirs(list(y), t)
        assert list(ours.vars.items()) == list(theirs.vars.items())      # same keys, same order
    theirs = reference.example_classes('stochastic_strains_model.py')['StrainsModel']()
    ours = dropin('stochastic_strains_model.py', 'StrainsModel')()
    for m in (theirs, ours):
        m.check_flow_transfers()
        m.compartments = {k: int(v) for k, v in m.init_compartments.items()}
        m.compartments['recovered_resident'] = 7
        m.calculate_vars()
        m.calculate_events()
    assert ours.events == theirs.events and len(ours.events) == 9
    assert ours.vars['rate_death'] == theirs.vars['rate_death']
