"""Drop-in behaviour of popdynamics_b200.basepop.BaseModel on the host side: setter asserts,
time grids, flow de-duplication, connectivity check, curve builders, clone -- the error paths the
reference exercises through its asserts."""
import math

import numpy as np
import pytest

from popdynamics_b200 import basepop
from popdynamics_b200.basepop import BaseModel

import helpers


def test_make_times_accumulates_and_keeps_ints():
    m = BaseModel()
    m.make_times(0, 50, 1)
    assert m.target_times == list(range(51)) and all(type(t) is int for t in m.target_times)
    m.make_times(1900, 2050, .05)
    assert len(m.target_times) == 3001 and m.target_times[-1] == 2049.9999999998727
    m.make_times_with_n_step(0., 10., 4)
    assert m.target_times == [0., 2.5, 5., 7.5, 10.]


@pytest.mark.parametrize('args', [('0', 1, 1), (0, '1', 1), (0, 1, None), (5, 1, 1)])
def test_make_times_asserts(args):
    with pytest.raises(AssertionError):
        BaseModel().make_times(*args)
    with pytest.raises(AssertionError):          # numpy scalars are rejected like the reference
        BaseModel().make_times(np.float64(0.), 1., 1.)


def test_setter_type_asserts():
    m = BaseModel()
    with pytest.raises(AssertionError):
        m.set_compartment(3, 1.)
    with pytest.raises(AssertionError):
        m.set_compartment('a', '1')
    with pytest.raises(AssertionError):
        m.set_compartment('a', -1.)
    with pytest.raises(AssertionError):
        m.set_param(1, 1.)
    with pytest.raises(AssertionError):
        m.set_param('p', 'x')
    with pytest.raises(AssertionError):
        m.set_scaleup_fn(1, lambda t: t)
    with pytest.raises(AssertionError):
        m.set_var_transfer_rate_flow('a', 'b', 3)
    with pytest.raises(AssertionError):
        m.set_var_entry_rate_flow(1, 'v')
    with pytest.raises(KeyError):
        m.set_fixed_transfer_rate_flow('a', 'b', 'missing_param')
    with pytest.raises(KeyError):
        m.set_infection_death_rate_flow('a', 'missing_param')
    with pytest.raises(KeyError):
        m.set_background_death_rate('missing_param')


def test_per_member_arrays_are_an_accepted_extension():
    m = BaseModel()
    m.set_param('beta', np.linspace(1., 2., 8))
    m.set_compartment('s', np.full(8, 10.))
    with pytest.raises(AssertionError):
        m.set_compartment('s', np.array([1., -1.]))
    with pytest.raises(AssertionError):
        m.set_param('beta', np.zeros((2, 2)))


def test_flow_dedup_replaces_in_place():
    m = BaseModel()
    m.set_param('a', 1.)
    m.set_param('b', 2.)
    m.set_fixed_transfer_rate_flow('x', 'y', 'a')
    m.set_fixed_transfer_rate_flow('y', 'z', 'a')
    m.set_fixed_transfer_rate_flow('x', 'y', 'b')
    assert m.fixed_transfer_rate_flows == [('x', 'y', 2.), ('y', 'z', 1.)]
    m.set_var_entry_rate_flow('x', 'v1')
    m.set_var_entry_rate_flow('x', 'v2')
    assert m.var_entry_rate_flow == [('x', 'v2')]
    m.set_infection_death_rate_flow('x', 'a')
    m.set_infection_death_rate_flow('x', 'b')
    assert m.infection_death_rate_flows == [('x', 2.)]
    assert basepop.label_intersects_tags('latent_early', ['active', 'latent'])
    assert not basepop.label_intersects_tags('susceptible', ['active', 'latent'])


def test_background_death_asserts_previous_value():
    m = BaseModel()
    m.set_param('mu', 1)                 # an int becomes the stored rate ...
    m.set_background_death_rate('mu')
    with pytest.raises(AssertionError):  # ... and the NEXT call asserts it is a float
        m.set_background_death_rate('mu')


def test_connectivity_assert():
    m = helpers.make_model('sir')
    m.set_compartment('orphan', 1.)
    with pytest.raises(AssertionError, match='orphan'):
        m.check_flow_transfers()


def test_integrate_preconditions_and_rejections():
    m = helpers.make_model('sir')
    with pytest.raises(AssertionError, match='times'):
        m.integrate()
    m.make_times(0, 5, 1)
    with pytest.raises(ValueError):
        m.integrate('runge_kutta')
    # a foreign derivative cannot be honoured by the device integrator: rejected loudly
    with pytest.raises(TypeError, match='make_derivate_fn'):
        m.integrate_explicit(m.get_init_list(), lambda y, t: [0.] * 3)


def test_host_derivative_closure_and_events_match_the_reference_semantics():
    # make_derivate_fn (basepop.py:613-626): vars cleared, scale-ups, hook, flows, checks
    m = helpers.make_model('tb')
    m.check_flow_transfers()
    f = m.make_derivate_fn()
    y = [1e6, 10., 20., 30., 4., 5.]
    flows = f(y, 1975.)
    assert m.time == 1975. and len(flows) == len(m.labels)
    assert list(m.vars)[0] == 'program_rate_detect'           # scale-ups first (basepop.py:608)
    assert 'rate_death' in m.vars and 'rate_infection_death' in m.vars
    total = sum(y)
    births = m.params['demo_rate_birth'] * total
    deaths = m.vars['rate_death'] + m.vars['rate_infection_death']
    assert abs(sum(flows) - (births - deaths)) < 1e-6
    with pytest.raises(AssertionError):                       # checks() on a negative state
        f([-1.] + y[1:], 1975.)
    # calculate_events (basepop.py:690-735): entry always, others only when positive
    s = helpers.make_model('strains')
    s.check_flow_transfers()
    s.compartments = {k: int(v) for k, v in s.init_compartments.items()}
    s.calculate_vars()
    s.calculate_events()
    assert s.events[0] == (None, 'susceptible', 2)
    kinds = [(e[0], e[1]) for e in s.events]
    assert ('susceptible', 'infectious_invader') in kinds
    assert ('recovered_resident', None) not in kinds          # zero count: no event
    assert all(e[2] > 0 for e in s.events[1:])
    assert s.vars['rate_death'] == sum(e[2] for e in s.events if e[1] is None)
    s.compartments['infectious_invader'] = 0
    s.calculate_vars()
    s.calculate_events()
    assert ('susceptible', 'infectious_invader') not in [(e[0], e[1]) for e in s.events]


def test_sigmoid_curves():
    curve = basepop.make_sigmoidal_curve(y_high=2, y_low=0, x_start=1950, x_inflect=1970,
                                         multiplier=4)
    b = 4. * (4 * 0.5 * 2 / 20.) / 2
    assert curve(1970) == 2 / (1. + math.exp(0.)) + 0
    assert curve(1900) == 0                      # arg > 10 early return
    assert curve(1980) == 2 / (1. + math.exp(b * (1970 - 1980))) + 0
    assert basepop.make_sigmoidal_curve(y_low=3, y_high=3)(17.) == 3
    assert basepop.make_constant_function(5.)(1.) == 5.
    two = basepop.make_two_step_curve(0., 1., 3., 10., 20., 30.)
    assert two(5.) == 0. and two(35.) == 3.
    assert 0. < two(15.) < 1. < two(25.) < 3.


def test_curves_match_reference_when_present():
    from oracle import reference
    if not reference.available():
        pytest.skip('reference tree not present')
    ref = reference.basepop()
    for kw in (dict(y_high=2, y_low=0, x_start=1950, x_inflect=1970, multiplier=4),
               dict(y_high=.95, y_low=0, x_start=1921, x_inflect=2006, multiplier=3)):
        ours, theirs = basepop.make_sigmoidal_curve(**kw), ref.make_sigmoidal_curve(**kw)
        for x in np.linspace(1900, 2050, 301):
            assert ours(float(x)) == theirs(float(x))
    ours = basepop.make_two_step_curve(0., 1., 3., 10., 20., 30.)
    theirs = ref.make_two_step_curve(0., 1., 3., 10., 20., 30.)
    for x in np.linspace(0, 40, 161):
        assert ours(float(x)) == theirs(float(x))


def test_pick_event_host_version(monkeypatch, golden):
    import random
    for row, u, want in list(zip(golden['pick_intervals'], golden['pick_uniform'],
                                 golden['pick_event']))[:50]:
        monkeypatch.setattr(random, 'random', lambda u=u: float(u))
        assert basepop.pick_event(list(row[row >= 0])) == want


def test_old_div():
    assert basepop.old_div(7, 2) == 3 and basepop.old_div(7., 2) == 3.5
    assert basepop.old_div(np.int64(7), 2) == 3


def test_clone_copies_what_the_reference_copies():
    m = helpers.make_model('sir')
    m.make_times(0, 5, 1)
    m.check_flow_transfers()
    c = m.clone()
    assert c.params == m.params and c.fixed_transfer_rate_flows == m.fixed_transfer_rate_flows
    assert c.labels != m.labels or not m.labels or c.labels == ['susceptible', 'infectious', 'immune']
    assert c.target_times == []                  # not copied, as in the reference


def test_host_diagnostics_pass(golden):
    # calculate_diagnostics is host Python over a stored soln_array (basepop.py:911-960)
    m = helpers.make_model('sir')
    m.make_times(0, 50, 1)
    m.check_flow_transfers()
    m.soln_array = golden['sir_explicit'].copy()
    m.times = list(m.target_times)
    m.calculate_diagnostics()
    assert m.var_labels == list(golden['sir_explicit_var_labels'])
    assert np.array_equal(m.var_array, golden['sir_explicit_var_array'])
    assert np.array_equal(m.flow_array, golden['sir_explicit_flow_array'])
    assert np.array_equal(np.array(m.fraction_soln['immune']),
                          golden['sir_explicit_fraction_immune'])
    assert m.check_converged_compartment_fraction('immune', 5, 1e-2)
    m.load_state(10)
    assert m.compartments['infectious'] == golden['sir_explicit'][10, 1]
