# -*- coding: utf-8 -*-
"""
Drop-in replacement for ``basepop`` (popdynamics) whose integrators run on a B200.

The model-definition API mirrors the reference class ``basepop.BaseModel``
(/root/reference/basepop.py:184-558): same method names, argument meaning, attribute names,
assertion behaviour and result attributes.  What changed is everything behind
``integrate(method)`` (/root/reference/basepop.py:854-903): the three pure-Python loops
(``integrate_explicit`` :658-688, ``integrate_discrete_time_stochastic`` :797-852,
``integrate_continuous_time_stochastic`` :737-795) are replaced by

    flow compiler + calculate_vars tracer  (popdynamics_b200/compiler.py)
      -> C-ABI library libpopdyn_b200.so   (popdynamics_b200/csrc, include/popdyn_b200.h)
      -> NVRTC-specialised sm_100a kernels (popdynamics_b200/csrc/kernels/*.cuh)

There is no CPU fallback: without the CUDA extension or without a GPU ``integrate`` raises.
``integrate('scipy')`` runs an adaptive Runge-Kutta integrator on the device in place of
``scipy.integrate.odeint`` (basepop.py:882-886).

Additive API (not in the reference): ``set_param`` / ``set_compartment`` also accept 1-D
``numpy.ndarray`` values (one entry per ensemble member), and ``integrate_ensemble`` runs
millions of members in one launch (see popdynamics_b200/ensemble.py).
"""

from __future__ import print_function, division

import copy
import glob
import math
import os
import platform
import random

import numpy

__all__ = [
    'BaseModel', 'add_unique_tuple_to_list', 'label_intersects_tags', 'make_sigmoidal_curve',
    'make_constant_function', 'make_two_step_curve', 'pick_event', 'ensure_out_dir',
    'open_pngs_in_dir', 'old_div',
]


def old_div(a, b):
    """``past.utils.old_div`` (third-party package ``future``, used at basepop.py:10): floor
    division for two integers, true division otherwise.  Works on traced values as well."""
    from .compiler import Sym, sym_old_div
    if isinstance(a, Sym) or isinstance(b, Sym):
        return sym_old_div(a, b)
    import numbers
    if isinstance(a, numbers.Integral) and isinstance(b, numbers.Integral):
        return a // b
    return a / b


# ---------------------------------------------------------------------------------------------
# small host-side helpers (reference basepop.py:30-80); no numerics on the hot path
# ---------------------------------------------------------------------------------------------

def ensure_out_dir(out_dir):
    """Create ``out_dir`` when missing (basepop.py:30-35)."""
    os.makedirs(out_dir, exist_ok=True)


def open_pngs_in_dir(out_dir):
    """Open the PNG files of a directory through the OS viewer (basepop.py:38-47)."""
    pngs = sorted(glob.glob(os.path.join(out_dir, '*png')))
    opener = {'Windows': 'start ', 'Darwin': 'open '}
    for system_name, command in opener.items():
        if system_name in platform.system() and pngs:
            os.system(command + ' '.join(pngs))


def add_unique_tuple_to_list(a_list, a_tuple):
    """Insert ``a_tuple``; a tuple whose fields before the last equal an existing entry's
    replaces that entry in place (basepop.py:53-63).  This is the de-duplication rule of every
    flow list: one flow per (from, to), one entry flow / infection death per compartment."""
    key = a_tuple[:-1]
    for position, present in enumerate(a_list):
        if present[:-1] == key:
            a_list[position] = a_tuple
            return
    a_list.append(a_tuple)


def label_intersects_tags(label, tags):
    """True when any of ``tags`` is a substring of ``label`` (basepop.py:66-80)."""
    return any(tag in label for tag in tags)


def make_sigmoidal_curve(y_low=0, y_high=1., x_start=0, x_inflect=0.5, multiplier=1.):
    """Sigmoid scale-up curve builder, arithmetic as basepop.py:86-124 (same operation order,
    including the ``arg > 10`` early return)."""
    amplitude = y_high - y_low
    if amplitude == 0:
        return make_constant_function(y_low)
    x_delta = x_inflect - x_start
    slope_at_inflection = multiplier * 0.5 * amplitude / x_delta
    b = 4. * slope_at_inflection / amplitude

    def curve(x):
        arg = b * (x_inflect - x)
        if arg > 10.:
            return y_low
        return old_div(amplitude, (1. + math.exp(arg))) + y_low

    curve.pd_sigmoid = dict(amplitude=amplitude, b=b, x_inflect=x_inflect, y_low=y_low)
    return curve


def make_constant_function(value):
    """basepop.py:127-131."""
    def curve(x):
        return value

    curve.pd_constant = value
    return curve


def make_two_step_curve(y_low, y_med, y_high, x_start, x_med, x_end):
    """Two chained sigmoids with hard limits outside [x_start, x_end] (basepop.py:134-158)."""
    first = make_sigmoidal_curve(y_high=y_med, y_low=y_low, x_start=x_start,
                                 x_inflect=(x_med - x_start) * 0.5 + x_start, multiplier=4)
    second = make_sigmoidal_curve(y_high=y_high, y_low=y_med, x_start=x_med,
                                  x_inflect=(x_end - x_med) * 0.5 + x_med, multiplier=4)

    def curve(x):
        if x < x_start:
            return y_low
        if x < x_med:
            return first(x)
        if x < x_end:
            return second(x)
        return y_high

    return curve


def pick_event(event_intervals):
    """Host restatement of the Gillespie event selection (basepop.py:161-178): sequential
    cumulative sum, one ``random.random()``, strict ``>`` linear search.  The device version
    lives in csrc/kernels/pd_gillespie.cuh."""
    running, cumulative = 0.0, []
    for interval in event_intervals:
        running += interval
        cumulative.append(running)
    point = random.random() * running
    i_event, i_last = 0, len(event_intervals) - 1
    while point > cumulative[i_event] and i_event < i_last:
        i_event += 1
    return i_event


def _is_number(val):
    return type(val) is float or type(val) is int


def _is_member_array(val):
    return isinstance(val, numpy.ndarray) and val.ndim == 1 and val.size > 0


class BaseModel(object):
    """
    Compartmental model whose differential equations are assembled from individual
    compartment-to-compartment connections; API of the reference ``basepop.BaseModel``
    (basepop.py:184-1131).

    Execution per step (now on the device): vars from params+compartments (the traced
    ``calculate_vars`` hook) -> transfers from vars/scale-ups/params -> flows -> update.
    """

    #: sub-step of the explicit Euler integrator (basepop.py:658, never overridden by integrate)
    explicit_min_dt = 0.05

    def __init__(self):
        # attribute names and defaults as basepop.py:237-319
        self.labels = []
        self.init_compartments = {}
        self.params = {}
        self.target_times = []
        self.time = 0
        self.times = []
        self.start_time = 0
        self.scaleup_fns = {}
        self.compartments = {}
        self.vars = {}
        self.flows = {}
        self.var_entry_rate_flow = []          # (label, var_label)
        self.fixed_transfer_rate_flows = []    # (from_label, to_label, rate value)
        self.var_transfer_rate_flows = []      # (from_label, to_label, var_label)
        self.infection_death_rate_flows = []   # (label, rate value)
        self.background_death_rate = 0.
        self.soln_array = None
        self.flow_array = None
        self.var_array = None
        self.var_labels = None
        # --- additions: which param label stands behind a resolved rate (for member sweeps) ---
        self._fixed_rate_params = {}           # (from, to) -> param label
        self._infection_death_params = {}      # label -> param label
        self._background_death_param = None
        #: seed of the native Philox streams; None draws one from ``random`` at integrate time
        self.seed = None
        #: GPU ordinal used by integrate()
        self.device = 0
        #: filled by integrate(): device statistics of the last launch
        self.last_run_info = None

    # ------------------------------------------------------------------ bookkeeping
    def clear_solns(self):
        """basepop.py:321-325."""
        self.soln_array = None
        self.flow_array = None
        self.var_labels = None
        self.var_array = None

    def clone(self):
        """Copy params, flows and stored arrays into a new instance (basepop.py:327-353).  As in
        the reference, labels / init_compartments / target_times are NOT copied."""
        duplicate = type(self)()
        for name in ('params', 'compartments', 'scaleup_fns', 'fixed_transfer_rate_flows',
                     'var_transfer_rate_flows', 'infection_death_rate_flows',
                     'var_entry_rate_flow', 'background_death_rate', 'var_labels',
                     '_fixed_rate_params', '_infection_death_params',
                     '_background_death_param'):
            setattr(duplicate, name, copy.deepcopy(getattr(self, name)))
        duplicate.soln_array = numpy.copy(self.soln_array)
        duplicate.var_array = numpy.copy(self.var_array)
        duplicate.flow_array = numpy.copy(self.flow_array)
        return duplicate

    # ------------------------------------------------------------------ time grid
    def make_times(self, start, end, delta):
        """Accumulating time grid (basepop.py:355-376): ``step += delta`` in the caller's number
        type, so the grid is NOT ``start + i * delta`` and ints stay ints."""
        assert type(start) is float or type(start) is int, 'Start time not specified with float'
        assert type(end) is float or type(end) is int, 'End time not specified with a number'
        assert type(delta) is float or type(delta) is int, \
            'Time increment not specified with a number'
        assert end >= start, 'End time is before start time'
        self.target_times = []
        step = start
        while step <= end:
            self.target_times.append(step)
            step += delta

    def make_times_with_n_step(self, start, end, n):
        """basepop.py:378-393: as make_times with ``delta = (end - start) / float(n)``."""
        self.target_times = []
        step = start
        delta = old_div((end - start), float(n))
        while step <= end:
            self.target_times.append(step)
            step += delta

    # ------------------------------------------------------------------ definition API
    def set_compartment(self, label, init_val=0.):
        """basepop.py:395-407.  Extension: a 1-D float array gives one initial value per member."""
        assert type(label) is str, 'Compartment label for initial setting not string'
        if _is_member_array(init_val):
            assert (init_val >= 0.).all(), 'Start with negative compartment not permitted'
        else:
            assert _is_number(init_val), 'Value to start %s compartment from not numeric' % label
            assert init_val >= 0., 'Start with negative compartment not permitted'
        if label not in self.labels:
            self.labels.append(label)
        self.init_compartments[label] = init_val

    def set_param(self, label, val):
        """basepop.py:409-418.  Extension: a 1-D float array sweeps the parameter over members."""
        assert type(label) is str, 'Parameter name "%s" is not string' % label
        assert _is_number(val) or _is_member_array(val), \
            'Fixed parameter value is not numeric for %s' % label
        self.params[label] = val

    def convert_list_to_compartments(self, vec):
        """basepop.py:420-427."""
        return {label: vec[i] for i, label in enumerate(self.labels)}

    def convert_compartments_to_list(self, compartments):
        """basepop.py:429-437."""
        return [compartments[label] for label in self.labels]

    def get_init_list(self):
        """basepop.py:439-447."""
        return self.convert_compartments_to_list(self.init_compartments)

    def set_scaleup_fn(self, label, fn):
        """basepop.py:449-459."""
        assert type(label) is str, 'Name of scale-up function is not string'
        self.scaleup_fns[label] = fn

    def set_background_death_rate(self, param_label='demo_rate_death'):
        """basepop.py:461-471 (the asserts look at the PREVIOUS value, as in the reference)."""
        previous = self.background_death_rate
        assert type(previous) is float or _is_member_array(previous), \
            'Background death rate is not float'
        assert numpy.all(numpy.asarray(previous) >= 0.), 'Background death rate is negative'
        self.background_death_rate = self.params[param_label]
        self._background_death_param = param_label

    def set_infection_death_rate_flow(self, label, param_label):
        """basepop.py:473-487."""
        assert type(label) is str, 'Compartment label not string when setting infection death rate'
        assert type(param_label) is str, \
            'Parameter label not string when setting infection death rate'
        add_unique_tuple_to_list(self.infection_death_rate_flows,
                                 (label, self.params[param_label]))
        self._infection_death_params[label] = param_label

    def set_fixed_transfer_rate_flow(self, from_label, to_label, param_label):
        """basepop.py:489-506."""
        assert type(from_label) is str, \
            'Origin compartment label not string for setting fixed transfer rate'
        assert type(to_label) is str, \
            'Destination compartment label not string for setting fixed transfer rate'
        add_unique_tuple_to_list(self.fixed_transfer_rate_flows,
                                 (from_label, to_label, self.params[param_label]))
        self._fixed_rate_params[(from_label, to_label)] = param_label

    def set_var_transfer_rate_flow(self, from_label, to_label, var_label):
        """basepop.py:508-529."""
        assert type(from_label) is str, \
            'Origin compartment label not string for setting variable transfer rate'
        assert type(to_label) is str, \
            'Destination compartment label not string for setting variable transfer rate'
        assert type(var_label) is str, 'Function label not string for setting fixed transfer rate'
        add_unique_tuple_to_list(self.var_transfer_rate_flows, (from_label, to_label, var_label))

    def set_var_entry_rate_flow(self, label, var_label):
        """basepop.py:531-542."""
        assert type(label) is str, \
            'Destination compartment label not string for setting variable transfer rate'
        assert type(var_label) is str, 'Function label not string for setting variable entry rate'
        add_unique_tuple_to_list(self.var_entry_rate_flow, (label, var_label))

    # ------------------------------------------------------------------ hooks
    def calculate_scaleup_vars(self):
        """basepop.py:544-551."""
        for label, fn in self.scaleup_fns.items():
            self.vars[label] = fn(self.time)

    def calculate_vars(self):
        """Hook (basepop.py:553-558): vars that depend on compartments and params only.  Traced
        ONCE with symbolic proxies by the flow compiler, then executed on the device."""
        pass

    def set_flows(self):
        """Hook (basepop.py:628-636): register the entry / var / fixed flows."""
        pass

    def calculate_diagnostic_vars(self):
        """Hook (basepop.py:905-909)."""
        pass

    def checks(self, error_margin=0.1):
        """Default invariant (basepop.py:1016-1025): every compartment >= 0.  On the device the
        same predicate raises an error flag that integrate() turns into AssertionError."""
        for label in self.labels:
            assert self.compartments[label] >= 0.

    # ------------------------------------------------------------------ host evaluators
    def calculate_flows(self):
        """Net flow per compartment in the canonical order entry -> var -> fixed -> background
        death -> infection death (basepop.py:560-598).  Host version, used by the diagnostics
        pass; the integrators use the compiled form of the same table."""
        flows = self.flows
        comps = self.compartments
        for label in self.labels:
            flows[label] = 0.
        for label, var_label in self.var_entry_rate_flow:
            flows[label] += self.vars[var_label]
        for from_label, to_label, var_label in self.var_transfer_rate_flows:
            val = comps[from_label] * self.vars[var_label]
            flows[from_label] -= val
            flows[to_label] += val
        for from_label, to_label, rate in self.fixed_transfer_rate_flows:
            val = comps[from_label] * rate
            flows[from_label] -= val
            flows[to_label] += val
        self.vars['rate_death'] = 0.
        for label in self.labels:
            val = comps[label] * self.background_death_rate
            flows[label] -= val
            self.vars['rate_death'] += val
        self.vars['rate_infection_death'] = 0.
        for label, rate in self.infection_death_rate_flows:
            val = comps[label] * rate
            flows[label] -= val
            self.vars['rate_infection_death'] += val

    def prepare_vars_and_flows(self):
        """basepop.py:600-611."""
        self.vars.clear()
        self.calculate_scaleup_vars()
        self.calculate_vars()
        self.calculate_flows()

    def check_flow_transfers(self):
        """Re-register the flows and assert every compartment is connected (basepop.py:638-656)."""
        self.set_flows()
        connected = set(flow[0] for flow in self.var_entry_rate_flow)
        for flow_list in (self.var_transfer_rate_flows, self.fixed_transfer_rate_flows):
            for flow in flow_list:
                connected.add(flow[0])
                connected.add(flow[1])
        for label in self.labels:
            assert label in connected, \
                'Compartment "%s" doesn\'t have any entry or transfer flows' % label

    # ------------------------------------------------------------------ host restatements
    def make_derivate_fn(self):
        """The derivative closure ``f(y, t)`` of basepop.py:613-626, on the HOST: time and
        compartments are loaded, ``prepare_vars_and_flows`` runs, ``checks`` asserts, the flow
        vector comes back in label order.  The device integrators use the compiled form of the
        same chain (``pd_derivative``); this closure is what scripts pass to their own solvers
        and what the tests feed to ``scipy.integrate.odeint``."""
        def derivative_fn(y, t):
            self.time = t
            self.compartments = self.convert_list_to_compartments(y)
            self.prepare_vars_and_flows()
            flow_vector = self.convert_compartments_to_list(self.flows)
            self.checks()
            return flow_vector

        derivative_fn.pd_model = self
        return derivative_fn

    def calculate_events(self):
        """``self.events``: the list of ``(from_label, to_label, rate)`` the stochastic loops draw
        from, in flow-table order (basepop.py:690-735).  Entry events are always present, the
        others only with a positive rate; the two death totals are side outputs.  Host version of
        what the kernels evaluate per step (``PD_FOR_EACH_LIVE_ROW``)."""
        events = []
        for label, var_label in self.var_entry_rate_flow:
            events.append((None, label, self.vars[var_label]))
        transfers = [(f, t, self.vars[v]) for f, t, v in self.var_transfer_rate_flows]
        transfers += list(self.fixed_transfer_rate_flows)
        for from_label, to_label, rate in transfers:
            val = self.compartments[from_label] * rate
            if val > 0:
                events.append((from_label, to_label, val))
        deaths = [('rate_death', [(label, self.background_death_rate) for label in self.labels]),
                  ('rate_infection_death', list(self.infection_death_rate_flows))]
        for total_label, rows in deaths:
            self.vars[total_label] = 0.
            for label, rate in rows:
                val = self.compartments[label] * rate
                if val > 0:
                    self.vars[total_label] += val
                    events.append((label, None, val))
        self.events = events

    # ------------------------------------------------------------------ integrate (device)
    def _run_on_device(self, y, method, is_continue, **options):
        from . import ensemble
        info = ensemble.integrate_single(self, y, method, is_continue, **options)
        self.last_run_info = info
        self.soln_array = info['soln']
        return info

    def integrate_explicit(self, y, derivative=None, min_dt=0.05):
        """basepop.py:658-688 on the device: explicit Euler from ``y`` over ``target_times`` with
        sub-step ``min_dt``, filling ``soln_array``.  ``derivative`` is accepted for signature
        compatibility and must be a closure of THIS model made by ``make_derivate_fn`` (or None):
        the device integrates the model's compiled derivative, so a foreign function cannot be
        honoured and is rejected instead of being silently ignored."""
        if derivative is not None and getattr(derivative, 'pd_model', None) is not self:
            raise TypeError('integrate_explicit runs the compiled derivative of this model on the '
                            'device; pass make_derivate_fn() of the same model (or None)')
        return self._run_on_device(y, 'explicit', getattr(self, '_continuing', False),
                                   min_dt=min_dt)

    def integrate_continuous_time_stochastic(self, y):
        """basepop.py:737-795 on the device (Gillespie direct method)."""
        return self._run_on_device(y, 'continuous_time_stochastic',
                                   getattr(self, '_continuing', False))

    def integrate_discrete_time_stochastic(self, y):
        """basepop.py:797-852 on the device (clamped Poisson tau-leap)."""
        return self._run_on_device(y, 'discrete_time_stochastic',
                                   getattr(self, '_continuing', False))

    def integrate_adaptive(self, y, rtol=1.49012e-8, atol=1.49012e-8):
        """What ``integrate('scipy')`` runs: the adaptive Dormand-Prince 5(4) integrator on the
        device, in place of ``odeint(derivative, y, self.target_times)`` (basepop.py:886)."""
        return self._run_on_device(y, 'scipy', getattr(self, '_continuing', False),
                                   rtol=rtol, atol=atol)

    def integrate(self, method='explicit', is_continue=False):
        """
        Drop-in for basepop.py:854-903.  ``method`` is one of ``'explicit'``, ``'scipy'``,
        ``'continuous_time_stochastic'``, ``'discrete_time_stochastic'``.  Runs one ensemble
        member on the GPU through the per-method entry points above (so a subclass that overrides
        one of them is honoured) and fills ``soln_array``, ``times`` and the diagnostics exactly
        like the reference.  ``'scipy'`` does not need scipy: LSODA is replaced by an adaptive
        Runge-Kutta pair with the same default tolerances (agreement within 1e-5 relative, not
        bit for bit).
        """
        assert bool(self.target_times), 'Haven\'t set times yet'
        if method not in ('explicit', 'scipy', 'continuous_time_stochastic',
                          'discrete_time_stochastic'):
            raise ValueError('unknown integration method %r' % (method,))

        if not is_continue:
            self.clear_solns()
            self.check_flow_transfers()
            y = self.get_init_list()
        else:
            self.save_soln_array = self.soln_array
            y = self.convert_compartments_to_list(self.compartments)

        self._continuing = bool(is_continue)
        try:
            if method == 'explicit':
                info = self.integrate_explicit(y, self.make_derivate_fn(), self.explicit_min_dt)
            elif method == 'scipy':
                info = self.integrate_adaptive(y)
            elif method == 'continuous_time_stochastic':
                info = self.integrate_continuous_time_stochastic(y)
            else:
                info = self.integrate_discrete_time_stochastic(y)
        finally:
            self._continuing = False
        info = info if isinstance(info, dict) else (self.last_run_info or {})

        if is_continue:
            self.soln_array = numpy.concatenate((self.save_soln_array, self.soln_array[1:]), 0)
            self.times.extend(self.target_times[1:])
        else:
            self.times = copy.deepcopy(self.target_times)

        if method in ('explicit', 'scipy'):
            # state the reference leaves in self.vars after its last derivative call: key order
            # of prepare_vars_and_flows and scale-ups frozen at the last evaluation time
            # (basepop.py:606-611); for 'scipy' that time is LSODA's
            # business in the reference and the last target time here
            last_eval_time = info.get('last_eval_time')
            if last_eval_time is not None:
                self.time = last_eval_time
                self.compartments = self.convert_list_to_compartments(
                    [v for v in self.soln_array[-1]])
                self.prepare_vars_and_flows()
        else:
            self.compartments = self.convert_list_to_compartments(
                [int(v) for v in self.soln_array[-1]])
        self.calculate_diagnostics()

    def integrate_ensemble(self, method='explicit', **kwargs):
        """Additive API: run many members at once; see ``ensemble.integrate_ensemble``."""
        from . import ensemble
        return ensemble.integrate_ensemble(self, method, **kwargs)

    # ------------------------------------------------------------------ diagnostics (host)
    def calculate_diagnostics(self):
        """Second pass over the stored time points (basepop.py:911-960) re-running the hooks to
        fill ``var_array``, ``flow_array``, ``population_soln`` and ``fraction_soln``.  Host
        Python for a single model (SURVEY.md section 8f row N1 is the batched device version)."""
        self.population_soln = {}
        for label in self.labels:
            self.population_soln[label] = self.get_compartment_soln(label)
        self.var_labels = None
        n_time = len(self.times)
        for i in range(n_time):
            self.time = self.times[i]
            for label in self.labels:
                self.compartments[label] = self.population_soln[label][i]
            self.calculate_vars()
            self.calculate_flows()
            self.calculate_diagnostic_vars()
            if self.var_labels is None:
                self.var_labels = list(self.vars.keys())
                self.var_array = numpy.zeros((n_time, len(self.var_labels)))
                self.flow_array = numpy.zeros((n_time, len(self.labels)))
            for i_label, label in enumerate(self.var_labels):
                self.var_array[i, i_label] = self.vars[label]
            for i_label, label in enumerate(self.labels):
                self.flow_array[i, i_label] = self.flows[label]
        self.fraction_soln = {}
        population = self.get_var_soln('population')
        for label in self.labels:
            self.fraction_soln[label] = [
                old_div(v, t) for v, t in zip(self.population_soln[label], population)]

    def get_compartment_soln(self, label):
        """basepop.py:962-974."""
        assert self.soln_array is not None, 'calculate_diagnostics has not been run'
        return self.soln_array[:, self.labels.index(label)]

    def get_var_soln(self, label):
        """basepop.py:976-988."""
        assert self.var_array is not None, 'calculate_diagnostics has not been run'
        return self.var_array[:, self.var_labels.index(label)]

    def get_flow_soln(self, label):
        """basepop.py:990-1000."""
        assert self.flow_array is not None, 'calculate_diagnostics has not been run'
        return self.flow_array[:, self.labels.index(label)]

    def load_state(self, i_time):
        """basepop.py:1002-1014."""
        self.time = self.target_times[i_time]
        for i_label, label in enumerate(self.labels):
            self.compartments[label] = self.soln_array[i_time, i_label]
        self.calculate_vars()

    def check_converged_compartment_fraction(self, label, equil_time, test_fraction_diff):
        """Equilibrium test on ``fraction_soln`` walking back from the end (basepop.py:1099-1131)."""
        self.calculate_diagnostics()
        times = self.target_times
        fraction = self.fraction_soln[label]
        i, time_diff = -2, 0
        while time_diff < equil_time:
            i -= 1
            if -i >= len(times):
                break
            time_diff = abs(times[-1] - times[i])
            if abs(fraction[-1] - fraction[i]) > test_fraction_diff:
                return False
        return True

    def make_flow_diagram_png(self, png):
        """Graphviz rendering of the flow graph (basepop.py:1027-1097) is visualisation and out
        of scope; kept as a no-op that says so."""
        print('make_flow_diagram_png: flow-diagram rendering is not part of popdynamics_b200')
