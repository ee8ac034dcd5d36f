# -*- coding: utf-8 -*-
"""
Flow compiler and ``calculate_vars`` tracer.

Lowers a ``BaseModel`` (popdynamics_b200/basepop.py; reference /root/reference/basepop.py) into

* a flat SoA **flow table** of ``(kind, from, to, rate slot)`` rows in the canonical evaluation
  order of ``calculate_flows`` / ``calculate_events`` (basepop.py:566-598, :703-735): entry flows,
  var transfers, fixed transfers, one background-death row per compartment, infection deaths;
* an fp64 **expression DAG** obtained by running the user's ``calculate_vars`` hook
  (basepop.py:553) ONCE with symbolic proxies for ``self.compartments`` / ``self.params`` /
  ``self.vars`` / ``self.time``.  Operations are recorded in program order and never
  re-associated, so the device evaluates exactly the reference's arithmetic (kernels are built
  with ``--fmad=false``).

The DAG is emitted as CUDA C for NVRTC (``cuda_source``); the CPU oracle builds its own bytecode
from the same graph in oracle/oracle.py (test infrastructure).  Python constructs that cannot be lowered
(``bool()`` of a traced value, ``math.*`` on it, ``int()``/``float()``, ``**``, numpy ufuncs) raise
``TraceError`` -- nothing is silently run on the CPU.
"""

from __future__ import division

import builtins
import contextlib
import hashlib
import math
import numbers
import os
import sys
import threading

import numpy

ENTRY, VAR, FIXED, BGDEATH, INFDEATH = 0, 1, 2, 3, 4
KIND_NAMES = ('entry', 'var', 'fixed', 'bgdeath', 'infdeath')

# node operations of the expression graph
OPS = ('const', 'add', 'sub', 'mul', 'div', 'neg', 'fabs', 'floor', 'lt', 'le', 'gt', 'ge',
       'eq', 'ne', 'and', 'or', 'not', 'isfinite', 'select', 'exp')
_BOOL_OPS = frozenset(('lt', 'le', 'gt', 'ge', 'eq', 'ne', 'and', 'or', 'not', 'isfinite'))


class TraceError(TypeError):
    """A Python construct in a traced hook cannot be lowered to the device."""


# =============================================================================================
# expression graph
# =============================================================================================

class Graph(object):
    """Hash-consed expression DAG.  Node = (op, a, b, c); leaves are ('comp', i), ('param', j),
    ('time',), ('su', k), ('const', repr) and ('comp2', i): compartment i AFTER the update of a
    stochastic step, seen only by a traced checks()."""

    def __init__(self):
        self.nodes = []
        self.index = {}
        self.const_values = {}
        self.assertions = []   # node ids recorded by traced `assert` statements
        self.collect_asserts = False
        self.explorer = None   # _BranchExplorer while a scale-up closure is being traced

    def add(self, op, a=-1, b=-1, c=-1):
        key = (op, a, b, c)
        hit = self.index.get(key)
        if hit is None:
            hit = len(self.nodes)
            self.nodes.append(key)
            self.index[key] = hit
        return hit

    def const(self, value):
        value = float(value)
        key = ('const', value.hex())
        hit = self.index.get(key)
        if hit is None:
            hit = len(self.nodes)
            self.nodes.append(key)
            self.index[key] = hit
            self.const_values[hit] = value
        return hit


def _is_intlike(x):
    return isinstance(x, numbers.Integral) and not isinstance(x, bool)


class Sym(object):
    """Symbolic fp64 value.  ``is_int`` records whether the Python value would be an ``int``
    (stochastic integrators hold integer compartments, basepop.py:747-748, :805-806): it steers
    ``old_div`` and the path Python's ``sum`` takes, the value itself is always carried in fp64
    (exact below 2**53)."""

    __slots__ = ('g', 'n', 'is_int', 'is_bool', 'first_int')
    __array_ufunc__ = None      # numpy scalars defer to our reflected operators
    __hash__ = None

    def __init__(self, g, n, is_int=False, is_bool=False, first_int=False):
        self.g, self.n, self.is_int, self.is_bool = g, n, is_int, is_bool
        # an ODE compartment whose INITIAL value is a Python int: the reference's first
        # derivative call sees that int, every later one a float (basepop.py:683-684)
        self.first_int = first_int

    # -- coercion ---------------------------------------------------------------------------
    def _lift(self, other):
        if isinstance(other, Sym):
            if other.g is not self.g:
                raise TraceError('traced values from two different traces were mixed')
            return other
        if isinstance(other, bool):
            return Sym(self.g, self.g.const(float(other)), is_int=True)
        if isinstance(other, numbers.Real):
            return Sym(self.g, self.g.const(float(other)), is_int=_is_intlike(other))
        if isinstance(other, numpy.ndarray):
            raise TraceError('numpy arrays cannot be mixed with traced values inside a hook; '
                             'pass per-member arrays through set_param / set_compartment')
        return None

    def _bin(self, op, other, swap=False):
        rhs = self._lift(other)
        if rhs is None:
            return NotImplemented
        a, b = (rhs, self) if swap else (self, rhs)
        if a.is_bool or b.is_bool:
            raise TraceError('arithmetic on a traced comparison result is not supported')
        as_int = a.is_int and b.is_int and op != 'div'
        return Sym(self.g, self.g.add(op, a.n, b.n), is_int=as_int,
                   first_int=(a.is_int or a.first_int) and (b.is_int or b.first_int)
                   and not as_int and op != 'div')

    def __add__(self, o): return self._bin('add', o)
    def __radd__(self, o): return self._bin('add', o, True)
    def __sub__(self, o): return self._bin('sub', o)
    def __rsub__(self, o): return self._bin('sub', o, True)
    def __mul__(self, o): return self._bin('mul', o)
    def __rmul__(self, o): return self._bin('mul', o, True)
    def __truediv__(self, o): return self._bin('div', o)
    def __rtruediv__(self, o): return self._bin('div', o, True)

    def __floordiv__(self, o):
        rhs = self._lift(o)
        if rhs is None:
            return NotImplemented
        return _floordiv(self, rhs)

    def __rfloordiv__(self, o):
        lhs = self._lift(o)
        if lhs is None:
            return NotImplemented
        return _floordiv(lhs, self)

    def __neg__(self):
        return Sym(self.g, self.g.add('neg', self.n), is_int=self.is_int)

    def __pos__(self):
        return self

    def __abs__(self):
        return Sym(self.g, self.g.add('fabs', self.n), is_int=self.is_int)

    def _cmp(self, op, other):
        rhs = self._lift(other)
        if rhs is None:
            return NotImplemented
        return Sym(self.g, self.g.add(op, self.n, rhs.n), is_bool=True)

    def __lt__(self, o): return self._cmp('lt', o)
    def __le__(self, o): return self._cmp('le', o)
    def __gt__(self, o): return self._cmp('gt', o)
    def __ge__(self, o): return self._cmp('ge', o)
    def __eq__(self, o): return self._cmp('eq', o)
    def __ne__(self, o): return self._cmp('ne', o)

    # -- loud rejections --------------------------------------------------------------------
    def __bool__(self):
        if self.g.explorer is not None and self.is_bool:
            # inside a scale-up closure: Python branches on the traced time are explored path
            # by path and joined with selects (_trace_scalar_function)
            return self.g.explorer.decide(self.n)
        if self.g.collect_asserts:
            # inside a traced checks(): `assert cond` becomes a device-side assertion
            self.g.assertions.append(self.n)
            return True
        raise TraceError(
            'a traced value was used in a Python branch (if / while / and / or / min / max); '
            'data-dependent control flow in calculate_vars cannot be lowered to the device')

    def _no_scalar(self, *args):
        raise TraceError(
            'a traced value was converted with float() / int() / math.* ; only + - * / '
            'abs() sum() old_div() and comparisons inside checks() are lowered to the device')

    __float__ = __int__ = __index__ = __complex__ = __round__ = __trunc__ = _no_scalar
    __floor__ = __ceil__ = _no_scalar

    def __pow__(self, o, mod=None):
        raise TraceError('** on a traced value is not lowered (libm pow has no bit-exact '
                         'device counterpart); write the product out')

    __rpow__ = __pow__

    def __mod__(self, o):
        raise TraceError('% on a traced value is not supported')

    __rmod__ = __mod__

    def __repr__(self):
        return 'Sym(#%d%s)' % (self.n, ',int' if self.is_int else '')


def _first_step_hazard(a, b):
    """old_div / // of two values that are ints on the FIRST derivative call only (initial
    compartments given as Python ints): the reference floors there and divides truly on every
    later sub-step; one compiled expression cannot be both."""
    if (a.first_int or b.first_int) and (a.is_int or a.first_int) and (b.is_int or b.first_int):
        raise TraceError(
            'old_div / // meets a compartment whose initial value is a Python int: the reference '
            'would floor on the first sub-step only; give the initial value as a float')


def _floordiv(a, b):
    _first_step_hazard(a, b)
    if not (a.is_int and b.is_int):
        raise TraceError('// is only lowered for two integer operands')
    q = a.g.add('div', a.n, b.n)
    return Sym(a.g, a.g.add('floor', q), is_int=True)


def sym_old_div(a, b):
    """``past.utils.old_div`` on traced values: floor division for two ints, else true division."""
    sym = a if isinstance(a, Sym) else b
    a, b = sym._lift(a), sym._lift(b)
    if a is None or b is None:
        raise TraceError('old_div operand cannot be traced')
    _first_step_hazard(a, b)
    if a.is_int and b.is_int:
        return _floordiv(a, b)
    return a / b


# how the correction term of the compensated sum() is emitted: 'ordered' (CPython's branch, operands
# selected first) or 'twosum' (branch-free, three additions more); same values either way.  A
# developer knob for A/B measurements, read once at import.
SUM_FORM = os.environ.get('POPDYN_SUM_FORM', 'ordered')


def traced_sum(iterable, start=0, naive=False):
    """Python's builtin ``sum`` over traced floats.

    CPython >= 3.12 (Python/bltinmodule.c, builtin_sum_impl) sums exact ``float`` objects with
    Neumaier compensation: leading ints are accumulated exactly, the first float is added
    plainly, later floats update ``(f, c)`` and ``c`` is added at the end when finite and
    non-zero.  ``naive=True`` emits the plain left fold instead (what ``sum`` does over
    ``numpy.float64`` items, i.e. after ``integrate(is_continue=True)``, and on CPython < 3.12).
    """
    items = list(iterable)
    if not any(isinstance(x, Sym) for x in items) and not isinstance(start, Sym):
        return _real_sum(items, start)
    total = start
    k = 0
    # exact part: while everything seen is an int (or nothing symbolic yet)
    while k < len(items) and (_is_intlike(items[k]) or
                              (isinstance(items[k], Sym) and items[k].is_int)):
        total = _fold0(total, items[k])
        k += 1
    if k == len(items):
        return total
    total = _fold0(total, items[k])            # first float: plain add
    k += 1
    if naive:
        for x in items[k:]:
            total = total + x
        return total
    g = total.g if isinstance(total, Sym) else next(x for x in items if isinstance(x, Sym)).g
    if not isinstance(total, Sym):
        total = Sym(g, g.const(total))
    comp = None
    for x in items[k:]:
        if _is_intlike(x) or (isinstance(x, Sym) and x.is_int):
            total = total + x                    # ints on the float path: plain add
            continue
        x = total._lift(x)
        t = total + x
        if SUM_FORM == 'twosum':
            # CPython's correction term -- (total - t) + x when |total| >= |x|, else
            # (x - t) + total -- is the rounding error of total + x, which is exactly
            # representable; Knuth's TwoSum yields that same number without the comparison, the
            # two magnitudes and the select (and the same non-finite compensation, which the end
            # test discards, on overflow): five additions and no select
            bb = t - total
            term = (total - (t - bb)) + (x - bb)
        else:
            # CPython's own branch with the operands selected BEFORE the arithmetic:
            # (big - t) + small, two additions behind one comparison and two selects.  The ODE
            # kernels are bound by fp64 additions and multiplications; a comparison costs them
            # next to nothing (profiles/r2_explicit_forms_ab.txt)
            big_first = g.add('ge', g.add('fabs', total.n), g.add('fabs', x.n))
            big = Sym(g, g.add('select', big_first, total.n, x.n))
            small = Sym(g, g.add('select', big_first, x.n, total.n))
            term = (big - t) + small
        comp = term if comp is None else comp + term
        total = t
    if comp is not None:
        nonzero = g.add('ne', comp.n, g.const(0.0))
        finite = g.add('isfinite', comp.n)
        cond = g.add('and', nonzero, finite)
        total = Sym(g, g.add('select', cond, (total + comp).n, total.n))
    return total


class _BranchExplorer(object):
    """Decision record of one run of a scale-up closure: the first ``len(forced)`` traced
    comparisons take the forced outcomes, later ones take True and are recorded."""

    def __init__(self, forced):
        self.forced = list(forced)
        self.taken = []          # (cond node, outcome) of every comparison met, in order

    def decide(self, node):
        k = len(self.taken)
        outcome = self.forced[k] if k < len(self.forced) else True
        self.taken.append((node, outcome))
        return outcome


def traced_exp(x):
    """``math.exp`` while a hook is traced: an 'exp' node for a traced value."""
    if isinstance(x, Sym):
        if x.is_bool:
            raise TraceError('exp() of a traced comparison result')
        return Sym(x.g, x.g.add('exp', x.n))
    return _real_exp(x)


_real_exp = math.exp


@contextlib.contextmanager
def _patched_exp():
    """Scale-up closures call ``math.exp`` (basepop.py:122): while one is traced, the stdlib
    function is shadowed by a version that records an 'exp' node for traced arguments."""
    math.exp = traced_exp
    try:
        yield
    finally:
        math.exp = _real_exp


def _trace_scalar_function(g, fn, x, what, max_paths=64):
    """Trace ``fn(x)`` for a traced scalar ``x`` where ``fn`` may branch on comparisons of ``x``
    (basepop.py:120-121 `if arg > 10.`, :148-155, examples/tb_model.py:93): every path is run
    once with its branch outcomes forced and the results are joined by selects on the recorded
    conditions.  ``fn`` must be a pure function of ``x``."""
    runs = [0]

    def run(forced):
        runs[0] += 1
        if runs[0] > max_paths:
            raise TraceError('%s branches into more than %d paths' % (what, max_paths))
        explorer = _BranchExplorer(forced)
        previous, g.explorer = g.explorer, explorer
        try:
            with _patched_exp():
                value = fn(x)
        finally:
            g.explorer = previous
        if isinstance(value, Sym):
            if value.is_bool:
                raise TraceError('%s returns a comparison result, not a number' % what)
        elif isinstance(value, numbers.Real):
            value = Sym(g, g.const(float(value)))
        else:
            raise TraceError('%s returned %s, not a number' % (what, type(value).__name__))
        return value, explorer.taken

    def build(forced):
        value, taken = run(forced)
        if len(taken) <= len(forced):
            return value
        # the first unforced comparison: this run took True there
        cond = taken[len(forced)][0]
        outcomes = [o for _, o in taken[:len(forced)]]
        if_true = build(outcomes + [True])
        if_false = build(outcomes + [False])
        return Sym(g, g.add('select', cond, if_true.n, if_false.n))

    return build([])


def _fold0(total, x):
    """``total + x`` where an exact int 0 on the left is dropped (``0 + x``)."""
    if _is_intlike(total) and total == 0 and isinstance(x, Sym):
        return x
    return total + x


_real_sum = builtins.sum


@contextlib.contextmanager
def _patched_sum(naive):
    def hooked(iterable, start=0):
        return traced_sum(iterable, start, naive=naive)
    builtins.sum = hooked
    try:
        yield
    finally:
        builtins.sum = _real_sum


# =============================================================================================
# traced views of the model's dictionaries
# =============================================================================================

class _ParamView(dict):
    """``self.params`` during a trace: numeric values come back as parameter symbols."""

    def __init__(self, tracer, source):
        dict.__init__(self, source)
        self._tracer = tracer

    def __getitem__(self, label):
        value = dict.__getitem__(self, label)
        sym = self._tracer.param_symbol(label, value)
        return value if sym is None else sym

    def get(self, label, default=None):
        return self[label] if label in self else default

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]


class CompiledModel(object):
    """Result of ``compile_model``: flow table, expression DAG, parameter / scale-up tables."""

    def __init__(self):
        self.labels = []
        self.method = None
        self.integer_state = False
        self.naive_sum = False
        self.graph = None
        self.order = []            # node ids in evaluation order (after dead-code elimination)
        self.param_labels = []     # slot -> label (None for anonymous constants)
        self.param_values = []     # slot -> float or 1-D array (per member)
        self.param_is_int = []
        self.su_labels = []
        self.flow_kind = []
        self.flow_from = []
        self.flow_to = []
        self.flow_node = []        # DAG node holding the rate multiplier
        self.flow_is_int = []
        self.flow_names = []
        self.check_nodes = []
        self.post_check_nodes = [] # user-defined checks() of the stochastic loops (after update)
        self.uses_time = False
        self.var_nodes = {}        # var label -> node id (all vars seen in the trace)
        self.diagnostics = False   # compiled for the diagnostics pass (basepop.py:911-960)
        self.diag_labels = []      # var_labels of the diagnostics pass, in the reference's order
        self.diag_nodes = []       # node of every diagnostic var
        self.diag_flow_nodes = []  # net flow of every compartment (flow_array)
        self.final_var_labels = [] # diagnostics vars of the final state, fused into the integrator
        self.final_var_nodes = []

    # ---- sizes
    @property
    def n_comp(self): return len(self.labels)
    @property
    def n_flow(self): return len(self.flow_kind)
    @property
    def n_param(self): return len(self.param_labels)
    @property
    def n_su(self): return len(self.su_labels)

    def member_param_slots(self):
        return [k for k, v in enumerate(self.param_values) if isinstance(v, numpy.ndarray)]

    def n_members_implied(self):
        sizes = set(v.size for v in self.param_values if isinstance(v, numpy.ndarray))
        if len(sizes) > 1:
            raise ValueError('per-member parameter arrays have different lengths: %s' % sizes)
        return sizes.pop() if sizes else None

    def uniform_params(self):
        """fp64 vector of all parameter slots; per-member slots hold NaN as a tripwire."""
        return numpy.array([numpy.nan if isinstance(v, numpy.ndarray) else float(v)
                            for v in self.param_values], dtype=numpy.float64)

    def flow_table(self):
        """The flat SoA flow table: int32 arrays kind / from / to (-1 = outside)."""
        as_i32 = lambda seq: numpy.asarray(seq, dtype=numpy.int32)
        return dict(kind=as_i32(self.flow_kind), frm=as_i32(self.flow_from),
                    to=as_i32(self.flow_to), is_int=as_i32(self.flow_is_int))

    # ---- CUDA C for NVRTC ------------------------------------------------------------------
    def cuda_source(self):
        """The generated model header (memoised: a CompiledModel is not mutated after
        compile_model returns)."""
        from .codegen import emit_model_header
        if getattr(self, '_cuda_source', None) is None:
            self._cuda_source = emit_model_header(self)
        return self._cuda_source

    def structure_key(self):
        """Hash of everything that shapes the generated kernel (NOT parameter values)."""
        h = hashlib.sha256()
        h.update(self.cuda_source().encode())
        return h.hexdigest()[:24]


class _Tracer(object):
    def __init__(self, model, integer_state):
        self.model = model
        self.g = Graph()
        self.integer_state = integer_state
        self.param_slot = {}
        self.param_labels, self.param_values, self.param_is_int = [], [], []

    def param_symbol(self, label, value):
        if isinstance(value, Sym):
            return value
        if isinstance(value, bool) or not (isinstance(value, numbers.Real) or
                                           isinstance(value, numpy.ndarray)):
            return None                       # strings, lists ...: static Python data
        slot = self.param_slot.get(label)
        if slot is None:
            slot = len(self.param_labels)
            self.param_slot[label] = slot
            self.param_labels.append(label)
            if isinstance(value, numpy.ndarray):
                if value.ndim != 1:
                    raise ValueError('parameter %r: only 1-D per-member arrays are supported'
                                     % (label,))
                value = numpy.ascontiguousarray(value, dtype=numpy.float64)
            self.param_values.append(value)
            self.param_is_int.append(_is_intlike(value))
        return Sym(self.g, self.g.add('param', slot), is_int=self.param_is_int[slot])

    def anonymous_param(self, value, tag):
        return self.param_symbol('<%s>' % tag, value)


# Tracing shadows builtins.sum and math.exp process-wide while a hook runs: one trace at a time.
_TRACE_LOCK = threading.RLock()


def compile_model(model, method, is_continue=False, naive_sum=None, diagnostics=False,
                  want_vars=None, final_vars=None):
    """Thread-safe front door of ``_compile_model`` (see there)."""
    with _TRACE_LOCK:
        return _compile_model(model, method, is_continue, naive_sum, diagnostics, want_vars,
                              final_vars)


def _compile_model(model, method, is_continue=False, naive_sum=None, diagnostics=False,
                   want_vars=None, final_vars=None):
    """Lower ``model`` for ``method`` ('explicit', 'discrete_time_stochastic',
    'continuous_time_stochastic', 'scipy').  ``model.set_flows()`` must already have run
    (``check_flow_transfers``, basepop.py:638-656).

    'scipy' (basepop.py:882-886: ``odeint`` on the derivative closure) is the explicit lowering
    with two differences: every member is at its own time, so the scale-up closures are traced on
    the symbolic time instead of tabulated per sub-step; and the state arrives as a numpy array,
    i.e. ``numpy.float64`` compartments, over which Python's ``sum`` folds plainly.

    ``diagnostics=True`` lowers the diagnostics pass instead (``calculate_diagnostics``,
    basepop.py:911-960): at every stored time the compartments are the ``numpy.float64`` values
    of ``soln_array`` -- floats whatever the integrator was, so ``old_div`` divides truly and
    ``sum`` folds naively -- then ``calculate_vars``, ``calculate_flows`` and the
    ``calculate_diagnostic_vars`` hook run; every numeric entry of ``self.vars`` is an output
    (``var_array``) and so is the net flow of every compartment (``flow_array``).  ``want_vars``
    (a list of var labels) keeps only those outputs, in the reference's order: everything that
    feeds none of them is eliminated from the device code.

    ``final_vars`` (ODE methods only) fuses the diagnostics pass of the FINAL state into the
    integrator: the same hooks are traced a second time on the final compartments the way
    ``calculate_diagnostics`` runs them (plain ``sum`` fold, vars not cleared) and the kernel
    writes those vars instead of the state -- one scalar per member for a calibration sweep that
    reads ``model.vars['proportion']`` (examples/rmit_tb_model.py:143)."""
    from .basepop import BaseModel

    integer_state = method not in ('explicit', 'scipy') and not diagnostics
    if diagnostics or method == 'scipy':
        naive_sum = True
    if naive_sum is None:
        # after is_continue the compartments are numpy.float64 (basepop.py:932-933) and Python's
        # sum() falls off its compensated float fast path
        naive_sum = bool(is_continue) and not integer_state
    if sys.version_info < (3, 12):
        naive_sum = True          # builtin sum() compensates float sums from CPython 3.12 on
    tr = _Tracer(model, integer_state)
    g = tr.g
    cm = CompiledModel()
    cm.labels = list(model.labels)
    cm.method = method
    cm.integer_state = integer_state
    cm.naive_sum = naive_sum
    cm.diagnostics = bool(diagnostics)
    cm.graph = g
    index_of = {label: i for i, label in enumerate(cm.labels)}

    first_int = {label: (not integer_state and not diagnostics and not is_continue and
                         method in ('explicit',) and
                         _is_intlike(model.init_compartments.get(label)))
                 for label in cm.labels}
    comps = {label: Sym(g, g.add('comp', i), is_int=integer_state, first_int=first_int[label])
             for i, label in enumerate(cm.labels)}
    time_sym = Sym(g, g.add('time'))

    # ---- vars dictionary at the start of the hook -----------------------------------------
    if method == 'explicit':
        # prepare_vars_and_flows: vars.clear(); scale-ups; calculate_vars (basepop.py:606-610)
        cm.su_labels = list(model.scaleup_fns.keys())
        traced_vars = {label: Sym(g, g.add('su', k)) for k, label in enumerate(cm.su_labels)}
    elif method == 'scipy':
        # same prologue, the closures traced on the member's own time
        traced_vars = {label: _trace_scalar_function(g, fn, time_sym,
                                                     'scale-up function %r' % (label,))
                       for label, fn in model.scaleup_fns.items()}
    else:
        # the stochastic loops never clear vars nor evaluate scale-ups (basepop.py:763-765,
        # :822-824): whatever is left in self.vars is visible as stale constants
        traced_vars = {k: v for k, v in model.vars.items() if isinstance(v, numbers.Real)}

    def traced_scaleup_vars():
        # a hook that calls self.calculate_scaleup_vars() itself (declared extension for the
        # stochastic integrators): the closures are traced on the symbolic time
        # and evaluated on the device at every member's own time
        for label, fn in model.scaleup_fns.items():
            model.vars[label] = _trace_scalar_function(g, fn, time_sym,
                                                       'scale-up function %r' % (label,))

    # flow lists whose resolved rates are parameter symbols while the hooks run
    fixed_syms = []
    for from_label, to_label, rate in model.fixed_transfer_rate_flows:
        plabel = model._fixed_rate_params.get((from_label, to_label))
        if plabel is not None and plabel in model.params and _same(model.params[plabel], rate):
            sym = tr.param_symbol(plabel, model.params[plabel])
        else:
            sym = tr.anonymous_param(rate, 'fixed:%s>%s' % (from_label, to_label))
        fixed_syms.append((from_label, to_label, sym))
    inf_syms = []
    for label, rate in model.infection_death_rate_flows:
        plabel = model._infection_death_params.get(label)
        if plabel is not None and plabel in model.params and _same(model.params[plabel], rate):
            sym = tr.param_symbol(plabel, model.params[plabel])
        else:
            sym = tr.anonymous_param(rate, 'infdeath:%s' % label)
        inf_syms.append((label, sym))
    bg = model.background_death_rate
    plabel = model._background_death_param
    bg_is_default = plabel is None and not isinstance(bg, numpy.ndarray) and bg == 0.
    if bg_is_default:
        bg_sym = None
    elif plabel is not None and plabel in model.params and _same(model.params[plabel], bg):
        bg_sym = tr.param_symbol(plabel, model.params[plabel])
    else:
        bg_sym = tr.anonymous_param(bg, 'background_death')

    initial_vars = dict(traced_vars)
    if final_vars and (integer_state or diagnostics):
        raise ValueError("final_vars is defined for the ODE methods ('explicit', 'scipy')")
    saved = dict(params=model.params, compartments=model.compartments, vars=model.vars,
                 time=model.time, fixed=model.fixed_transfer_rate_flows,
                 inf=model.infection_death_rate_flows, bg=model.background_death_rate,
                 flows=model.flows)
    try:
        model.params = _ParamView(tr, saved['params'])
        model.compartments = comps
        model.vars = traced_vars
        model.time = time_sym
        model.fixed_transfer_rate_flows = fixed_syms
        model.infection_death_rate_flows = inf_syms
        if bg_sym is not None:
            model.background_death_rate = bg_sym
        if method != 'explicit':
            model.calculate_scaleup_vars = traced_scaleup_vars
        else:
            # a hook that calls it again re-reads the values of this sub-step (host table)
            su_syms = dict(traced_vars)
            model.calculate_scaleup_vars = lambda: model.vars.update(su_syms)
        try:
            with _patched_sum(naive_sum):
                model.calculate_vars()
        finally:
            model.__dict__.pop('calculate_scaleup_vars', None)
        hook_vars = dict(model.vars)

        def as_sym(value, what):
            if isinstance(value, Sym):
                if value.is_bool:
                    raise TraceError('%s is a comparison result, not a number' % what)
                return value
            if isinstance(value, numbers.Real):
                return Sym(g, g.const(float(value)), is_int=_is_intlike(value))
            raise TraceError('%s has unsupported type %s' % (what, type(value).__name__))

        # ---- flow table in canonical order (basepop.py:566-598 / :703-735) -----------------
        def add_flow(kind, frm, to, sym, name):
            cm.flow_kind.append(kind)
            cm.flow_from.append(-1 if frm is None else index_of[frm])
            cm.flow_to.append(-1 if to is None else index_of[to])
            cm.flow_node.append(sym.n)
            cm.flow_is_int.append(1 if (kind == ENTRY and sym.is_int) else 0)
            cm.flow_names.append(name)

        for label, var_label in model.var_entry_rate_flow:
            add_flow(ENTRY, None, label, as_sym(hook_vars[var_label], 'vars[%r]' % var_label),
                     '->%s [%s]' % (label, var_label))
        for from_label, to_label, var_label in model.var_transfer_rate_flows:
            add_flow(VAR, from_label, to_label,
                     as_sym(hook_vars[var_label], 'vars[%r]' % var_label),
                     '%s->%s [%s]' % (from_label, to_label, var_label))
        for from_label, to_label, sym in fixed_syms:
            add_flow(FIXED, from_label, to_label, sym, '%s->%s' % (from_label, to_label))
        if bg_sym is not None:
            for label in cm.labels:
                add_flow(BGDEATH, label, None, bg_sym, '%s->death' % label)
        for label, sym in inf_syms:
            add_flow(INFDEATH, label, None, sym, '%s->infection death' % label)

        # ---- checks() ----------------------------------------------------------------------
        default_checks = type(model).checks is BaseModel.checks
        if default_checks:
            zero = g.const(0.0)
            cm.check_nodes = [g.add('ge', comps[label].n, zero) for label in cm.labels]
            if integer_state:
                cm.check_nodes = []       # integer counts never go negative (clamped draws)
        else:
            if integer_state:
                # The stochastic loops call checks() AFTER the update of a step / an event
                # (basepop.py:789, :846): it sees the UPDATED integer compartments (a second set
                # of leaves, 'comp2'), while self.vars still holds what calculate_vars and
                # calculate_events left at the start of the step -- including the two death
                # totals, which calculate_events accumulates over positive terms only (:724-735).
                rate_death, rate_inf = _death_totals(cm, g, comps, bg_sym, inf_syms,
                                                     positive_only=True)
                model.compartments = {label: Sym(g, g.add('comp2', i), is_int=True)
                                      for i, label in enumerate(cm.labels)}
            else:
                # vars visible to checks(): calculate_flows has also defined the two death totals
                rate_death, rate_inf = _death_totals(cm, g, comps, bg_sym, inf_syms)
            model.vars['rate_death'] = rate_death
            model.vars['rate_infection_death'] = rate_inf
            g.collect_asserts = True
            try:
                with _patched_sum(naive_sum):
                    model.checks()
            finally:
                g.collect_asserts = False
                model.compartments = comps
            if integer_state:
                cm.post_check_nodes = list(g.assertions)
            else:
                cm.check_nodes = list(g.assertions)
        cm.var_nodes = {k: as_sym(v, 'vars[%r]' % k).n for k, v in hook_vars.items()
                        if isinstance(v, (Sym, numbers.Real)) and not getattr(v, 'is_bool', False)}

        # ---- diagnostics pass: calculate_flows + the diagnostic hook on the traced state -----
        if diagnostics:
            model.vars = dict(hook_vars)
            model.flows = {}
            with _patched_sum(naive_sum):
                model.calculate_flows()                  # basepop.py:560-598, incl. rate_death
                model.calculate_diagnostic_vars()        # hook, basepop.py:905
            for label, value in model.vars.items():
                if isinstance(value, Sym) and value.is_bool:
                    continue
                if want_vars is not None and label not in want_vars:
                    continue
                if isinstance(value, (Sym, numbers.Real)) and not isinstance(value, bool):
                    cm.diag_labels.append(label)
                    cm.diag_nodes.append(as_sym(value, 'vars[%r]' % label).n)
            if want_vars is not None:
                missing = [label for label in want_vars if label not in cm.diag_labels]
                if missing:
                    raise KeyError('the diagnostics pass defines no var(s) %s' % missing)
            cm.diag_flow_nodes = [as_sym(model.flows[label], 'flows[%r]' % label).n
                                  for label in cm.labels]
        if final_vars:
            # calculate_diagnostics on the final state (basepop.py:927-948): compartments are the
            # numpy.float64 of soln_array (plain sum fold), vars still hold the scale-up values
            # of the last derivative call, then the three hooks run
            model.vars = dict(initial_vars)
            model.flows = {}
            with _patched_sum(True):
                model.calculate_vars()
                model.calculate_flows()
                model.calculate_diagnostic_vars()
            for label in final_vars:
                if label not in model.vars:
                    raise KeyError('the diagnostics pass defines no var %r' % (label,))
                cm.final_var_labels.append(label)
                cm.final_var_nodes.append(as_sym(model.vars[label], 'vars[%r]' % label).n)
    finally:
        model.params = saved['params']
        model.compartments = saved['compartments']
        model.vars = saved['vars']
        model.time = saved['time']
        model.fixed_transfer_rate_flows = saved['fixed']
        model.infection_death_rate_flows = saved['inf']
        model.background_death_rate = saved['bg']
        model.flows = saved['flows']

    cm.param_labels = tr.param_labels
    cm.param_values = tr.param_values
    cm.param_is_int = tr.param_is_int

    cm.order = live_order(cm)
    cm.uses_time = any(g.nodes[n][0] == 'time' for n in cm.order)
    return cm


def params_read_by(cm, roots):
    """Parameter slots the expression graph reads on the way to ``roots`` (node ids)."""
    g = cm.graph
    seen, slots, stack = set(), set(), list(roots)
    while stack:
        n = stack.pop()
        if n in seen:
            continue
        seen.add(n)
        node = g.nodes[n]
        if node[0] == 'param':
            slots.add(node[1])
        elif node[0] not in ('comp', 'comp2', 'time', 'su', 'const'):
            stack.extend(x for x in node[1:] if x >= 0)
    return slots


def live_order(cm):
    """Dead-code elimination + evaluation order: the nodes reachable from the flow rates, the
    checks and (diagnostics pass) the requested vars / net flows, in creation order, which is
    topological."""
    g = cm.graph
    roots = list(cm.flow_node) + list(cm.check_nodes) + list(cm.post_check_nodes) + \
        list(cm.diag_nodes) + list(cm.diag_flow_nodes) + list(cm.final_var_nodes)
    live = set()
    stack = list(roots)
    while stack:
        n = stack.pop()
        if n in live:
            continue
        live.add(n)
        node = g.nodes[n]
        if node[0] not in ('comp', 'comp2', 'param', 'time', 'su', 'const'):
            stack.extend(x for x in node[1:] if x >= 0)
    return [n for n in range(len(g.nodes)) if n in live]


def _same(a, b):
    if isinstance(a, numpy.ndarray) or isinstance(b, numpy.ndarray):
        return a is b
    return a == b


def _death_totals(cm, g, comps, bg_sym, inf_syms, positive_only=False):
    """vars['rate_death'] / vars['rate_infection_death'] as calculate_flows accumulates them
    (basepop.py:587-598), or -- ``positive_only`` -- as calculate_events does, where a term only
    counts when it is positive (basepop.py:724-735)."""
    def add(total, val):
        if not positive_only or not isinstance(val, Sym):
            return total + val
        grown = total + val
        if not isinstance(total, Sym):
            total = Sym(g, g.const(float(total)))
        cond = g.add('gt', val.n, g.const(0.0))
        return Sym(g, g.add('select', cond, grown.n, total.n))

    rate_death = 0.
    if bg_sym is not None:
        for label in cm.labels:
            rate_death = add(rate_death, comps[label] * bg_sym)
    rate_inf = 0.
    for label, sym in inf_syms:
        rate_inf = add(rate_inf, comps[label] * sym)
    return rate_death, rate_inf


def explicit_schedule(target_times, min_dt=0.05):
    """Host replay of the sub-step schedule of ``integrate_explicit`` (basepop.py:672-682).  It is
    member independent: fp64 accumulation ``time = time + min_dt`` with a snap to the target on
    overshoot, so some intervals take 21 sub-steps, the last with dt ~ 1e-16.  Returns
    ``(t_eval, dt, n_sub)``: evaluation time and step size of every derivative call, and the
    number of sub-steps in each of the T-1 target intervals."""
    t_eval, dts, n_sub = [], [], []
    time = float(target_times[0])
    min_dt = float(min_dt)
    for new_time in target_times[1:]:
        new_time = float(new_time)
        count = 0
        while time < new_time:
            old_time = time
            time = time + min_dt
            dt = min_dt
            if time > new_time:
                dt = new_time - old_time
                time = new_time
            t_eval.append(old_time)
            dts.append(dt)
            count += 1
        n_sub.append(count)
    return (numpy.asarray(t_eval, dtype=numpy.float64), numpy.asarray(dts, dtype=numpy.float64),
            numpy.asarray(n_sub, dtype=numpy.int32))
