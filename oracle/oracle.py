# -*- coding: utf-8 -*-
"""
ctypes front-end of the CPU oracle ``libpopdyn_oracle.so`` (oracle/popdyn_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the CPU arms of
``bench.py``.  The product package ``popdynamics_b200`` never imports this module.
"""

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libpopdyn_oracle.so')
_lib = None

ERRORS = {0: 'ok', 1: 'checks() failed', 2: 'bad Poisson mean', 3: 'zero total rate',
          4: 'uniform stream exhausted', 5: 'log(0)'}


class PdoModel(C.Structure):
    _fields_ = [('n_comp', C.c_int32), ('n_param', C.c_int32), ('n_su', C.c_int32),
                ('n_ins', C.c_int32), ('n_flow', C.c_int32), ('n_check', C.c_int32),
                ('n_post_check', C.c_int32), ('pad1', C.c_int32),
                ('ins', C.c_void_p), ('consts', C.c_void_p), ('kind', C.c_void_p),
                ('frm', C.c_void_p), ('to', C.c_void_p), ('slot', C.c_void_p),
                ('is_int', C.c_void_p), ('check', C.c_void_p), ('post_check', C.c_void_p)]


class PdoRng(C.Structure):
    _fields_ = [('kind', C.c_int32), ('mti', C.c_int32), ('mt', C.c_uint32 * 624),
                ('u', C.c_void_p), ('n_u', C.c_int64), ('stride', C.c_int64),
                ('cursor', C.c_int64), ('key', C.c_uint32 * 2), ('ctr', C.c_uint32 * 4),
                ('spare', C.c_double), ('has_spare', C.c_int32), ('exhausted', C.c_int32)]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, 'popdyn_oracle.c')
    stale = (not os.path.exists(_LIB_PATH) or
             os.path.getmtime(_LIB_PATH) < os.path.getmtime(src) or
             os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, 'popdyn_oracle.h')))
    if force or stale:
        subprocess.check_call(['make', '-s', '-C', _HERE, '-B', 'libpopdyn_oracle.so'])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        P = C.POINTER
        L.pdo_rng_seed_numpy.argtypes = [P(PdoRng), C.c_uint32]
        L.pdo_rng_seed_python.argtypes = [P(PdoRng), C.c_uint64]
        L.pdo_rng_inject.argtypes = [P(PdoRng), C.c_void_p, C.c_int64, C.c_int64]
        L.pdo_rng_philox.argtypes = [P(PdoRng), C.c_uint64, C.c_uint64]
        L.pdo_rng_double.argtypes = [P(PdoRng)]
        L.pdo_rng_double.restype = C.c_double
        L.pdo_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdo_poisson.argtypes = [P(PdoRng), C.c_double]
        L.pdo_poisson.restype = C.c_int64
        L.pdo_pick_event.argtypes = [C.c_void_p, C.c_int32, C.c_double]
        L.pdo_pick_event.restype = C.c_int32
        L.pdo_explicit_schedule.argtypes = [C.c_int32, C.c_void_p, C.c_double, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int64]
        L.pdo_explicit_schedule.restype = C.c_int64
        L.pdo_explicit.argtypes = [P(PdoModel), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_double, C.c_void_p, C.c_void_p, P(C.c_int32)]
        L.pdo_explicit.restype = C.c_int64
        for fn in (L.pdo_tau_leap, L.pdo_gillespie):
            fn.argtypes = [P(PdoModel), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                           P(PdoRng), C.c_void_p, P(C.c_int32)]
            fn.restype = C.c_int64
        for fn in (L.pdo_tau_leap_batch, L.pdo_gillespie_batch):
            fn.argtypes = [P(PdoModel), C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                           C.c_int64, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_int32, P(C.c_int32)]
            fn.restype = C.c_int64
        L.pdo_explicit_batch.argtypes = [P(PdoModel), C.c_int64, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                         C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int32, P(C.c_int32)]
        L.pdo_explicit_batch.restype = C.c_int64
        L.pdo_eval_vars.argtypes = [P(PdoModel), C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                    C.c_void_p]
        L.pdo_max_threads.restype = C.c_int32
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# opcode numbering of the oracle's bytecode interpreter (oracle/popdyn_oracle.h)
OPS = ('const', 'add', 'sub', 'mul', 'div', 'neg', 'fabs', 'floor', 'lt', 'le', 'gt', 'ge',
       'eq', 'ne', 'and', 'or', 'not', 'isfinite', 'select', 'exp')
OPCODE = {name: i for i, name in enumerate(OPS)}


def oracle_program(cm):
    """The traced expression graph of a ``CompiledModel`` as the bytecode the C oracle
    interprets.  Value slots: [compartments | parameters | time | scale-ups | updated
    compartments (only when a checks() was traced for the stochastic loops) | one per
    instruction]; instructions in the graph's evaluation order (program order of the hook)."""
    g = cm.graph
    C, NP, NSU = cm.n_comp, cm.n_param, cm.n_su
    post = list(getattr(cm, 'post_check_nodes', []))
    base = C + NP + 1 + NSU + (C if post else 0)
    slot_of, ins, consts = {}, [], []
    for n in cm.order:
        node = g.nodes[n]
        op = node[0]
        if op == 'comp':
            slot_of[n] = node[1]
        elif op == 'param':
            slot_of[n] = C + node[1]
        elif op == 'time':
            slot_of[n] = C + NP
        elif op == 'su':
            slot_of[n] = C + NP + 1 + node[1]
        elif op == 'comp2':
            slot_of[n] = C + NP + 1 + NSU + node[1]
        elif op == 'const':
            consts.append(g.const_values[n])
            slot_of[n] = base + len(ins)
            ins.append((OPCODE['const'], len(consts) - 1, -1, -1))
        else:
            args = [slot_of[x] if x >= 0 else -1 for x in node[1:]]
            slot_of[n] = base + len(ins)
            ins.append((OPCODE[op], args[0], args[1], args[2]))
    return dict(
        ins=np.asarray(ins, dtype=np.int32).reshape(-1, 4),
        consts=np.asarray(consts, dtype=np.float64),
        slot=np.asarray([slot_of[n] for n in cm.flow_node], dtype=np.int32),
        check=np.asarray([slot_of[n] for n in cm.check_nodes], dtype=np.int32),
        post_check=np.asarray([slot_of[n] for n in post], dtype=np.int32),
        n_values=base + len(ins), slot_of=slot_of)


class OracleModel(object):
    """A ``CompiledModel`` (popdynamics_b200.compiler) handed to the C oracle."""

    def __init__(self, cm):
        self.cm = cm
        prog = oracle_program(cm)
        table = cm.flow_table()
        self._keep = dict(ins=np.ascontiguousarray(prog['ins']), consts=prog['consts'],
                          kind=table['kind'], frm=table['frm'], to=table['to'],
                          slot=prog['slot'], is_int=table['is_int'], check=prog['check'],
                          post_check=prog['post_check'])
        k = self._keep
        self.n_values = prog['n_values']
        self.slot_of = prog['slot_of']
        self.struct = PdoModel(cm.n_comp, cm.n_param, cm.n_su, len(k['ins']), cm.n_flow,
                               len(k['check']), len(k['post_check']), 0, _ptr(k['ins']),
                               _ptr(k['consts']), _ptr(k['kind']), _ptr(k['frm']), _ptr(k['to']),
                               _ptr(k['slot']), _ptr(k['is_int']), _ptr(k['check']),
                               _ptr(k['post_check']))
        self.C = cm.n_comp

    # ---- parameter helpers -------------------------------------------------------------
    def member_params(self, i=None):
        """fp64 parameter vector of member ``i`` (or the uniform vector)."""
        vals = []
        for v in self.cm.param_values:
            vals.append(float(v[i]) if isinstance(v, np.ndarray) else float(v))
        return np.asarray(vals, dtype=np.float64)

    def params_matrix(self, n_member):
        out = np.empty((n_member, max(self.cm.n_param, 1)), dtype=np.float64)
        for k, v in enumerate(self.cm.param_values):
            out[:, k] = v if isinstance(v, np.ndarray) else float(v)
        return out

    # ---- single-member integrators ------------------------------------------------------
    def explicit(self, y0, target_times, params=None, min_dt=0.05, su_table=None):
        tt = _f64(target_times)
        p = _f64(self.member_params() if params is None else params)
        y0 = _f64(y0)
        soln = np.zeros((len(tt), self.C))
        err = C.c_int32(0)
        su = None if su_table is None else _f64(su_table)
        n = lib().pdo_explicit(C.byref(self.struct), _ptr(y0), _ptr(p), len(tt), _ptr(tt),
                               float(min_dt), _ptr(su), _ptr(soln), C.byref(err))
        return soln, int(n), err.value

    def explicit_first_failure(self, y0, target_times, params=None, min_dt=0.05, su_table=None):
        """Target index of the interval in which checks() first fails, 0 if never."""
        tt = _f64(target_times)
        p = _f64(self.member_params() if params is None else params)
        su = None if su_table is None else _f64(su_table)
        fn = lib().pdo_explicit_first_failure
        fn.restype = C.c_int32
        return int(fn(C.byref(self.struct), _ptr(_f64(y0)), _ptr(p), len(tt), _ptr(tt),
                      C.c_double(float(min_dt)), _ptr(su)))

    def _stochastic(self, fn, y0, target_times, rng, params):
        tt = _f64(target_times)
        p = _f64(self.member_params() if params is None else params)
        y0 = _f64(y0)
        soln = np.zeros((len(tt), self.C))
        err = C.c_int32(0)
        n = fn(C.byref(self.struct), _ptr(y0), _ptr(p), len(tt), _ptr(tt), C.byref(rng.struct),
               _ptr(soln), C.byref(err))
        return soln, int(n), err.value

    def tau_leap(self, y0, target_times, rng, params=None):
        return self._stochastic(lib().pdo_tau_leap, y0, target_times, rng, params)

    def gillespie(self, y0, target_times, rng, params=None):
        return self._stochastic(lib().pdo_gillespie, y0, target_times, rng, params)

    def eval_values(self, y, params=None, time=0.0, su=None):
        p = _f64(self.member_params() if params is None else params)
        y = _f64(y)
        v = np.zeros(self.n_values)
        lib().pdo_eval_vars(C.byref(self.struct), _ptr(y), _ptr(p), float(time),
                            _ptr(None if su is None else _f64(su)), _ptr(v))
        return v

    # ---- batch drivers ------------------------------------------------------------------
    def _batch_inputs(self, n_member, y0):
        y0 = _f64(y0)
        if y0.ndim == 1:
            y0_stride = 0
        else:
            assert y0.shape == (n_member, self.C)
            y0_stride = self.C
        if self.cm.member_param_slots():
            params = self.params_matrix(n_member)
            p_stride = params.shape[1]
        else:
            params = self.member_params()
            p_stride = 0
        return y0, y0_stride, np.ascontiguousarray(params), p_stride

    def stochastic_batch(self, method, n_member, y0, target_times, seed, member0=0,
                         want_soln=True, want_moments=False, n_thread=0):
        fn = lib().pdo_tau_leap_batch if method == 'discrete_time_stochastic' \
            else lib().pdo_gillespie_batch
        tt = _f64(target_times)
        y0, y0_stride, params, p_stride = self._batch_inputs(n_member, y0)
        soln = np.zeros((n_member, len(tt), self.C)) if want_soln else None
        s = np.zeros((len(tt), self.C), dtype=np.int64) if want_moments else None
        q = np.zeros((len(tt), self.C), dtype=np.int64) if want_moments else None
        err = C.c_int32(0)
        n = fn(C.byref(self.struct), n_member, member0, _ptr(y0), y0_stride, _ptr(params),
               p_stride, len(tt), _ptr(tt), int(seed), _ptr(soln), _ptr(s), _ptr(q),
               int(n_thread), C.byref(err))
        return dict(soln=soln, sum=s, sumsq=q, n_steps=int(n), err=err.value)

    def explicit_batch(self, n_member, y0, target_times, min_dt=0.05, su_table=None,
                       want_soln=True, n_thread=0):
        tt = _f64(target_times)
        y0, y0_stride, params, p_stride = self._batch_inputs(n_member, y0)
        soln = np.zeros((n_member, len(tt), self.C)) if want_soln else None
        final = np.zeros((n_member, self.C))
        err = C.c_int32(0)
        su = None if su_table is None else _f64(su_table)
        n = lib().pdo_explicit_batch(C.byref(self.struct), n_member, _ptr(y0), y0_stride,
                                     _ptr(params), p_stride, len(tt), _ptr(tt), float(min_dt),
                                     _ptr(su), _ptr(soln), _ptr(final), int(n_thread),
                                     C.byref(err))
        return dict(soln=soln, final=final, n_steps=int(n), err=err.value)


class Rng(object):
    """Uniform source for the oracle: 'numpy' / 'python' MT19937 seeding, an injected stream, or
    the Philox stream of a given (seed, member)."""

    def __init__(self, kind, seed=0, member=0, uniforms=None, stride=1):
        self.struct = PdoRng()
        self._keep = None
        L = lib()
        if kind == 'numpy':
            L.pdo_rng_seed_numpy(C.byref(self.struct), int(seed))
        elif kind == 'python':
            L.pdo_rng_seed_python(C.byref(self.struct), int(seed))
        elif kind == 'inject':
            self._keep = uniforms
            assert uniforms.dtype == np.float64
            n = (uniforms.size - 1) // stride + 1 if uniforms.size else 0
            L.pdo_rng_inject(C.byref(self.struct), _ptr(uniforms), n, stride)
        elif kind == 'philox':
            L.pdo_rng_philox(C.byref(self.struct), int(seed), int(member))
        else:
            raise ValueError(kind)

    def random(self):
        return lib().pdo_rng_double(C.byref(self.struct))

    def poisson(self, lam):
        return int(lib().pdo_poisson(C.byref(self.struct), float(lam)))

    @property
    def consumed(self):
        return int(self.struct.cursor)

    @property
    def exhausted(self):
        return bool(self.struct.exhausted)


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().pdo_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
    return out


def explicit_schedule(target_times, min_dt=0.05):
    tt = _f64(target_times)
    n = lib().pdo_explicit_schedule(len(tt), _ptr(tt), float(min_dt), None, None, None, 0)
    t_eval = np.zeros(n)
    dt = np.zeros(n)
    interval = np.zeros(n, dtype=np.int32)
    lib().pdo_explicit_schedule(len(tt), _ptr(tt), float(min_dt), _ptr(t_eval), _ptr(dt),
                                _ptr(interval), n)
    return t_eval, dt, interval


def pick_event(intervals, u):
    a = _f64(intervals)
    return int(lib().pdo_pick_event(_ptr(a), len(a), float(u)))
