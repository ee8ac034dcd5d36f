# -*- coding: utf-8 -*-
"""
ctypes binding of the C-ABI library ``libpopdyn_b200.so`` (include/popdyn_b200.h).

Buffers are passed as raw addresses: ``numpy.ndarray`` (host) and ``torch.Tensor`` (host or CUDA)
are both accepted -- the library inspects each pointer and stages host memory itself.  The
library must be present: a missing or unloadable extension raises ``ExtensionMissing`` -- there
is no Python or CPU substitute for the kernels.
"""

import ctypes as C
import os

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libpopdyn_b200.so')

EXPLICIT, TAU_LEAP, GILLESPIE, DIAGNOSTICS, ADAPTIVE = 0, 1, 2, 3, 4
FULL, FINAL, STATS, FINAL_VARS = 0, 1, 2, 3
PHILOX, INJECT = 0, 1
METHOD_CODES = {'explicit': EXPLICIT, 'discrete_time_stochastic': TAU_LEAP,
                'continuous_time_stochastic': GILLESPIE, 'scipy': ADAPTIVE}
OUTPUT_CODES = {'full': FULL, 'final': FINAL, 'stats': STATS, 'final_vars': FINAL_VARS}

PD_E_NO_DEVICE, PD_E_MODEL = -4, -5
DEVICE_ERRORS = {
    1: 'checks() assertion failed on the device',
    2: 'negative or NaN Poisson mean (numpy.random.poisson would raise ValueError)',
    3: 'total event rate is zero while events exist (ZeroDivisionError in the reference)',
    4: 'injected uniform stream exhausted',
    5: 'log(0) in the Gillespie waiting time',
    6: 'compartment count exceeds the range of the statistics accumulators (2**31)',
    7: 'adaptive integrator: step budget exhausted (raise max_steps)',
    8: 'adaptive integrator: step size underflow',
    9: 'PTRS squeeze and exact acceptance test disagree (PD_PTRS_CHECK build)',
}

EXPORTED_SYMBOLS = (
    'pd_version', 'pd_last_error', 'pd_device_count', 'pd_struct_size', 'pd_model_create', 'pd_model_destroy',
    'pd_model_cubin_size', 'pd_model_cubin_copy', 'pd_explicit_schedule', 'pd_run_explicit',
    'pd_run_tau_leap', 'pd_run_gillespie', 'pd_run_diagnostics', 'pd_run_adaptive',
    'pd_hist_quantiles',
    'pd_hist_moments', 'pd_philox_fill', 'pd_microbench', 'pd_math_selfcheck')


class ExtensionMissing(RuntimeError):
    pass


class PopdynError(RuntimeError):
    def __init__(self, status, message):
        RuntimeError.__init__(self, 'libpopdyn_b200 status %d: %s' % (status, message))
        self.status = status


class ModelDesc(C.Structure):
    _fields_ = [(name, C.c_int32) for name in
                ('method', 'out_mode', 'rng_mode', 'count_bits', 'block', 'device', 'n_comp',
                 'n_param', 'n_member_param', 'n_scaleup', 'n_flow', 'min_blocks', 'flags',
                 'reserved')]


FLAG_MEMBER_MASK = 1
FLAG_HIST_MOMENTS = 2


class RunResult(C.Structure):
    _fields_ = [('kernel_ms', C.c_double), ('h2d_ms', C.c_double), ('d2h_ms', C.c_double),
                ('h2d_bytes', C.c_uint64), ('d2h_bytes', C.c_uint64), ('n_steps', C.c_uint64),
                ('err_code', C.c_int32), ('err_where', C.c_int32), ('err_member', C.c_int64),
                ('grid', C.c_int32), ('block', C.c_int32), ('regs_per_thread', C.c_int32),
                ('smem_bytes', C.c_int32), ('blocks_per_sm', C.c_int32), ('n_launch', C.c_int32)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class StochRun(C.Structure):
    _fields_ = [('n_member', C.c_int64), ('member0', C.c_int64),
                ('y0', C.c_void_p), ('y0_per_member', C.c_int32), ('n_time', C.c_int32),
                ('uniform_params', C.c_void_p), ('member_params', C.c_void_p),
                ('target_times', C.c_void_p), ('seed', C.c_uint64),
                ('uniforms', C.c_void_p), ('n_uniform', C.c_int64),
                ('soln', C.c_void_p), ('final_state', C.c_void_p), ('moments', C.c_void_p),
                ('hist', C.c_void_p), ('hist_row', C.c_void_p), ('hist_shift', C.c_void_p),
                ('hist_bins', C.c_int32), ('n_hist_row', C.c_int32), ('stream', C.c_void_p),
                ('mask', C.c_void_p), ('mask_value', C.c_int32),
                ('moments_from_hist', C.c_int32), ('member_param_rows', C.c_void_p)]


class ExplicitRun(C.Structure):
    _fields_ = [('n_member', C.c_int64),
                ('y0', C.c_void_p), ('y0_per_member', C.c_int32), ('n_time', C.c_int32),
                ('uniform_params', C.c_void_p), ('member_params', C.c_void_p),
                ('n_sub', C.c_void_p), ('dt', C.c_void_p), ('t_eval', C.c_void_p),
                ('scaleup_table', C.c_void_p), ('n_steps', C.c_int64),
                ('soln', C.c_void_p), ('final_state', C.c_void_p), ('stream', C.c_void_p),
                ('mask', C.c_void_p), ('mask_value', C.c_int32), ('n_final_var', C.c_int32),
                ('member_param_rows', C.c_void_p), ('t_final', C.c_double)]


class DiagRun(C.Structure):
    _fields_ = [('n_member', C.c_int64), ('n_time', C.c_int32), ('n_var', C.c_int32),
                ('soln', C.c_void_p), ('uniform_params', C.c_void_p),
                ('member_params', C.c_void_p), ('times', C.c_void_p),
                ('scaleup_row', C.c_void_p), ('var_soln', C.c_void_p),
                ('flow_soln', C.c_void_p), ('stream', C.c_void_p),
                ('member_param_rows', C.c_void_p)]


class AdaptiveRun(C.Structure):
    _fields_ = [('n_member', C.c_int64), ('member0', C.c_int64),
                ('y0', C.c_void_p), ('y0_per_member', C.c_int32), ('n_time', C.c_int32),
                ('uniform_params', C.c_void_p), ('member_params', C.c_void_p),
                ('member_param_rows', C.c_void_p), ('target_times', C.c_void_p),
                ('rtol', C.c_double), ('atol', C.c_double), ('first_step', C.c_double),
                ('max_steps', C.c_int64), ('soln', C.c_void_p), ('final_state', C.c_void_p),
                ('stream', C.c_void_p), ('n_final_var', C.c_int32), ('reserved', C.c_int32)]


_lib = None


def lib():
    """The loaded C-ABI library (built in-tree by popdynamics_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ExtensionMissing(
            '%s is missing: build it with `python -m popdynamics_b200.build` '
            '(popdynamics_b200 has no CPU fallback)' % LIB_PATH)
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as exc:
        raise ExtensionMissing('cannot load %s: %s' % (LIB_PATH, exc))
    P = C.POINTER
    L.pd_version.restype = C.c_int
    L.pd_last_error.restype = C.c_char_p
    L.pd_device_count.restype = C.c_int
    L.pd_model_create.argtypes = [C.c_char_p, P(ModelDesc), P(C.c_void_p)]
    L.pd_model_destroy.argtypes = [C.c_void_p]
    L.pd_model_destroy.restype = None
    L.pd_model_cubin_size.argtypes = [C.c_void_p, P(C.c_size_t)]
    L.pd_model_cubin_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.pd_explicit_schedule.argtypes = [C.c_int32, C.c_void_p, C.c_double, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int64, P(C.c_int64)]
    L.pd_run_explicit.argtypes = [C.c_void_p, P(ExplicitRun), P(RunResult)]
    L.pd_run_tau_leap.argtypes = [C.c_void_p, P(StochRun), P(RunResult)]
    L.pd_run_gillespie.argtypes = [C.c_void_p, P(StochRun), P(RunResult)]
    L.pd_run_diagnostics.argtypes = [C.c_void_p, P(DiagRun), P(RunResult)]
    L.pd_run_adaptive.argtypes = [C.c_void_p, P(AdaptiveRun), P(RunResult)]
    L.pd_hist_quantiles.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    L.pd_hist_moments.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_int32, C.c_void_p]
    L.pd_philox_fill.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_uint64,
                                 C.c_int32, C.c_void_p]
    L.pd_microbench.argtypes = [C.c_int32, C.c_int32, P(C.c_double)]
    L.pd_math_selfcheck.argtypes = [C.c_int32, C.c_int64, C.c_uint64, P(C.c_uint64), P(C.c_double)]
    _lib = L
    return L


def check(status):
    if status != 0:
        raise PopdynError(status, lib().pd_last_error().decode(errors='replace'))


def address(buf):
    """Raw address of a numpy array or torch tensor (host or device); None -> NULL."""
    if buf is None:
        return None
    if isinstance(buf, numpy.ndarray):
        if not buf.flags['C_CONTIGUOUS']:
            raise ValueError('buffer must be C-contiguous')
        return buf.ctypes.data
    if hasattr(buf, 'data_ptr'):
        if not buf.is_contiguous():
            raise ValueError('tensor must be contiguous')
        return buf.data_ptr()
    raise TypeError('expected numpy.ndarray or torch.Tensor, got %s' % type(buf).__name__)


def row_pointers(buffers):
    """HOST array of row addresses for ``member_param_rows``: one pointer per swept parameter
    column (numpy or torch, each [M] float64).  Returns (keep-alive object, address)."""
    table = (C.c_void_p * len(buffers))(*[address(b) for b in buffers])
    return table, C.addressof(table)


def device_count():
    return lib().pd_device_count()


class Model(object):
    """One NVRTC-specialised kernel (``pd_model``)."""

    def __init__(self, header_source, method, out_mode, rng_mode, device, n_comp, n_param,
                 n_member_param, n_scaleup, n_flow, count_bits=64, block=256, min_blocks=0,
                 flags=0):
        self.desc = ModelDesc(method, out_mode, rng_mode, count_bits, block, device, n_comp,
                              n_param, n_member_param, n_scaleup, n_flow, min_blocks, flags, 0)
        handle = C.c_void_p()
        check(lib().pd_model_create(header_source.encode(), C.byref(self.desc), C.byref(handle)))
        self.handle = handle

    def cubin(self):
        size = C.c_size_t()
        check(lib().pd_model_cubin_size(self.handle, C.byref(size)))
        buf = C.create_string_buffer(size.value)
        check(lib().pd_model_cubin_copy(self.handle, buf, size.value))
        return buf.raw

    def close(self):
        if getattr(self, 'handle', None):
            lib().pd_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def explicit_schedule(target_times, min_dt=0.05):
    """Host replay of integrate_explicit's sub-step schedule (/root/reference/basepop.py:672-682),
    native version of ``compiler.explicit_schedule``.  Returns (t_eval, dt, n_sub)."""
    tt = numpy.ascontiguousarray(target_times, dtype=numpy.float64)
    n_sub = numpy.zeros(max(len(tt) - 1, 1), dtype=numpy.int32)
    n = C.c_int64()
    check(lib().pd_explicit_schedule(len(tt), tt.ctypes.data, float(min_dt), None, None,
                                     n_sub.ctypes.data, 0, C.byref(n)))
    t_eval = numpy.zeros(n.value, dtype=numpy.float64)
    dt = numpy.zeros(n.value, dtype=numpy.float64)
    check(lib().pd_explicit_schedule(len(tt), tt.ctypes.data, float(min_dt), t_eval.ctypes.data,
                                     dt.ctypes.data, n_sub.ctypes.data, n.value, C.byref(n)))
    return t_eval, dt, n_sub[:max(len(tt) - 1, 0)]


def run(model, run_struct):
    """Launch; returns the RunResult.  A raised device error flag comes back as PD_E_MODEL: the
    caller decides which Python exception it maps to."""
    res = RunResult()
    fn = {EXPLICIT: lib().pd_run_explicit, TAU_LEAP: lib().pd_run_tau_leap,
          GILLESPIE: lib().pd_run_gillespie, ADAPTIVE: lib().pd_run_adaptive,
          DIAGNOSTICS: lib().pd_run_diagnostics}[model.desc.method]
    status = fn(model.handle, C.byref(run_struct), C.byref(res))
    if status != 0 and status != PD_E_MODEL:
        check(status)
    return res


def hist_quantiles(hist, n_row, n_comp, n_bins, hist_shift, q, out, device=0, stream=None):
    shift = numpy.ascontiguousarray(hist_shift, dtype=numpy.int32)
    q = numpy.ascontiguousarray(q, dtype=numpy.float64)
    check(lib().pd_hist_quantiles(address(hist), n_row, n_comp, n_bins, shift.ctypes.data,
                                  q.ctypes.data, len(q), address(out), device, stream))
    return out


def hist_moments(hist, n_row, n_comp, n_bins, moments, device=0, stream=None):
    """moments[n_row][n_comp][2] = (sum x, sum x^2) from width-1 histograms (overwritten)."""
    check(lib().pd_hist_moments(address(hist), n_row, n_comp, n_bins, address(moments), device,
                                stream))
    return moments


def philox_fill(out, n, n_member, member0, seed, device=0, stream=None):
    check(lib().pd_philox_fill(address(out), n, n_member, member0, seed, device, stream))
    return out


def math_selfcheck(n, seed=0, device=0):
    """(mismatch counts, one offending argument each) of pd_math.cuh against the math library."""
    bad = (C.c_uint64 * 4)()
    first = (C.c_double * 4)()
    check(lib().pd_math_selfcheck(device, n, seed, bad, first))
    return list(bad), list(first)


def microbench(kind, device=0):
    g = C.c_double()
    check(lib().pd_microbench(kind, device, C.byref(g)))
    return g.value
