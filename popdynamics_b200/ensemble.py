# -*- coding: utf-8 -*-
"""
Ensemble engine: runs batches of independent ``BaseModel`` members on one B200 through the C-ABI.

``integrate_single``   -- what ``BaseModel.integrate`` calls: one member, full trajectory; the
                          drop-in replacement of the loops at /root/reference/basepop.py:658-852.
``integrate_ensemble`` -- additive API (the reference's ensembles are Python ``for`` loops, e.g.
                          /root/reference/examples/stochastic_strains_model.py:210-218): M members,
                          per-member parameter / initial-state sweeps, output as full trajectories
                          ``[T][C][M]``, final states ``[C][M]`` or streaming statistics (exact
                          integer moments + histograms -> mean / variance / quantiles).

Buffers may be numpy arrays (host: the library stages them, this is the end-to-end path) or torch
CUDA tensors (resident in HBM, no copies).  Multi-GPU sharding lives in ``distributed.py``.
"""

import ctypes
import math
import random
import re

import numpy

from . import capi, compiler

_KERNEL_CACHE = {}
_SCHEDULE_CACHE = {}


def _kernel_for(cm, method, out_mode, rng_mode, device, count_bits, block, min_blocks=0,
                flags=0):
    source = cm.cuda_source()
    key = (source, method, out_mode, rng_mode, device, count_bits, block, min_blocks, flags)
    kernel = _KERNEL_CACHE.get(key)
    if kernel is None:
        kernel = capi.Model(source, method, out_mode, rng_mode, device, cm.n_comp, cm.n_param,
                            len(cm.member_param_slots()), cm.n_su, cm.n_flow,
                            count_bits=count_bits, block=block, min_blocks=min_blocks,
                            flags=flags)
        _KERNEL_CACHE[key] = kernel
    return kernel


def clear_kernel_cache():
    for kernel in _KERNEL_CACHE.values():
        kernel.close()
    _KERNEL_CACHE.clear()


def _schedule(target_times, min_dt):
    key = (tuple(float(t) for t in target_times), float(min_dt))
    hit = _SCHEDULE_CACHE.get(key)
    if hit is None:
        if len(_SCHEDULE_CACHE) > 64:
            _SCHEDULE_CACHE.clear()
        hit = capi.explicit_schedule(target_times, min_dt)
        _SCHEDULE_CACHE[key] = hit
    return hit


def _scaleup_table(model, cm, t_eval):
    """Values of the model's scale-up closures at every sub-step time: they depend on time only
    (/root/reference/basepop.py:544-551) and the explicit time grid is the same for every member,
    so the ORIGINAL Python closures are evaluated on the host -- exact by construction."""
    if not cm.su_labels:
        return None
    table = numpy.empty((len(t_eval), len(cm.su_labels)), dtype=numpy.float64)
    for k, label in enumerate(cm.su_labels):
        fn = model.scaleup_fns[label]
        table[:, k] = [fn(t) for t in t_eval.tolist()]
    return table


def _is_torch(x):
    return hasattr(x, 'data_ptr')


def _member_param_block(cm, n_member, like=None):
    slots = cm.member_param_slots()
    if not slots:
        return None
    block = numpy.empty((len(slots), n_member), dtype=numpy.float64)
    for j, k in enumerate(slots):
        values = cm.param_values[k]
        if values.size != n_member:
            raise ValueError('parameter %r has %d entries for %d members'
                             % (cm.param_labels[k], values.size, n_member))
        block[j] = values
    return block


def _member_param_rows(cm, n_member):
    """The swept parameter columns as the model holds them (no [P][M] copy): list of [M] float64
    arrays in slot order, or None."""
    slots = cm.member_param_slots()
    if not slots:
        return None
    rows = []
    for k in slots:
        values = cm.param_values[k]
        if values.size != n_member:
            raise ValueError('parameter %r has %d entries for %d members'
                             % (cm.param_labels[k], values.size, n_member))
        rows.append(values)
    return rows


def _initial_state(model, y, n_member):
    """[C] shared or [C][M] per-member initial compartments."""
    per_member = any(isinstance(v, numpy.ndarray) for v in y)
    if not per_member:
        return numpy.asarray([float(v) for v in y], dtype=numpy.float64), 0
    block = numpy.empty((len(y), n_member), dtype=numpy.float64)
    for c, v in enumerate(y):
        block[c] = v
    return block, 1


class EnsembleResult(object):
    """Outputs of one ensemble launch (arrays are numpy unless torch buffers were supplied)."""

    def __init__(self):
        self.method = None
        self.output = None
        self.labels = []
        self.times = None
        self.n_members = 0        # members integrated by THIS launch (the shard)
        self.n_total = None       # members behind the statistics once shards have been reduced
        self.member_offset = 0
        self.seed = None
        self.soln = None          # [T][C][M]
        self.final = None         # [C][M]
        self.final_vars = None    # [V][M]: diagnostics vars of the final state (output='final_vars')
        self.final_var_labels = []
        self.moments = None       # uint64 [T][C][2]
        self.hist = None          # uint32 [rows][C][bins]
        self.hist_times = None    # target index of every histogram row
        self.hist_shift = None
        self.hist_bins = 0
        self.device = 0
        self.info = {}

    # ---- statistics -------------------------------------------------------------------
    def _moments_host(self):
        m = self.moments
        if _is_torch(m):
            m = m.cpu().numpy()
        return m.astype(numpy.float64) if m.dtype != numpy.float64 else m

    def mean(self, n_total=None):
        n = float(n_total or self.n_total or self.n_members)
        return self._moments_host()[..., 0] / n

    def var(self, n_total=None, ddof=0):
        n = float(n_total or self.n_total or self.n_members)
        m = self._moments_host()
        mean = m[..., 0] / n
        return (m[..., 1] / n - mean * mean) * (n / (n - ddof))

    def quantiles(self, q):
        """[rows][C][len(q)] int64: inverted-CDF quantiles from the histograms (exact when the
        bin width is 1, i.e. ``hist_shift == 0``, and nothing fell into the overflow bin)."""
        if self.hist is None:
            raise ValueError('no histograms were recorded')
        n_row = len(self.hist_times)
        C = len(self.labels)
        if _is_torch(self.hist):
            import torch
            out = torch.empty((n_row, C, len(q)), dtype=torch.int64, device=self.hist.device)
        else:
            out = numpy.empty((n_row, C, len(q)), dtype=numpy.int64)
        capi.hist_quantiles(self.hist, n_row, C, self.hist_bins, self.hist_shift, q, out,
                            device=self.device)
        return out

    def derive_moments(self):
        """Exact moments from lossless histograms (``moments_from='hist'``); call again after the
        histograms of several shards / ranks have been added up."""
        capi.hist_moments(self.hist, len(self.hist_times), len(self.labels), self.hist_bins,
                          self.moments, device=self.device)
        return self.moments

    def final_var(self, label):
        """[M] values of one fused final-state diagnostics var (``output='final_vars'``)."""
        return self.final_vars[self.final_var_labels.index(label)]

    def trajectory(self, member):
        """(T, C) array of one member, the layout of the reference's ``soln_array``."""
        s = self.soln[:, :, member]
        return s.cpu().numpy() if _is_torch(s) else numpy.array(s)


def _hist_setup(n_time, n_comp, hist_bins, hist_max, hist_times):
    if hist_bins % 2:
        raise ValueError('hist_bins must be even')
    if hist_times is None:
        hist_times = list(range(n_time))
    hist_times = [int(t) % n_time for t in hist_times]
    row = numpy.full(n_time, -1, dtype=numpy.int32)
    for r, t in enumerate(hist_times):
        row[t] = r
    hist_max = numpy.broadcast_to(numpy.asarray(hist_max, dtype=numpy.float64), (n_comp,))
    shift = numpy.zeros(n_comp, dtype=numpy.int32)
    for c in range(n_comp):
        need = (hist_max[c] + 1.0) / hist_bins
        shift[c] = max(0, int(math.ceil(math.log2(need)))) if need > 1 else 0
    return row, shift, numpy.asarray(hist_times, dtype=numpy.int32)


def run_compiled(model, cm, method, y, target_times, n_members, output='full', seed=0,
                 member_offset=0, uniforms=None, device=0, count_bits=64, block=256,
                 hist_bins=0, hist_max=None, hist_times=None, out=None, moments=None,
                 hist=None, member_params=None, y0=None, stream=None, min_dt=None,
                 min_blocks=0, statistics='host', mask=None, mask_value=0, moments_from='kernel',
                 rtol=1.49012e-8, atol=1.49012e-8, first_step=0.0, max_steps=10 ** 7):
    """Launch one compiled model.  Returns an EnsembleResult.  ``out`` / ``moments`` / ``hist`` /
    ``member_params`` / ``y0`` let the caller supply resident (torch CUDA) buffers.
    ``statistics='device'`` keeps the streaming-statistics accumulators (moments, histograms)
    in HBM as torch tensors created here: only what ``mean`` / ``var`` / ``quantiles`` return
    ever crosses PCIe, not the histograms themselves.  ``mask`` ([M] uint8) restricts the launch to
    the members with ``mask[m] == mask_value``; the outputs of the others are left untouched.
    ``moments_from='hist'``: with lossless histograms (width-1 bins at every target time) the exact
    moments are derived from them after the launch instead of being reduced inside the kernel; a
    count that reaches the last bin raises the device error flag (choose a larger ``hist_max``).
    ``rtol`` / ``atol`` / ``first_step`` / ``max_steps`` steer the adaptive integrator
    (``method='scipy'``; the defaults are scipy.integrate.odeint's tolerances)."""
    method_code = capi.METHOD_CODES[method]
    out_code = capi.OUTPUT_CODES[output]
    rng_code = capi.INJECT if uniforms is not None else capi.PHILOX
    live_rows = cm.n_flow
    if method == 'discrete_time_stochastic':
        # model-size limits of the tau-leap kernel, reported here instead of as an NVRTC log
        live_rows = int(re.search(r'#define PD_LIVE_NROW (\d+)', cm.cuda_source()).group(1))
        if live_rows > 64:
            raise ValueError('the tau-leap kernel handles at most 64 flow-table rows that can '
                             'fire (its live-row mask is one 64-bit word); this model has %d'
                             % live_rows)
    if output == 'stats' and cm.n_comp > 64:
        raise ValueError("output='stats' supports at most 64 compartments (this model has %d); "
                         "use output='full' or 'final'" % cm.n_comp)
    if not min_blocks and ((method == 'discrete_time_stochastic' and live_rows <= 24) or
                           (method == 'continuous_time_stochastic' and cm.n_flow <= 16)):
        # measured on B200 (profiles/): the stochastic kernels are fastest at 64 registers per
        # thread, i.e. 1024 resident threads per SM, whatever the block size; kernels of a large
        # flow table (TB Gillespie: 20 rows, ~100 registers) are better left uncapped, and a
        # tau-leap draw list beyond two dozen rows would only spill under the cap
        min_blocks = max(1, 1024 // int(block))
    if method == 'continuous_time_stochastic' and cm.n_flow > 16 and block == 256 and \
            output == 'final':
        # a large flow table keeps ~100 registers per thread alive (rates and cumulative
        # intervals of every row): five 128-thread blocks at 94 registers fit an SM instead of
        # two of 256 (TB model, B200: 2.84e10 against 2.51e10 events/s,
        # profiles/r2_tune_gillespie.txt)
        block = 128
        min_blocks = min_blocks or 5
    if method in ('explicit', 'scipy') and block == 256:
        # measured on B200, SEIR-demography sweep, 10^6 members: 128-thread blocks are 1.3 %
        # faster than 256 for both ODE kernels (2.825e11 against 2.789e11 Euler steps/s)
        block = 128
    kernel = _kernel_for(cm, method_code, out_code, rng_code, device, count_bits, block,
                         min_blocks,
                         (capi.FLAG_MEMBER_MASK if mask is not None else 0) |
                         (capi.FLAG_HIST_MOMENTS if (moments_from == 'hist' and output == 'stats')
                          else 0))
    C, T, M = cm.n_comp, len(target_times), int(n_members)
    tt = numpy.ascontiguousarray([float(t) for t in target_times], dtype=numpy.float64)

    res = EnsembleResult()
    res.method, res.output, res.labels = method, output, list(cm.labels)
    res.times, res.n_members, res.member_offset, res.seed = tt, M, member_offset, seed
    res.device = device

    if y0 is None:
        y0, y0_per_member = _initial_state(model, y, M)
    else:
        y0_per_member = 1 if (y0.ndim == 2) else 0
    param_rows = None
    if method in ('discrete_time_stochastic', 'continuous_time_stochastic') and count_bits == 32:
        # 32-bit counts: an initial state at or above 2**31 would be silently saturated by the
        # double -> int conversion on the device; numpy / Python ints never wrap
        peak = float(y0.max()) if (y0.numel() if _is_torch(y0) else y0.size) else 0.0
        if not peak < 2.0 ** 31:
            raise ValueError('initial compartment %.17g does not fit count_bits=32; use '
                             'count_bits=64' % peak)
    if member_params is None:
        # the swept columns go to the library as they are: one pointer per parameter, staged
        # chunk by chunk inside the run (no [P][M] host copy)
        rows = _member_param_rows(cm, M)
        if rows is not None:
            param_rows = capi.row_pointers(rows)
    uniform = cm.uniform_params()
    keep = [tt, y0, member_params, uniform, param_rows]

    def alloc(shape, dtype):
        return numpy.empty(shape, dtype=dtype)

    if method == 'explicit':
        run = capi.ExplicitRun()
        min_dt = model.explicit_min_dt if min_dt is None else min_dt
        t_eval, dt, n_sub = _schedule(target_times, min_dt)
        su = _scaleup_table(model, cm, t_eval)
        keep += [t_eval, dt, n_sub, su]
        run.n_member = M
        run.y0, run.y0_per_member, run.n_time = capi.address(y0), y0_per_member, T
        run.uniform_params = capi.address(uniform)
        run.member_params = capi.address(member_params)
        run.member_param_rows = param_rows[1] if param_rows else None
        n_sub_arg = n_sub if len(n_sub) else numpy.zeros(1, dtype=numpy.int32)
        keep.append(n_sub_arg)
        run.n_sub, run.dt = capi.address(n_sub_arg), capi.address(dt) if len(dt) else None
        run.t_eval = capi.address(t_eval) if (cm.uses_time and len(t_eval)) else None
        run.scaleup_table = capi.address(su) if su is not None and len(su) else None
        if su is not None and not len(su):
            run.scaleup_table = capi.address(numpy.zeros((1, cm.n_su)))
        run.n_steps = len(dt)
        res.info['last_eval_time'] = float(t_eval[-1]) if len(t_eval) else None
    elif method == 'scipy':
        if mask is not None:
            raise ValueError("a member mask is not supported by method='scipy'")
        run = capi.AdaptiveRun()
        run.n_member, run.member0 = M, int(member_offset)
        run.y0, run.y0_per_member, run.n_time = capi.address(y0), y0_per_member, T
        run.uniform_params = capi.address(uniform)
        run.member_params = capi.address(member_params)
        run.member_param_rows = param_rows[1] if param_rows else None
        run.target_times = capi.address(tt)
        run.rtol, run.atol, run.first_step = float(rtol), float(atol), float(first_step)
        run.max_steps = int(max_steps)
        res.info['last_eval_time'] = float(tt[-1])
    else:
        run = capi.StochRun()
        run.n_member, run.member0 = M, int(member_offset)
        run.y0, run.y0_per_member, run.n_time = capi.address(y0), y0_per_member, T
        run.uniform_params = capi.address(uniform)
        run.member_params = capi.address(member_params)
        run.member_param_rows = param_rows[1] if param_rows else None
        run.target_times = capi.address(tt)
        run.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        if uniforms is not None:
            if tuple(uniforms.shape)[-1] != M:
                raise ValueError('uniforms must have shape [n_uniform][n_members]')
            run.uniforms, run.n_uniform = capi.address(uniforms), int(uniforms.shape[0])
            keep.append(uniforms)

    if output == 'full':
        res.soln = out if out is not None else alloc((T, C, M), numpy.float64)
        run.soln = capi.address(res.soln)
    elif output == 'final':
        res.final = out if out is not None else alloc((C, M), numpy.float64)
        run.final_state = capi.address(res.final)
    elif output == 'final_vars':
        V = len(cm.final_var_labels)
        if not V:
            raise ValueError("output='final_vars' needs a model compiled with final_vars=[...] "
                             "(integrate_ensemble(final_vars=[...]))")
        res.final_var_labels = list(cm.final_var_labels)
        res.final_vars = out if out is not None else alloc((V, M), numpy.float64)
        run.final_state, run.n_final_var = capi.address(res.final_vars), V
        if method == 'explicit':
            run.t_final = float(tt[-1])
    else:
        if method in ('explicit', 'scipy'):
            raise ValueError("output='stats' is defined for the stochastic integrators")
        on_device = statistics == 'device'
        if statistics not in ('host', 'device'):
            raise ValueError("statistics must be 'host' or 'device'")
        if on_device:
            import torch
            cuda = torch.device('cuda', device)
        if moments is not None:
            res.moments = moments
        elif on_device:
            res.moments = torch.zeros((T, C, 2), dtype=torch.int64, device=cuda)
        else:
            res.moments = numpy.zeros((T, C, 2), numpy.uint64)
        run.moments = capi.address(res.moments)
        if hist_bins:
            if hist_max is None:
                raise ValueError('hist_max (largest expected count per compartment) is needed '
                                 'with hist_bins; see ensemble.pilot_hist_max')
            # the second moment is an exact u64 sum of x^2 over ALL members that accumulate into
            # this buffer: refuse a launch that could overflow it on its own
            worst = float(numpy.max(numpy.asarray(hist_max, dtype=numpy.float64)))
            if float(M) * worst * worst >= 2.0 ** 64:
                raise ValueError(
                    'n_members * hist_max^2 = %.3g does not fit the exact 64-bit second moment; '
                    'run the ensemble in smaller launches with separate moment buffers and add '
                    'them as Python integers' % (float(M) * worst * worst))
            row, shift, rows = _hist_setup(T, C, hist_bins, hist_max, hist_times)
            if hist is not None:
                res.hist = hist
            elif on_device:
                res.hist = torch.zeros((len(rows), C, hist_bins), dtype=torch.int32, device=cuda)
            else:
                res.hist = numpy.zeros((len(rows), C, hist_bins), numpy.uint32)
            res.hist_times, res.hist_shift, res.hist_bins = rows, shift, hist_bins
            keep += [row, shift]
            run.hist, run.hist_row = capi.address(res.hist), capi.address(row)
            run.hist_shift = capi.address(shift)
            run.hist_bins, run.n_hist_row = hist_bins, len(rows)
        if moments_from == 'hist':
            if not hist_bins or len(rows) != T or int(numpy.max(shift)) != 0:
                raise ValueError("moments_from='hist' needs width-1 histogram bins at every "
                                 "target time (hist_bins > hist_max, no hist_times)")
            run.moments_from_hist = 1
        elif moments_from != 'kernel':
            raise ValueError("moments_from must be 'kernel' or 'hist'")
    run.stream = stream
    if mask is not None:
        run.mask, run.mask_value = capi.address(mask), int(mask_value)
        keep.append(mask)

    result = capi.run(kernel, run)
    res.info.update(result.as_dict())
    res.moments_from = moments_from if output == 'stats' else None
    if res.moments_from == 'hist' and not result.err_code:
        res.derive_moments()
    del keep
    return res


def empty_result(model, method, output='full', member_offset=0, hist_bins=0, hist_max=None,
                 hist_times=None, statistics='host', device=None, seed=0,
                 moments_from='kernel', **_ignored):
    """The result of a launch over ZERO members (a rank whose shard is empty): zeroed statistics
    of the right shape, so that the rank can still join the all-reduce; no kernel runs."""
    device = model.device if device is None else device
    res = EnsembleResult()
    res.method, res.output, res.labels = method, output, list(model.labels)
    res.times = numpy.asarray([float(t) for t in model.target_times])
    res.member_offset, res.seed, res.device = member_offset, seed, device
    res.info = dict(err_code=0, err_member=-1, err_where=0, kernel_ms=0.0, n_steps=0,
                    h2d_bytes=0, d2h_bytes=0)
    C, T = len(model.labels), len(model.target_times)
    res.moments_from = moments_from if output == 'stats' else None
    if output == 'stats':
        def zeros(shape, np_dtype, torch_dtype):
            if statistics == 'device':
                import torch
                return torch.zeros(shape, dtype=getattr(torch, torch_dtype),
                                   device=torch.device('cuda', device))
            return numpy.zeros(shape, np_dtype)
        res.moments = zeros((T, C, 2), numpy.uint64, 'int64')
        if hist_bins:
            row, shift, rows = _hist_setup(T, C, hist_bins, hist_max, hist_times)
            res.hist = zeros((len(rows), C, hist_bins), numpy.uint32, 'int32')
            res.hist_times, res.hist_shift, res.hist_bins = rows, shift, hist_bins
    elif output == 'final':
        res.final = numpy.zeros((C, 0))
    else:
        res.soln = numpy.zeros((T, C, 0))
    return res


def raise_device_error(info, method):
    code = info.get('err_code', 0)
    if not code:
        return
    message = '%s (member %d, target index %d)' % (
        capi.DEVICE_ERRORS.get(code, 'device error %d' % code), info.get('err_member', -1),
        info.get('err_where', 0))
    if code == 1:
        raise AssertionError(message)          # checks(), basepop.py:1023-1025
    if code == 2:
        raise ValueError(message)              # numpy.random.poisson(lam < 0)
    if code == 3:
        raise ZeroDivisionError(message)       # basepop.py:776
    if code == 5:
        raise ValueError('math domain error: ' + message)
    raise RuntimeError(message)              # 4 stream exhausted, 6 count range, 7 / 8 adaptive


def integrate_single(model, y, method, is_continue, **options):
    """One member, full trajectory: the body of ``BaseModel.integrate``.  ``options`` go to
    ``run_compiled`` (``min_dt`` for the explicit, ``rtol`` / ``atol`` for the adaptive method)."""
    cm = compiler.compile_model(model, method, is_continue=is_continue)
    seed = model.seed if model.seed is not None else random.getrandbits(63)
    # With a fixed model.seed every integrate(is_continue=True) segment gets its OWN Philox
    # stream -- stream id = number of the segment -- instead of replaying the first one: the
    # reference's global generators simply keep going across segments (examples/mix_model.py:
    # 123-134), so its segments are independent.
    segment = getattr(model, '_segment', 0) + 1 if is_continue else 0
    model._segment = segment
    res = run_compiled(model, cm, method, y, model.target_times, 1, output='full', seed=seed,
                       member_offset=segment, device=model.device, **options)
    raise_device_error(res.info, method)
    info = dict(res.info)
    info['soln'] = numpy.ascontiguousarray(res.soln[:, :, 0])
    info['seed'] = seed
    info['stream'] = segment
    return info


def integrate_ensemble(model, method='explicit', n_members=None, output='full', seed=0,
                       member_offset=0, uniforms=None, device=None, count_bits=64, block=256,
                       hist_bins=0, hist_max=None, hist_times=None, check=True, min_blocks=0,
                       statistics='host', continue_from=None, moments_from='kernel',
                       final_vars=None, **buffers):
    """
    Run ``n_members`` independent members of ``model`` (additive API, see module docstring).

    ``continue_from=previous_result`` is ``integrate(..., is_continue=True)`` (basepop.py:874-876)
    for an ensemble: every member starts from its last state in ``previous_result`` (``final``
    or the last row of ``soln``), explicit runs take the continued-run arithmetic (plain ``sum``
    fold), and the Philox streams move on to the next segment
    (stream id = segment << 40 | global member id) instead of replaying the previous one.

    Per-member sweeps: give ``set_param`` / ``set_compartment`` 1-D arrays of length M.  Native
    Philox streams are keyed by ``(seed, member_offset + i)``: results do not depend on how the
    members are split over launches or GPUs.  ``uniforms[n][M]`` switches to the injected-stream
    protocol (parity tests).  Raises the reference's exception types when a member trips the
    device error flag (AssertionError for ``checks()``).

    ``final_vars=['proportion']`` (ODE methods) makes the integrator kernel itself run the
    diagnostics pass on every member's FINAL state and return those vars, ``[V][M]``, instead of
    the compartments (``output`` becomes 'final_vars'): a calibration sweep that reads one scalar
    per member (/root/reference/examples/rmit_tb_model.py:143) moves 8 bytes per member off the
    device and needs no second pass.
    """
    if final_vars:
        output = 'final_vars'
    assert bool(model.target_times), 'Haven\'t set times yet'
    if method not in capi.METHOD_CODES:
        raise ValueError('unknown integration method %r' % (method,))
    model.clear_solns()
    model.check_flow_transfers()
    y = model.get_init_list()
    segment = 0
    if continue_from is not None:
        prev = continue_from
        state = prev.final if prev.final is not None else \
            (prev.soln[-1] if prev.soln is not None else None)
        if state is None:
            raise ValueError("continue_from needs a result with output='final' or 'full'")
        if _is_torch(state):
            state = state.contiguous()
        else:
            state = numpy.ascontiguousarray(state)
        buffers = dict(buffers, y0=state)
        n_members = prev.n_members if n_members is None else n_members
        segment = getattr(prev, 'segment', 0) + 1
        member_offset = prev.member_offset
    cm = compiler.compile_model(model, method, is_continue=continue_from is not None,
                                final_vars=final_vars)
    implied = cm.n_members_implied()
    for v in y:
        if isinstance(v, numpy.ndarray):
            if implied not in (None, v.size):
                raise ValueError('per-member arrays have different lengths')
            implied = v.size
    if n_members is None:
        n_members = implied if implied is not None else 1
    elif implied is not None and implied != n_members:
        raise ValueError('n_members=%d but per-member arrays have %d entries'
                         % (n_members, implied))
    device = model.device if device is None else device
    res = run_compiled(model, cm, method, y, model.target_times, n_members, output=output,
                       seed=seed, member_offset=member_offset + (segment << 40),
                       uniforms=uniforms, device=device,
                       count_bits=count_bits, block=block, hist_bins=hist_bins,
                       hist_max=hist_max, hist_times=hist_times, min_blocks=min_blocks,
                       statistics=statistics, moments_from=moments_from, **buffers)
    res.member_offset, res.segment = member_offset, segment
    if method == 'explicit' and res.info.get('err_code'):
        res.info['err_member'] += int(member_offset)     # the explicit kernel counts from 0
    if check:
        raise_device_error(res.info, method)
    return res


def integrate_hybrid(model, segments, n_members, cutoff, seed=0, member_offset=0, device=None,
                     stochastic_method='continuous_time_stochastic', output='final',
                     count_bits=64, block=256):
    """
    Per-segment switching between a stochastic integrator and explicit Euler for every member of
    an ensemble -- the loop of /root/reference/examples/mix_model.py:107-136 over
    ``integrate(method, is_continue=True)`` (basepop.py:874-876, 894-899), batched:

      * ``segments`` is a list of ``make_times`` argument tuples, one per segment;
      * the first segment is stochastic (mix_model.py:118-121); before every later segment a
        member whose smallest compartment exceeds ``cutoff`` switches to 'explicit', the others
        to ``stochastic_method`` (:126-133);
      * a segment starts from the last row of the previous one; entering a stochastic segment
        truncates the state with ``int()`` (basepop.py:747-748); explicit segments are continued
        runs, where Python's ``sum`` folds plainly over ``numpy.float64``;
      * every segment restarts its clock at its first target time, as ``integrate`` does.

    Both integrators are launched over the same state buffer with complementary member masks.
    Philox streams: stream id = segment << 40 | global member id, so segments are independent
    and the result does not depend on sharding.  Returns an EnsembleResult with ``final``
    [C][M] and, for ``output='full'``, ``soln`` [T][C][M] / ``times`` of the concatenated grid,
    plus ``modes`` [n_segments][M] (1 = explicit).  Buffers are torch CUDA tensors."""
    import torch
    if output not in ('final', 'full'):
        raise ValueError("integrate_hybrid supports output='final' or 'full'")
    device = model.device if device is None else device
    cuda = torch.device('cuda', device)
    M = int(n_members)
    model.check_flow_transfers()
    y = model.get_init_list()
    init, per_member = _initial_state(model, y, M)
    C = len(model.labels)
    state = torch.empty((C, M), dtype=torch.float64, device=cuda)
    state[:] = torch.from_numpy(numpy.ascontiguousarray(init)).to(cuda).reshape(C, -1)
    res = EnsembleResult()
    res.method, res.output, res.labels = 'hybrid', output, list(model.labels)
    res.n_members, res.member_offset, res.seed, res.device = M, member_offset, seed, device
    rows, times, modes = [], [], []
    member_blocks, compiled = {}, {}
    for k, seg in enumerate(segments):
        model.make_times(*seg)
        tt = list(model.target_times)
        T = len(tt)
        if k == 0:
            explicit = torch.zeros(M, dtype=torch.uint8, device=cuda)
        else:
            explicit = (state.min(dim=0).values > cutoff).to(torch.uint8)
        modes.append(explicit)
        n_explicit = int(explicit.sum().item())
        buf = torch.empty((T, C, M), dtype=torch.float64, device=cuda) if output == 'full' else state
        for method, value, count in ((stochastic_method, 0, M - n_explicit),
                                     ('explicit', 1, n_explicit)):
            if not count:
                continue
            if (method, k > 0) not in compiled:    # the trace does not depend on the time grid
                compiled[(method, k > 0)] = compiler.compile_model(model, method,
                                                                   is_continue=(k > 0))
            cm = compiled[(method, k > 0)]
            if method not in member_blocks:        # slot order depends on the trace
                block_ = _member_param_block(cm, M)
                member_blocks[method] = None if block_ is None else \
                    torch.from_numpy(block_).to(cuda)
            part = run_compiled(model, cm, method, y, tt, M, output=output, seed=seed,
                                member_offset=member_offset + (k << 40), device=device,
                                count_bits=count_bits, block=block, y0=state, out=buf,
                                member_params=member_blocks[method], mask=explicit,
                                mask_value=value)
            raise_device_error(part.info, method)
            res.info.setdefault('launches', []).append(
                dict(segment=k, method=method, members=count, kernel_ms=part.info['kernel_ms']))
        if output == 'full':
            rows.append(buf if k == 0 else buf[1:])
            times.extend(tt if k == 0 else tt[1:])
            state = buf[-1].clone()
    res.final = state
    res.modes = torch.stack(modes)
    if output == 'full':
        res.soln = torch.cat(rows, dim=0)
        res.times = numpy.asarray([float(t) for t in times])
    return res


class DiagnosticsResult(object):
    """Outputs of the batched diagnostics pass: ``var_labels`` in the reference's order,
    ``vars`` [T][V][M] (``var_array`` of every member), ``flows`` [T][C][M] (``flow_array``) or
    None.  For a final-state input T is 1 and the arrays are [V][M] / [C][M]."""

    def __init__(self):
        self.var_labels, self.labels, self.times = [], [], None
        self.vars, self.flows, self.info = None, None, {}

    def var(self, label):
        """[T][M] (or [M]) slice of one var, e.g. ``var('proportion')``."""
        k = self.var_labels.index(label)
        return self.vars[:, k] if self.vars.ndim == 3 else self.vars[k]

    def fraction(self, label, soln):
        """``fraction_soln[label]`` of every member (basepop.py:954-960): compartment over
        vars['population']."""
        c = self.labels.index(label)
        return soln[..., c, :] / self.var('population')


def diagnostics(model, result, method=None, want_flows=False, device=None, block=256,
                want_vars=None, out=None):
    """The diagnostics pass (``calculate_diagnostics``, /root/reference/basepop.py:911-960) for
    every member of an ensemble result: ``calculate_vars`` / ``calculate_flows`` /
    ``calculate_diagnostic_vars`` traced once and evaluated on the device at every stored time
    of ``result.soln`` ([T][C][M]) or on ``result.final`` ([C][M]).  ``model`` must still carry
    the parameters (incl. per-member arrays) the result was produced with.

    Scale-up vars: for the explicit integrator the reference never re-evaluates them in this pass,
    they keep the value of the last derivative call -- reproduced here; a hook
    that calls ``calculate_scaleup_vars()`` itself gets them at the stored time.

    ``want_vars=['proportion']`` evaluates and stores only the named vars (BASELINE config 5 reads
    one scalar per member, /root/reference/examples/rmit_tb_model.py:143): 8 bytes per member and
    time leave the device instead of 8 V.  ``out`` supplies the [T][V][M] (or [V][M]) buffer."""
    method = method or result.method
    device = result.device if device is None else device
    model.check_flow_transfers()
    cm = compiler.compile_model(model, method, diagnostics=True, want_vars=want_vars)
    if not cm.diag_nodes:
        raise ValueError('the model defines no vars')
    soln = result.soln if result.soln is not None else result.final
    if soln is None:
        raise ValueError("diagnostics need output='full' or output='final'")
    final_only = result.soln is None
    M = int(result.n_members)
    C, V = cm.n_comp, len(cm.diag_nodes)
    times = numpy.ascontiguousarray(result.times[-1:] if final_only else result.times,
                                    dtype=numpy.float64)
    T = len(times)
    source = cm.cuda_source()
    key = (source, capi.DIAGNOSTICS, device, block)
    kernel = _KERNEL_CACHE.get(key)
    if kernel is None:
        kernel = capi.Model(source, capi.DIAGNOSTICS, capi.FULL, capi.PHILOX, device, C,
                            cm.n_param, len(cm.member_param_slots()), cm.n_su, cm.n_flow,
                            block=block)
        _KERNEL_CACHE[key] = kernel
    out_buffer, out = out, DiagnosticsResult()
    out.var_labels, out.labels, out.times = list(cm.diag_labels), list(cm.labels), times

    def like(shape):
        if _is_torch(soln):
            import torch
            return torch.empty(shape, dtype=torch.float64, device=soln.device)
        return numpy.empty(shape, dtype=numpy.float64)

    if out_buffer is not None:
        out.vars = out_buffer.reshape((T, V, M))
    else:
        out.vars = like((T, V, M))
    out.flows = like((T, C, M)) if want_flows else None
    run = capi.DiagRun()
    run.n_member, run.n_time, run.n_var = M, T, V
    run.soln = capi.address(soln)
    uniform = cm.uniform_params()
    member_params = None
    rows = _member_param_rows(cm, M)
    if rows is not None:
        # only the swept columns the requested outputs read are staged (config 5's `proportion`
        # reads none: 8 bytes per member leave the device and nothing but the state goes in)
        used = compiler.params_read_by(cm, list(cm.diag_nodes) +
                                       (list(cm.diag_flow_nodes) if want_flows else []))
        rows = [r if k in used else None for k, r in zip(cm.member_param_slots(), rows)]
    param_rows = capi.row_pointers(rows) if rows is not None else None
    su_row = None
    if cm.n_su:
        t_last = result.info.get('last_eval_time')
        if t_last is None:
            raise ValueError('the result does not record the time of the last derivative call')
        su_row = numpy.array([model.scaleup_fns[label](t_last) for label in cm.su_labels],
                             dtype=numpy.float64)
    keep = [uniform, member_params, su_row, times, param_rows]
    run.uniform_params = capi.address(uniform)
    run.member_params = capi.address(member_params)
    run.member_param_rows = param_rows[1] if param_rows else None
    run.times = capi.address(times)
    run.scaleup_row = capi.address(su_row)
    run.var_soln = capi.address(out.vars)
    run.flow_soln = capi.address(out.flows)
    res = capi.run(kernel, run)
    out.info = res.as_dict()
    del keep
    if final_only:
        out.vars = out.vars[0]
        out.flows = out.flows[0] if want_flows else None
    return out


def pilot_hist_max(model, method, n_pilot=2048, seed=12345, margin=1.5, device=None):
    """Largest count per compartment over a small pilot ensemble, times ``margin``: a bound for
    ``hist_max`` when nothing better is known."""
    res = integrate_ensemble(model, method, n_members=n_pilot, output='full', seed=seed,
                             device=device)
    return numpy.ceil(res.soln.max(axis=(0, 2)) * margin)
