#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
Headline benchmark: trajectory-compartment-steps/sec of the multi-strain SIR ensemble
(BASELINE.json config 3: examples/stochastic_strains_model.py, discrete-time stochastic,
10^7 members per GPU, make_times(0, 1000, 1), ensemble mean / variance / quantiles), plus the
other BASELINE.json configurations at their stated sizes as a ``configs`` list in the same line.

    python bench.py --gpus N --steps K --warmup W            # this framework, N GPUs (torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the reference's loop on host cores

One "step" = one pass of the hot path over the whole ensemble: M members x 1000 tau-leap steps,
statistics fused, plus (N > 1) the NCCL all-reduce of the integer histograms and the moment /
quantile kernels.  ``value`` is measured with the statistics buffers resident in HBM, ``e2e``
through the public Python API (``BaseModel.integrate_ensemble``) with host numpy buffers.

``configs`` (outside the headline's timed region, one entry each with its own ms / value / e2e /
roofline / cpu_baseline / parity spot-check; ``--configs headline`` skips them):
  2  SEIR with demography, 10^6-member parameter sweep x 50 years, explicit fp64
  4  TB model, Gillespie, 1950 -> 2000, 125 000 members per GPU (10^6 on 8 GPUs)
  5  RMIT TB calibration sweep, 10^8 members over the N ranks, output = proportion per member
  5b mix_model hybrid stochastic / deterministic switching, 2^20 members per GPU
  6  adaptive (odeint-equivalent) solver on the SEIR-demography sweep
``strong_scaling``: the headline with 10^7 members in total over the N ranks.
Prints ONE JSON line on rank 0.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'trajectory_compartment_steps_per_sec'
T_END = 1000
QUANTILES = [0.025, 0.25, 0.5, 0.75, 0.975]
HIST_BINS = 2048
HIST_MAX = 2047          # strains counts stay well below 2048 (S ~ 1000-2000): width-1 bins
SM_COUNT = 148
LANES_PER_SM_PER_CLK = 128          # 4 schedulers x 1 warp instruction x 32 lanes
# SURVEY.md 8(d)'s back-of-envelope unit for the strains tau-leap step (S Poisson draws x (one exp
# ~ 25-30 fp64 ops + (lambda + 1) uniforms x ~45 int32 Philox ops) ~ 300 fp64 + ~1000 int32): kept
# as a NAMED SIDE VALUE (`frac_survey_unit`).  `roofline.frac` itself is counter based: useful
# (predicated-on) thread-instructions the kernel executes per member-step, from the committed ncu
# capture below, over the SM issue peak.
SURVEY_OPS_PER_MEMBER_STEP = 1300.0
NCU_UNIT_FILE = os.path.join(ROOT, 'profiles', 'r2_tau_leap_headline_ncu.json')
# fp64 operations per derivative evaluation, SURVEY.md 8(d): V + E + 3(Fv+Ff+D) + 3C + 2C + ~3
EXPLICIT_OPS = {'seir': 39.0, 'rmit': 84.0}


def workload_name(members_per_gpu):
    return ('stochastic_strains_model DTS (clamped Poisson tau-leap), make_times(0,1000,1), C=5, '
            '%d members per GPU, streaming %d-bin width-1 histograms -> exact mean/var + 5 '
            'quantiles per (t,c)' % (members_per_gpu, HIST_BINS))


def make_model():
    from examples import configs
    return configs.strains()


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') \
        else (os.cpu_count() or 1)


# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.device_index), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_max, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); sm_max.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(sm_max)),
                'power_w_max': float(max(power)), 'samples': len(sm),
                'reasons': sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as handle:
            return json.load(handle), 'measured'
    return {'hbm_gbs': 6650.0}, 'fallback'


def ncu_unit():
    """Per-member-step unit costs of pd_tau_leap_kernel from the committed ncu capture
    (written by tools/ncu_summary.py --json): useful thread-instructions and DRAM bytes."""
    if not os.path.exists(NCU_UNIT_FILE):
        return None
    with open(NCU_UNIT_FILE) as handle:
        return json.load(handle)


# ---------------------------------------------------------------------------------------------
# CPU arms.  The oracle (oracle/) is test infrastructure: it is executed here ONLY as the CPU
# baseline, never on the measured GPU path.
# ---------------------------------------------------------------------------------------------
def port_arm(n_member, n_thread=0):
    """The reference's algorithm on host cores: oracle port (C, OpenMP), the headline workload.
    All host cores are used whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    from oracle import oracle
    n_thread = n_thread or host_threads()
    from popdynamics_b200 import compiler
    model = make_model()
    model.check_flow_transfers()
    om = oracle.OracleModel(compiler.compile_model(model, 'discrete_time_stochastic'))
    t0 = time.perf_counter()
    out = om.stochastic_batch('discrete_time_stochastic', n_member, model.get_init_list(),
                              model.target_times, seed=0, want_soln=False, want_moments=True,
                              n_thread=n_thread)
    dt = time.perf_counter() - t0
    assert out['err'] == 0
    return out['n_steps'] * len(model.labels) / dt, dt, n_thread


def sized_port_sample(target_seconds):
    rate, dt, cores = port_arm(2000)
    n_member = max(2000, int(target_seconds * rate / (T_END * 5)))
    return n_member, cores


def port_baseline(seconds):
    members, cores = sized_port_sample(seconds)
    value, dt, cores = port_arm(members)
    return {'value': value, 'unit': 'comp-steps/s', 'cores': cores, 'kind': 'port',
            'sample': '%d members x 1000 steps x 5 compartments (%.1f s), OpenMP over members; C '
                      'restatement of basepop.py:797-852 (oracle/popdyn_oracle.c)'
                      % (members, dt)}


def reference_python_baseline(seconds):
    """The UNMODIFIED reference's own Python loop (oracle/_ref, compiled here from
    /root/reference by `make -C oracle _ref`) on every host core; None when it did not ship."""
    try:
        from oracle import ref_runner
    except ImportError:
        return None
    if not ref_runner.available():
        return None
    return ref_runner.time_strains_dts(seconds, host_threads())


def run_reference_arm(args):
    """CPU arm: the reference's own implementation of the headline path on every host core --
    the unmodified Python loop from oracle/_ref when it shipped with this run, else the oracle
    port.  Each step is a bounded sample of the workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = host_threads()
    ref_ok = False
    try:
        from oracle import ref_runner
        ref_ok = ref_runner.available()
    except ImportError:
        pass
    steps = max(1, args.steps)
    if ref_ok:
        per_step = max(2.0, min(15.0, 60.0 / steps))
        for _ in range(args.warmup):
            ref_runner.time_strains_dts(1.0, cores)
        t0 = time.perf_counter()
        runs = [ref_runner.time_strains_dts(per_step, cores) for _ in range(steps)]
        elapsed = time.perf_counter() - t0
        value = float(np.mean([r['value'] for r in runs]))
        whole = float(np.mean([r['whole_integrate_value'] for r in runs]))
        members = sum(r['members'] for r in runs) // steps
        kind = 'reference'
        sample = ('%d members x 1000 tau-leap steps x 5 compartments per step over %d worker '
                  'processes, >= %.0f s each, timed inside the workers; unmodified '
                  'basepop.py:797-852 compiled into oracle/_ref, integrator only'
                  % (members, cores, per_step))
        extra = {'whole_integrate_value': whole, 'port': port_baseline(3.0)}
    else:
        members, cores = sized_port_sample(2.0)
        for _ in range(args.warmup):
            port_arm(max(500, members // 4))
        t0 = time.perf_counter()
        for _ in range(steps):
            port_arm(members)
        elapsed = time.perf_counter() - t0
        value = steps * members * T_END * 5 / elapsed
        kind = 'port'
        sample = ('%d members x %d tau-leap steps x 5 compartments per step, streaming moments, '
                  'OpenMP over members; C restatement of basepop.py:797-852 '
                  '(oracle/popdyn_oracle.c)' % (members, T_END))
        extra = {'note': 'oracle/_ref (the reference\'s own Python loop) did not ship with this '
                         'run; the pure-Python reference measured 1.09e5 comp-steps/s per core '
                         '(BASELINE.md)'}
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'comp-steps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * elapsed / steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int64+f64', 'data': 'synthetic',
        'config': {'workload': workload_name(args.members),
                   'sample_members_per_step': members,
                   'note': 'each step is a bounded sample of the workload run by the CPU '
                           'implementation on all host cores'},
        'cpu_baseline': dict({'value': value, 'unit': 'comp-steps/s', 'cores': cores,
                              'kind': kind, 'sample': sample}, **extra),
        'e2e': {'value': value, 'unit': 'comp-steps/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
class Gpu(object):
    """Process-wide handles of the GPU arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, self.world))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py needs a CUDA device: there is no CPU fallback')
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device('cuda', self.local_rank)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # > 126 MB L2
        self.launches = 0

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, value):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, value):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, flush=True):
        """Device time of fn() in ms (CUDA events on the current stream, L2 flushed before)."""
        torch = self.torch
        if flush:
            self.flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out


def strains_step_fn(g, args, model, cm, M, start, moments, hist, quant):
    """One pass of the headline hot path over this rank's members [start, start + M)."""
    from popdynamics_b200 import capi, ensemble
    C, T = len(model.labels), len(model.target_times)
    y = model.get_init_list()
    hist_shift = np.zeros(C, dtype=np.int32)

    def one_step(seed, count_bits=None):
        moments.zero_()
        hist.zero_()
        res = ensemble.run_compiled(model, cm, 'discrete_time_stochastic', y, model.target_times,
                                    M, output='stats', seed=seed, member_offset=start,
                                    device=g.local_rank,
                                    count_bits=count_bits or args.count_bits,
                                    block=args.block, min_blocks=args.min_blocks,
                                    hist_bins=HIST_BINS, hist_max=HIST_MAX,
                                    moments=moments, hist=hist, moments_from=args.moments_from)
        ensemble.raise_device_error(res.info, 'discrete_time_stochastic')
        if g.world > 1:
            g.dist.all_reduce(hist)
            if args.moments_from == 'hist':
                res.derive_moments()              # exact moments of the whole ensemble
            else:
                g.dist.all_reduce(moments)
        capi.hist_quantiles(hist, T, C, HIST_BINS, hist_shift, QUANTILES, quant,
                            device=g.local_rank)
        return res
    return one_step


def run_headline(g, args):
    torch = g.torch
    from popdynamics_b200 import capi, compiler, distributed, ensemble

    M = args.members
    start = g.rank * M                      # weak scaling: every rank owns M members
    model = make_model()
    model.check_flow_transfers()
    model.device = g.local_rank
    C, T = len(model.labels), len(model.target_times)
    cm = compiler.compile_model(model, 'discrete_time_stochastic')

    moments = torch.zeros((T, C, 2), dtype=torch.int64, device=g.dev)
    hist = torch.zeros((T, C, HIST_BINS), dtype=torch.int32, device=g.dev)
    quant = torch.empty((T, C, len(QUANTILES)), dtype=torch.int64, device=g.dev)
    one_step = strains_step_fn(g, args, model, cm, M, start, moments, hist, quant)

    for w in range(args.warmup):
        one_step(w)
    peaks, peak_kind = measured_peaks()
    int_peak = capi.microbench(2, g.local_rank)        # G int32 op/s (IMAD.WIDE / LOP3 mix)
    fp64_peak = capi.microbench(0, g.local_rank)       # G fp64 op/s (unfused DMUL + DADD)

    sampler = ClockSampler(g.local_rank)
    if g.rank == 0:
        sampler.start()
    g.barrier()
    step_ms, kernel_ms = [], []
    info = None
    t_wall = time.perf_counter()
    for k in range(args.steps):
        ms, res = g.timed(lambda: one_step(1000 + k))       # L2 evicted between timed iterations
        step_ms.append(ms)
        kernel_ms.append(res.info['kernel_ms'])
        info = res.info
    g.barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if g.rank == 0 else None
    # per timed step: pd_tau_table_kernel, pd_tau_leap_kernel, pd_hist_moments_kernel,
    # pd_quantiles_kernel
    g.launches += 4 * args.steps

    total_ms = g.max_over_ranks(sum(step_ms))
    steps_per_member = T - 1
    comp_steps_per_step = float(g.world) * M * steps_per_member * C
    value = comp_steps_per_step * args.steps / (total_ms * 1e-3)

    # ---- the same step with 64-bit counts (the reference holds Python ints / numpy.int64) -----
    one_step(7, count_bits=64)
    ms64 = [g.timed(lambda: one_step(3000 + k, count_bits=64))[0] for k in range(2)]
    ms64 = g.max_over_ranks(sum(ms64))
    value64 = comp_steps_per_step * 2 / (ms64 * 1e-3)

    # ---- strong scaling: 10^7 members in total over the ranks ---------------------------------
    strong = None
    total_members = args.strong_members
    s0, s1 = distributed.shard_range(total_members, g.rank, g.world)
    if s1 > s0:
        strong_step = strains_step_fn(g, args, model, cm, s1 - s0, s0, moments, hist, quant)
        strong_step(11)
        sms = [g.timed(lambda: strong_step(4000 + k))[0] for k in range(args.steps)]
        sms = g.max_over_ranks(sum(sms))
        strong = {'members_total': total_members, 'members_per_gpu': s1 - s0,
                  'value': float(total_members) * steps_per_member * C * args.steps / (sms * 1e-3),
                  'unit': 'comp-steps/s', 'ms_per_step': sms / args.steps, 'scaling': 'strong',
                  'limiter': 'fixed per-step cost: all-reduce of the %.0f MB int32 histograms + '
                             'pd_hist_moments_kernel + pd_quantiles_kernel'
                             % (hist.numel() * 4 / 1e6)}

    # ---- N > 1: the reduced statistics must equal a single-GPU run of the same global ids -----
    multi_check = None
    if g.world > 1:
        n_chk = 100000
        c0, c1 = distributed.shard_range(n_chk, g.rank, g.world)
        chk = strains_step_fn(g, args, model, cm, c1 - c0, c0, moments, hist, quant)
        chk(99)
        reduced_hist, reduced_mom = hist.clone(), moments.clone()
        if g.rank == 0:
            saved = g.world
            g.world = 1                                   # rank 0 alone, no collective
            try:
                alone = strains_step_fn(g, args, model, cm, n_chk, 0, moments, hist, quant)
                alone(99)
            finally:
                g.world = saved
            same = bool(torch.equal(hist, reduced_hist) and torch.equal(moments, reduced_mom))
            multi_check = {'members': n_chk, 'ranks': g.world, 'identical': same,
                           'what': 'all-reduced histograms + derived moments of a sharded run '
                                   'vs rank 0 integrating the same global member ids alone'}
            if not same:
                raise SystemExit('multi-GPU statistics differ from the single-GPU run')
        g.barrier()

    # ---- end to end through the public API: HOST inputs (model, initial state, target times)
    # in, HOST statistics (mean, variance, quantiles) out.  statistics='device' keeps the
    # accumulators in HBM, so the 41 MB of histograms never cross PCIe: only what the user asked
    # for does.
    def e2e_step(seed):
        res = model.integrate_ensemble('discrete_time_stochastic', n_members=M, output='stats',
                                       seed=seed, member_offset=start, device=g.local_rank,
                                       count_bits=args.count_bits, block=args.block,
                                       min_blocks=args.min_blocks, statistics='device',
                                       moments_from=args.moments_from,
                                       hist_bins=HIST_BINS, hist_max=HIST_MAX)
        if g.world > 1:
            distributed.allreduce_stats(res, device=g.dev)
            if args.moments_from == 'hist':
                res.derive_moments()
        q = res.quantiles(QUANTILES).cpu().numpy()
        mean, var = res.mean(g.world * M), res.var(g.world * M)
        out_bytes = q.nbytes + 2 * res.moments.numel() * 8      # quantiles + moments (x2 reads)
        return res, (q, mean, var), out_bytes

    e2e_step(0)
    g.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    h2d = d2h = 0
    for k in range(e2e_steps):
        res, _, out_bytes = e2e_step(2000 + k)
        h2d, d2h = res.info['h2d_bytes'], res.info['d2h_bytes'] + out_bytes
    g.barrier()
    e2e_s = g.max_over_ranks(time.perf_counter() - t0)
    e2e_value = comp_steps_per_step * e2e_steps / e2e_s

    if g.rank != 0:
        return None
    # ---- roofline of the dominant kernel (pd_tau_leap_kernel) ---------------------------------
    k_ms = float(np.mean(kernel_ms))
    member_steps = float(M) * steps_per_member
    sm_mhz = (clocks or {}).get('sm_mhz') or peaks.get('sm_max_mhz') or 1965.0
    issue_peak = SM_COUNT * LANES_PER_SM_PER_CLK * sm_mhz * 1e6 / 1e9      # G thread-inst/s
    unit = ncu_unit()
    hbm_peak = peaks.get('hbm_gbs')
    survey = SURVEY_OPS_PER_MEMBER_STEP * member_steps / (k_ms * 1e-3) / 1e9
    if unit:
        useful_per = float(unit['thread_inst_executed_per_member_step'])
        pred_on_per = float(unit['thread_inst_executed_true_per_member_step'])
        achieved = useful_per * member_steps / (k_ms * 1e-3) / 1e9
        frac = achieved / issue_peak
        frac_pred_on = pred_on_per * member_steps / (k_ms * 1e-3) / 1e9 / issue_peak
        traffic = float(unit['dram_bytes_per_member_step']) * member_steps
        # the busiest pipe of the capture (FMA-heavy: the IMAD.WIDE.U32 of the Philox rounds),
        # scaled from the capture's throughput to this run's: busy cycles per member-step are a
        # property of the code, so the busy fraction moves with member-steps/s
        capture_rate = float(unit['units_in_launch']) / (float(unit['duration_ms']) * 1e-3)
        frac_pipe = float(unit['pipe_fmaheavy_pct']) / 100.0 * \
            (member_steps / (k_ms * 1e-3)) / capture_rate
        per_unit = ('%.0f thread-instructions executed per member-step (active lanes of every '
                    'issued warp instruction; %.0f with the predicate on), measured by ncu on '
                    'this kernel (%s, %s); the SURVEY.md 8d estimate of %.0f operations is '
                    'reported as frac_survey_unit'
                    % (useful_per, pred_on_per, os.path.relpath(NCU_UNIT_FILE, ROOT),
                       unit.get('capture', ''), SURVEY_OPS_PER_MEMBER_STEP))
    else:
        achieved, frac, traffic, frac_pred_on, frac_pipe = None, None, None, None, None
        per_unit = 'no ncu capture committed for this build (%s missing)' \
            % os.path.relpath(NCU_UNIT_FILE, ROOT)
    roofline = {
        'bound': 'issue', 'kernel': 'pd_tau_leap_kernel',
        'achieved': achieved, 'peak': issue_peak, 'unit': 'G thread-inst/s',
        'frac': frac, 'frac_pred_on': frac_pred_on, 'frac_survey_unit': survey / issue_peak,
        'frac_bounding_pipe': frac_pipe,
        'ncu': ({k: unit.get(k) for k in ('issue_active_pct', 'lanes_per_inst',
                                          'frac_issue_x_lanes', 'pipe_fmaheavy_pct',
                                          'warp_inst_per_unit')} if unit else None),
        'traffic': traffic,
        'peak_source': '%d SMs x %d lanes/clk x %.0f MHz (median SM clock sampled during the '
                       'timed region)' % (SM_COUNT, LANES_PER_SM_PER_CLK, sm_mhz),
        'per_unit': per_unit,
        'kernel_ms': k_ms, 'share_of_step': k_ms / float(np.mean(step_ms)),
        'pipe_peaks_gops': {'int32_imad_wide_lop3': int_peak, 'fp64_dmul_dadd': fp64_peak},
        'hbm': {'algorithmic_bytes_per_member_step': 0.0,
                'full_trajectory_bytes_per_member_step': 8.0 * C,
                'peak_gbs': hbm_peak, 'peak_kind': peak_kind,
                'full_trajectory_ceiling_member_steps_per_s':
                    (hbm_peak * 1e9 / (8.0 * C)) if hbm_peak else None},
        'note': 'neither HBM- nor tensor-bound: streaming-statistics mode moves ~0 B per '
                'member-step, and even full-trajectory output (8*C B per member-step) has an '
                'HBM ceiling ~10x above what the Poisson draws allow; the bound is instruction '
                'issue, and inside it the integer datapath: IMAD.WIDE.U32 (two per Philox round) '
                'holds the FMA-heavy pipe for 4 cycles per scheduler and every further ALU '
                'instruction adds 2 more instead of hiding behind it (pd_microbench kinds 2-4, '
                '9-10: profiles/r2_coissue_probe.jsonl).  frac = executed thread-instructions '
                'over the issue peak (issue-active x lanes per instruction); frac_bounding_pipe '
                '= busy fraction of the busiest pipe (FMA-heavy) as ncu measured it',
    }
    cpu = port_baseline(args.cpu_seconds)
    try:
        ref = reference_python_baseline(min(6.0, args.cpu_seconds))
    except Exception as exc:                                   # pragma: no cover
        ref = {'unavailable': str(exc)}
    if ref:
        cpu['reference_python'] = ref
    return {
        'metric': METRIC, 'value': value, 'unit': 'comp-steps/s', 'n_gpus': g.world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'int%d+f64' % args.count_bits, 'data': 'synthetic',
        'members_per_sec': float(g.world) * M * args.steps / (total_ms * 1e-3),
        'value_count_bits_64': value64,
        'config': {'workload': workload_name(M),
                   'members_per_gpu': M, 'steps_per_member': steps_per_member,
                   'parallelism': 'members sharded x%d, all-reduce of integer statistics'
                                  % g.world,
                   'l2': 'flushed between timed iterations (256 MiB fill)',
                   'block': info['block'], 'regs_per_thread': info['regs_per_thread'],
                   'blocks_per_sm': info['blocks_per_sm']},
        'roofline': roofline,
        'cpu_baseline': cpu,
        'e2e': {'value': e2e_value, 'unit': 'comp-steps/s', 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'steps': e2e_steps},
        'strong_scaling': strong,
        'multi_gpu_check': multi_check,
        'clocks': clocks, 'wall_s': wall,
    }


# ---------------------------------------------------------------------------------------------
# The other BASELINE.json configurations
# ---------------------------------------------------------------------------------------------
def golden():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_vectors.npz'))


def oracle_explicit_rate(model_fn, n_sample, seconds_hint):
    """Oracle port (CPU baseline leg) on a bounded sample of an explicit sweep."""
    from oracle import oracle
    from popdynamics_b200 import compiler
    model = model_fn(0, n_sample)
    model.check_flow_transfers()
    cm = compiler.compile_model(model, 'explicit')
    om = oracle.OracleModel(cm)
    cores = host_threads()
    t0 = time.perf_counter()
    out = om.explicit_batch(n_sample, model.get_init_list(), model.target_times,
                            want_soln=False, n_thread=cores)
    dt = time.perf_counter() - t0
    steps = out['n_steps']
    return {'value': steps * cm.n_comp / dt, 'unit': 'comp-steps/s', 'cores': cores,
            'kind': 'port', 'steps_per_s': steps / dt,
            'sample': '%d members of the same sweep, full length (%.1f s)' % (n_sample, dt)}


def explicit_sweep_config(g, args, name, model_fn, total_members, ops_key, want_var, check):
    """An explicit-Euler parameter sweep over the N ranks: device-timed with resident inputs,
    then end to end through BaseModel.integrate_ensemble with HOST parameter columns."""
    torch = g.torch
    from popdynamics_b200 import capi, compiler, distributed, ensemble
    s0, s1 = distributed.shard_range(total_members, g.rank, g.world)
    M = s1 - s0
    model = model_fn(s0, s1)
    model.device = g.local_rank
    model.check_flow_transfers()
    cm = compiler.compile_model(model, 'explicit')
    C = cm.n_comp
    y = model.get_init_list()
    rows = ensemble._member_param_rows(cm, M)
    mp = torch.stack([torch.from_numpy(r) for r in rows]).to(g.dev)
    out = torch.empty((C, M), dtype=torch.float64, device=g.dev)

    def device_step():
        res = ensemble.run_compiled(model, cm, 'explicit', y, model.target_times, M,
                                    output='final', device=g.local_rank, member_params=mp, out=out)
        ensemble.raise_device_error(res.info, 'explicit')
        return res
    device_step()
    g.barrier()
    ms, res = g.timed(device_step)
    ms = g.max_over_ranks(ms)
    g.launches += 1
    steps_per_member = int(res.info['n_steps']) // M
    total_steps = float(total_members) * steps_per_member
    value = total_steps * C / (ms * 1e-3)
    kernel_ms = res.info['kernel_ms']
    del mp, out
    torch.cuda.empty_cache()

    # end to end: host parameter columns in (chunked, double-buffered inside the C-ABI), host
    # result out -- the final state, or one diagnostic var per member written by the integrator
    # kernel itself (final_vars: the diagnostics pass of the final state fused into it)
    def e2e_step():
        if want_var:
            r = model.integrate_ensemble('explicit', n_members=M, final_vars=[want_var],
                                         device=g.local_rank)
            return r, r.final_var(want_var), [r.info['h2d_bytes'], r.info['d2h_bytes']], \
                r.info['n_launch']
        r = model.integrate_ensemble('explicit', n_members=M, output='final',
                                     device=g.local_rank)
        return r, r.final, [r.info['h2d_bytes'], r.info['d2h_bytes']], r.info['n_launch']
    if want_var:                   # NVRTC of the fused final-vars variant, outside the timing
        tiny = model_fn(s0, s0 + min(M, 2048))
        tiny.device = g.local_rank
        tiny.integrate_ensemble('explicit', n_members=min(M, 2048), final_vars=[want_var],
                                device=g.local_rank)
    g.barrier()
    t0 = time.perf_counter()
    r, host_out, moved, n_chunk = e2e_step()
    g.barrier()
    e2e_s = g.max_over_ranks(time.perf_counter() - t0)
    g.launches += n_chunk
    parity = check(s0, s1, r, host_out) if g.rank == 0 else None
    del r
    torch.cuda.empty_cache()
    if g.rank != 0:
        return None
    fp64_peak = capi.microbench(0, g.local_rank)
    ops = EXPLICIT_OPS[ops_key]
    achieved = ops * (float(M) * steps_per_member) / (kernel_ms * 1e-3) / 1e9
    return {
        'config': name, 'n_gpus': g.world, 'scaling': 'strong', 'members_total': total_members,
        'members_per_gpu': M, 'steps_per_member': steps_per_member, 'compartments': C,
        'ms': ms, 'value': value, 'unit': 'comp-steps/s', 'steps_per_s': total_steps / (ms * 1e-3),
        'members_per_sec': total_members / (ms * 1e-3),
        'dtype': 'f64', 'data': 'synthetic',
        'e2e': {'value': total_steps * C / e2e_s, 'unit': 'comp-steps/s',
                'h2d_bytes_per_step': int(moved[0]), 'd2h_bytes_per_step': int(moved[1]),
                'seconds': e2e_s, 'chunks': n_chunk, 'over_device_value': (total_steps * C / e2e_s) / value},
        'roofline': {'bound': 'fp64 pipe', 'kernel': 'pd_explicit_kernel',
                     'achieved': achieved, 'peak': fp64_peak, 'unit': 'G fp64 op/s',
                     'frac': achieved / fp64_peak, 'traffic': None,
                     'per_unit': '%.0f fp64 operations per derivative evaluation (SURVEY.md 8d), '
                                 'unfused (--fmad=false); peak = pd_microbench DMUL+DADD on this '
                                 'GPU' % ops},
        'parity': parity,
    }


def config_seir(g, args):
    from examples import configs
    gold = golden()[configs.SEIR_GOLDEN_KEY]

    def check(s0, s1, r, final):
        got = np.asarray(final)[:, 0]
        ok = bool(np.array_equal(got, gold))
        if not ok:
            raise SystemExit('config 2: member 0 differs from the reference golden: %r vs %r'
                             % (got, gold))
        return {'what': 'member 0 (measles base case) equals the unmodified reference\'s '
                        '18 250-day final state bit for bit (tests/golden, key %s)'
                        % configs.SEIR_GOLDEN_KEY, 'ok': ok}
    line = explicit_sweep_config(
        g, args, '2: SEIR-demography sweep (r0, duration_infectious, duration_preinfectious), '
        'explicit Euler, make_times(0, 18250, 1)', configs.seir_sweep, args.seir_members,
        'seir', None, check)
    if line is not None:
        line['cpu_baseline'] = oracle_explicit_rate(configs.seir_sweep, 16 * host_threads(), 3)
    return line


def config_rmit(g, args):
    from examples import configs
    gold = golden()['rmit_fig2_proportion']

    def check(s0, s1, r, proportion):
        n = min(len(gold), s1 - s0) if s0 == 0 else 0
        ok = bool(np.array_equal(np.asarray(proportion)[:n], gold[:n]))
        if not ok:
            raise SystemExit('config 5: embedded golden members differ: %r vs %r'
                             % (np.asarray(proportion)[:n], gold[:n]))
        finite = bool(np.isfinite(np.asarray(proportion)).all())
        return {'what': 'members 0..5 (six beta points of the reference\'s figure 2) equal '
                        'model.vars[\'proportion\'] of the unmodified reference bit for bit '
                        '(tests/golden, key rmit_fig2_proportion); every proportion finite',
                'ok': ok and finite}
    line = explicit_sweep_config(
        g, args, '5: RMIT TB calibration sweep (beta, rho, sigma1..3), explicit Euler, '
        'make_times(0, 500, 1), output = vars[\'proportion\'] of the final state per member',
        configs.rmit_sweep, args.rmit_members, 'rmit', 'proportion', check)
    if line is not None:
        line['cpu_baseline'] = oracle_explicit_rate(configs.rmit_sweep, 1024 * host_threads(), 3)
    return line


def config_tb(g, args):
    """Config 4: the TB model under Gillespie for 50 simulated years; members scale with N
    (125 000 per GPU: the stated 10^6 on 8 GPUs -- about 5.5e6 events per member)."""
    from examples import configs
    from oracle import oracle
    from popdynamics_b200 import capi, compiler, ensemble
    torch = g.torch
    M = args.tb_members
    start = g.rank * M
    method = 'continuous_time_stochastic'
    warm = configs.tb_gillespie()
    warm.device = g.local_rank
    warm.make_times(1950, 1950.01, 0.01)                  # NVRTC compile + module load
    warm.integrate_ensemble(method, n_members=256, output='final', seed=0, count_bits=32,
                            device=g.local_rank)
    model = configs.tb_gillespie()
    model.device = g.local_rank
    g.barrier()
    t0 = time.perf_counter()
    res = model.integrate_ensemble(method, n_members=M, output='final', seed=0,
                                   member_offset=start, count_bits=32, device=g.local_rank)
    g.barrier()
    wall = g.max_over_ranks(time.perf_counter() - t0)
    ms = g.max_over_ranks(res.info['kernel_ms'])
    events = g.sum_over_ranks(res.info['n_steps'])
    g.launches += 1
    final = np.asarray(res.final)
    mean_pop = g.sum_over_ranks(final.sum()) / (g.world * M)
    if g.rank != 0:
        return None
    C = final.shape[0]
    # the mean-field limit of the same model: explicit Euler over the same 50 years (population
    # 1e6, so the ensemble mean of the total population sits within a fraction of a percent)
    ode = configs.tb_gillespie()
    ode.device = g.local_rank
    ode.integrate('explicit')
    ode_pop = float(ode.soln_array[-1].sum())
    cm = compiler.compile_model(model, method)
    # CPU baseline leg: the oracle port, one member per host core, the full 50 years
    cores = host_threads()
    om = oracle.OracleModel(cm)
    t0 = time.perf_counter()
    out = om.stochastic_batch(method, cores, model.get_init_list(), model.target_times, seed=0,
                              want_soln=False, n_thread=cores)
    dt = time.perf_counter() - t0
    S = cm.n_flow
    ops_per_event = 11 + 1 + 2 * S + 7 * S + 50       # vars, rates, cumsum, Neumaier sum, log/div
    fp64_peak = capi.microbench(0, g.local_rank)
    achieved = ops_per_event * (res.info['n_steps'] / (res.info['kernel_ms'] * 1e-3)) / 1e9
    return {
        'config': '4: TB model (scale-ups evaluated in the loop), Gillespie direct method, '
                  'make_times(1950, 2000, 1), population 1e6',
        'n_gpus': g.world, 'scaling': 'weak', 'members_total': g.world * M,
        'members_per_gpu': M, 'events': events, 'events_per_member': events / (g.world * M),
        'ms': ms, 'value': events / (ms * 1e-3), 'unit': 'events/s',
        'comp_steps_per_s': events * C / (ms * 1e-3),
        'members_per_sec': g.world * M / (ms * 1e-3), 'dtype': 'int32+f64', 'data': 'synthetic',
        'note': '125 000 members per GPU by default: the stated 10^6 members on 8 GPUs',
        'e2e': {'value': events / wall, 'unit': 'events/s',
                'h2d_bytes_per_step': int(res.info['h2d_bytes']),
                'd2h_bytes_per_step': int(res.info['d2h_bytes']), 'seconds': wall},
        'roofline': {'bound': 'issue / fp64 pipe', 'kernel': 'pd_gillespie_kernel',
                     'achieved': achieved, 'peak': fp64_peak, 'unit': 'G fp64 op/s',
                     'frac': achieved / fp64_peak, 'traffic': None,
                     'regs_per_thread': res.info['regs_per_thread'],
                     'blocks_per_sm': res.info['blocks_per_sm'],
                     'per_unit': '%d fp64 operations per event (11 vars + %d rates + %d cumulative '
                                 'adds + 7 x %d compensated-sum operations + ~50 for log and '
                                 'divide)' % (ops_per_event, S, S, S)},
        'cpu_baseline': {'value': out['n_steps'] / dt, 'unit': 'events/s', 'cores': cores,
                         'kind': 'port',
                         'sample': '%d members x 50 years (%.1f s)' % (cores, dt)},
        'parity': {'what': 'no member raised the device error flag; ensemble mean of the final '
                           'total population %.7g vs %.7g from explicit Euler on the same model '
                           '(tolerance 1 %%); bit-exact comparison with the oracle is in '
                           'tests/test_gpu_configs.py' % (mean_pop, ode_pop),
                   'ok': bool(res.info['err_code'] == 0 and
                              abs(mean_pop - ode_pop) < 0.01 * ode_pop)},
    }


def hybrid_reference_baseline():
    """The reference's own mix_model loop (oracle/_ref) on the host cores; None when it did not
    ship with this run."""
    try:
        from oracle import ref_runner
        if ref_runner.available():
            return ref_runner.time_mix_hybrid(3.0, host_threads())
    except Exception as exc:                                   # pragma: no cover
        return {'unavailable': str(exc)}
    return None


def config_hybrid(g, args):
    """Config 5b: mix_model.py's day-by-day switching between Gillespie and explicit Euler."""
    from examples import configs
    from popdynamics_b200 import ensemble
    torch = g.torch
    M = args.hybrid_members
    model = configs.mix_model()
    model.device = g.local_rank
    ensemble.integrate_hybrid(model, configs.MIX_SEGMENTS, 4096, configs.MIX_CUTOFF,
                              seed=1, device=g.local_rank)       # warm-up: NVRTC of both kernels
    wall = None
    for rep in range(3):          # ~100 tiny launches: host-bound, so the best of three passes
        g.barrier()
        t0 = time.perf_counter()
        res = ensemble.integrate_hybrid(model, configs.MIX_SEGMENTS, M, configs.MIX_CUTOFF,
                                        seed=5, member_offset=g.rank * M, device=g.local_rank)
        torch.cuda.synchronize()
        g.barrier()
        took = g.max_over_ranks(time.perf_counter() - t0)
        wall = took if wall is None else min(wall, took)
    launches = res.info.get('launches', [])
    g.launches += len(launches)
    kernel_ms = sum(l['kernel_ms'] for l in launches)
    final = res.final
    # closed population of 100; entering a stochastic segment truncates with int()
    # (basepop.py:747-748), so the total can only shrink, by less than one per compartment and
    # switch
    total = final.sum(dim=0)
    conserved = bool(total.max().item() <= 100.0 + 1e-9 and total.min().item() > 50.0 and
                     final.min().item() >= 0.0)
    explicit_share = float(res.modes.float().mean().item())
    if g.rank != 0:
        return None
    return {
        'config': '5b: mix_model.py SIR (population 100), 50 one-day segments of step 0.2, '
                  'Gillespie while min(compartments) <= 3 else explicit Euler',
        'n_gpus': g.world, 'scaling': 'weak', 'members_total': g.world * M, 'members_per_gpu': M,
        'ms': wall * 1e3, 'kernel_ms_sum': kernel_ms, 'launches': len(launches),
        'value': g.world * M / wall, 'unit': 'members/s', 'dtype': 'int64+f64',
        'data': 'synthetic',
        'e2e': {'value': g.world * M / wall, 'unit': 'members/s', 'h2d_bytes_per_step': 24,
                'd2h_bytes_per_step': 0, 'seconds': wall},
        'roofline': None,
        'cpu_baseline': hybrid_reference_baseline(),
        'parity': {'what': 'closed population never grows and no compartment is negative; share of '
                           '(member, segment) pairs integrated explicitly %.3f; per-member '
                           'bit-exact comparison with the oracle loop is in '
                           'tests/test_gpu_configs.py' % explicit_share,
                   'ok': conserved},
    }


def config_adaptive(g, args):
    """Config 6: the adaptive solver behind integrate('scipy') on the SEIR-demography sweep.
    The reference itself cannot run this model through odeint beyond a few hundred days (LSODA
    evaluates the derivative on a slightly negative state and the reference's checks() assert
    fires), so the parity spot-check is a separate 300-day run against the odeint golden."""
    from examples import configs
    from popdynamics_b200 import capi, distributed
    torch = g.torch
    total = args.adaptive_members
    s0, s1 = distributed.shard_range(total, g.rank, g.world)
    M = s1 - s0
    short = configs.seir_sweep(s0, min(s1, s0 + 4096))
    short.device = g.local_rank
    short.make_times(0, 300, 1)
    chk = short.integrate_ensemble('scipy', n_members=min(M, 4096), output='full',
                                   device=g.local_rank)
    short.integrate_ensemble('scipy', n_members=min(M, 4096), output='final',
                             device=g.local_rank)          # NVRTC of the FINAL-output variant
    model = configs.seir_sweep(s0, s1)
    model.device = g.local_rank
    g.barrier()
    t0 = time.perf_counter()
    res = model.integrate_ensemble('scipy', n_members=M, output='final', device=g.local_rank,
                                   check=False)
    g.barrier()
    wall = g.max_over_ranks(time.perf_counter() - t0)
    ms = g.max_over_ranks(res.info['kernel_ms'])
    steps = g.sum_over_ranks(res.info['n_steps'])
    g.launches += res.info['n_launch'] + 1
    if g.rank != 0:
        return None
    want = golden()['seird_odeint_300']
    got = np.asarray(chk.soln)[:, :, 0]
    rel = float(np.abs(got - want).max() / np.abs(want).max())
    fp64_peak = capi.microbench(0, g.local_rank)
    # 7 derivative evaluations of 39 fp64 operations + ~28 C stage / error operations per step
    ops = 7 * EXPLICIT_OPS['seir'] + 28 * 4 + 60
    achieved = ops * (res.info['n_steps'] / (res.info['kernel_ms'] * 1e-3)) / 1e9
    return {
        'config': '6: SEIR-demography sweep through the adaptive Dormand-Prince 5(4) integrator '
                  '(integrate(\'scipy\') equivalent, rtol = atol = 1.49e-8), '
                  'make_times(0, 18250, 1)',
        'n_gpus': g.world, 'scaling': 'strong', 'members_total': total, 'members_per_gpu': M,
        'ms': ms, 'value': steps / (ms * 1e-3), 'unit': 'attempted steps/s',
        'steps_per_member': steps / total, 'members_per_sec': total / (ms * 1e-3),
        'dtype': 'f64', 'data': 'synthetic',
        'e2e': {'value': steps / wall, 'unit': 'attempted steps/s',
                'h2d_bytes_per_step': int(res.info['h2d_bytes']),
                'd2h_bytes_per_step': int(res.info['d2h_bytes']), 'seconds': wall},
        'roofline': {'bound': 'fp64 pipe', 'kernel': 'pd_adaptive_kernel', 'achieved': achieved,
                     'peak': fp64_peak, 'unit': 'G fp64 op/s', 'frac': achieved / fp64_peak,
                     'traffic': None,
                     'per_unit': '%.0f fp64 operations per attempted step (7 derivative '
                                 'evaluations + stage combinations + error norm)' % ops},
        'cpu_baseline': {'unavailable': 'the reference\'s integrate(\'scipy\') aborts with '
                                        'AssertionError on this model beyond ~300 days (its '
                                        'checks() fires inside LSODA, oracle/make_golden.py)'},
        'parity': {'what': 'member 0 over 300 days vs the unmodified reference\'s '
                           'integrate(\'scipy\') (odeint / LSODA) trajectory: max difference '
                           '%.2e of the largest compartment (tolerance 2e-6, DESIGN.md); members '
                           'flagged by checks() in the 50-year run: %s'
                           % (rel, 'first is %d' % res.info['err_member']
                              if res.info['err_code'] else 'none'),
                   'ok': rel < 2e-6},
    }


def run_gpu_arm(args):
    g = Gpu(args)
    line = run_headline(g, args)
    configs = []
    if args.configs != 'headline':
        wanted = ['seir', 'tb', 'rmit', 'hybrid', 'adaptive'] if args.configs == 'all' \
            else args.configs.split(',')
        table = {'seir': config_seir, 'tb': config_tb, 'rmit': config_rmit,
                 'hybrid': config_hybrid, 'adaptive': config_adaptive}
        for name in wanted:
            t0 = time.perf_counter()
            entry = table[name](g, args)
            g.barrier()
            if g.rank == 0 and entry is not None:
                entry['bench_wall_s'] = time.perf_counter() - t0
                configs.append(entry)
    if g.rank == 0:
        line['configs'] = configs
        line['gpu_launches'] = g.launches
        print(json.dumps(line), flush=True)
    if g.world > 1:
        g.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--members', type=int, default=10**7, help='members per GPU (weak scaling)')
    ap.add_argument('--strong-members', type=int, default=10**7,
                    help='members in total for the strong-scaling figure')
    ap.add_argument('--count-bits', type=int, default=32, choices=[32, 64],
                    help='width of the integer counts (32: overflow raises the device error flag)')
    ap.add_argument('--min-blocks', type=int, default=4, help='__launch_bounds__ resident-block cap')
    ap.add_argument('--block', type=int, default=256)
    ap.add_argument('--moments-from', default='hist', choices=['hist', 'kernel'],
                    help="'hist': exact moments derived from the width-1 histograms after the "
                         "launch; 'kernel': reduced inside the integrator kernel")
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--configs', default='all',
                    help="'all', 'headline', or a comma list of seir,tb,rmit,hybrid,adaptive")
    ap.add_argument('--seir-members', type=int, default=10**6, help='config 2, in total')
    ap.add_argument('--rmit-members', type=int, default=10**8, help='config 5, in total')
    ap.add_argument('--tb-members', type=int, default=125000, help='config 4, per GPU')
    ap.add_argument('--hybrid-members', type=int, default=1 << 20, help='config 5b, per GPU')
    ap.add_argument('--adaptive-members', type=int, default=10**6, help='config 6, in total')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
