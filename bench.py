#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
Headline benchmark: trajectory-compartment-steps/sec of the multi-strain SIR ensemble
(BASELINE.json config 3: examples/stochastic_strains_model.py, discrete-time stochastic,
10^7 members per GPU, make_times(0, 1000, 1), ensemble mean / variance / quantiles).

    python bench.py --gpus N --steps K --warmup W            # this framework, N GPUs (torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port on host cores

One "step" = one pass of the hot path over the whole ensemble: M members x 1000 tau-leap steps,
statistics fused, plus (N > 1) the NCCL all-reduce of the integer moments / histograms and the
quantile kernel.  ``value`` is measured with the statistics buffers resident in HBM, ``e2e``
through the public Python API (``BaseModel.integrate_ensemble``) with host numpy buffers.
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
# Roofline unit costs of pd_tau_leap_kernel on this workload.  The kernel
# is bound by SM instruction issue, so the work unit is the thread-operation.
#   ALGORITHMIC_OPS_PER_MEMBER_STEP: SURVEY.md 8(d)'s per-unit figure for the strains model --
#   S Poisson draws x (one exp ~ 25-30 fp64 ops + (lambda + 1) uniforms x ~45 int32 Philox ops)
#   ~ 300 fp64 + ~1000 int32 operations per member-step.  It does not depend on this
#   implementation (which needs fewer: see USEFUL_INST).
#   USEFUL_INST_PER_MEMBER_STEP: predicated-on thread instructions the current kernel executes
#   per member-step (ncu thread_inst_executed_true / member-steps of the captured launch,
#   profiles/r1_tau_leap_stats_bench_ncu_summary.txt): 1.914e12 / 2e9.
#   DRAM_BYTES_PER_MEMBER_STEP: dram__bytes_read + write of the same capture per member-step
#   (8.28 MB over 2e9 member-steps); the streaming-statistics mode reads y0 once and writes
#   nothing per step.
ALGORITHMIC_OPS_PER_MEMBER_STEP = 1300.0
USEFUL_INST_PER_MEMBER_STEP = 957.0
DRAM_BYTES_PER_MEMBER_STEP = 0.0041
SM_COUNT = 148
LANES_PER_SM_PER_CLK = 128          # 4 schedulers x 1 warp instruction x 32 lanes


def workload_name(members_per_gpu):
    return ('stochastic_strains_model DTS (clamped Poisson tau-leap), make_times(0,1000,1), C=5, '
            '%d members per GPU, streaming %d-bin width-1 histograms -> exact mean/var + 5 '
            'quantiles per (t,c)' % (members_per_gpu, HIST_BINS))


def make_model():
    from examples import models
    model = models.StrainsModel()
    model.make_times(0, T_END, 1)
    return model


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


# ---------------------------------------------------------------------------------------------
def cpu_arm(n_member, n_thread=0):
    """The reference's algorithm on host cores: oracle port (C, OpenMP), same workload.  All host
    cores are used whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    from oracle import oracle
    if not n_thread:
        n_thread = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') \
            else (os.cpu_count() or 1)
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


def sized_cpu_sample(target_seconds):
    rate, dt, cores = cpu_arm(2000)
    n_member = max(2000, int(target_seconds * rate / (T_END * 5)))
    return n_member, cores


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_member, cores = sized_cpu_sample(2.0)
    for _ in range(args.warmup):
        cpu_arm(max(500, n_member // 4))
    t0 = time.perf_counter()
    total = 0.0
    for _ in range(args.steps):
        rate, dt, cores = cpu_arm(n_member)
        total += n_member * T_END * 5
    elapsed = time.perf_counter() - t0
    value = total / elapsed
    sample = ('%d members x %d tau-leap steps x 5 compartments per step, streaming moments, '
              'OpenMP over members' % (n_member, T_END))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'comp-steps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * elapsed / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'int64+f64', 'data': 'synthetic',
        'config': {'workload': workload_name(args.members),
                   'sample_members_per_step': n_member,
                   'note': 'each step is a bounded sample of the workload: the oracle port runs '
                           'sample_members_per_step members with streaming moments'},
        'cpu_baseline': {'value': value, 'unit': 'comp-steps/s', 'cores': cores, 'kind': 'port',
                         'sample': sample,
                         'note': 'C restatement of basepop.py:797-852 (oracle/popdyn_oracle.c); '
                                 'the pure-Python reference itself measured 1.09e5 comp-steps/s '
                                 'per core (BASELINE.md)'},
        'e2e': {'value': value, 'unit': 'comp-steps/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from popdynamics_b200 import capi, compiler, distributed, ensemble

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: there is no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    M = args.members
    start = rank * M                      # weak scaling: every rank owns M members
    model = make_model()
    model.check_flow_transfers()
    model.device = local_rank
    C, T = len(model.labels), len(model.target_times)
    cm = compiler.compile_model(model, 'discrete_time_stochastic')
    y = model.get_init_list()

    n_rows = T
    moments = torch.zeros((T, C, 2), dtype=torch.int64, device=dev)
    hist = torch.zeros((n_rows, C, HIST_BINS), dtype=torch.int32, device=dev)
    quant = torch.empty((n_rows, C, len(QUANTILES)), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    hist_shift = np.zeros(C, dtype=np.int32)

    def one_step(seed):
        moments.zero_()
        hist.zero_()
        res = ensemble.run_compiled(model, cm, 'discrete_time_stochastic', y, model.target_times,
                                    M, output='stats', seed=seed, member_offset=start,
                                    device=local_rank, count_bits=args.count_bits,
                                    block=args.block, min_blocks=args.min_blocks,
                                    hist_bins=HIST_BINS, hist_max=HIST_MAX,
                                    moments=moments, hist=hist, moments_from=args.moments_from)
        ensemble.raise_device_error(res.info, 'discrete_time_stochastic')
        if world > 1:
            dist.all_reduce(hist)
            if args.moments_from == 'hist':
                res.derive_moments()              # exact moments of the whole ensemble
            else:
                dist.all_reduce(moments)
        capi.hist_quantiles(hist, n_rows, C, HIST_BINS, hist_shift, QUANTILES, quant,
                            device=local_rank)
        return res

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        one_step(w)
    peaks, peak_kind = measured_peaks()
    int_peak = capi.microbench(2, local_rank)        # G int32 op/s (IMAD / LOP3 mix)
    fp64_peak = capi.microbench(0, local_rank)       # G fp64 op/s (unfused DMUL + DADD)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    step_ms, kernel_ms, draws = [], [], []
    info = None
    t_wall = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xff)                         # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        res = one_step(1000 + k)
        e1.record()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        kernel_ms.append(res.info['kernel_ms'])
        info = res.info
    barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None

    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    steps_per_member = T - 1
    comp_steps_per_step = float(world) * M * steps_per_member * C
    value = comp_steps_per_step * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API: HOST inputs (model, initial state, target times)
    # in, HOST statistics (mean, variance, quantiles) out.  statistics='device' keeps the
    # accumulators in HBM, so the 41 MB of histograms never cross PCIe: only what the user asked
    # for does.
    def e2e_step(seed):
        res = model.integrate_ensemble('discrete_time_stochastic', n_members=M, output='stats',
                                       seed=seed, member_offset=start, device=local_rank,
                                       count_bits=args.count_bits, block=args.block,
                                       min_blocks=args.min_blocks, statistics='device',
                                       moments_from=args.moments_from,
                                       hist_bins=HIST_BINS, hist_max=HIST_MAX)
        if world > 1:
            distributed.allreduce_stats(res, device=dev)
        q = res.quantiles(QUANTILES).cpu().numpy()
        mean, var = res.mean(world * M), res.var(world * M)
        out_bytes = q.nbytes + 2 * res.moments.numel() * 8      # quantiles + moments (x2 reads)
        return res, (q, mean, var), out_bytes

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    h2d = d2h = 0
    for k in range(e2e_steps):
        res, _, out_bytes = e2e_step(2000 + k)
        h2d, d2h = res.info['h2d_bytes'], res.info['d2h_bytes'] + out_bytes
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = comp_steps_per_step * e2e_steps / float(e2e_s.item())

    if rank == 0:
        # ---- roofline of the dominant kernel (pd_tau_leap_kernel) ---------------------------
        k_ms = float(np.mean(kernel_ms))
        member_steps = float(M) * steps_per_member
        sm_mhz = (clocks or {}).get('sm_mhz') or peaks.get('sm_max_mhz') or 1965.0
        issue_peak = SM_COUNT * LANES_PER_SM_PER_CLK * sm_mhz * 1e6 / 1e9      # G thread-inst/s
        achieved = ALGORITHMIC_OPS_PER_MEMBER_STEP * member_steps / (k_ms * 1e-3) / 1e9
        useful = USEFUL_INST_PER_MEMBER_STEP * member_steps / (k_ms * 1e-3) / 1e9
        hbm_peak = peaks.get('hbm_gbs')
        roofline = {
            'bound': 'issue', 'kernel': 'pd_tau_leap_kernel',
            'achieved': achieved, 'peak': issue_peak, 'unit': 'G thread-op/s',
            'frac': achieved / issue_peak,
            'frac_executed_useful': useful / issue_peak,
            'traffic': (DRAM_BYTES_PER_MEMBER_STEP * member_steps
                        if DRAM_BYTES_PER_MEMBER_STEP is not None else None),
            'peak_source': '%d SMs x %d lanes/clk x %.0f MHz (median SM clock sampled during the '
                           'timed region)' % (SM_COUNT, LANES_PER_SM_PER_CLK, sm_mhz),
            'per_unit': '%.0f algorithmic operations per member-step (SURVEY.md 8d); the kernel '
                        'executes %.0f useful thread-instructions per member-step (ncu, profiles/)'
                        % (ALGORITHMIC_OPS_PER_MEMBER_STEP, USEFUL_INST_PER_MEMBER_STEP),
            'kernel_ms': k_ms, 'share_of_step': k_ms / float(np.mean(step_ms)),
            'pipe_peaks_gops': {'int32_imad_lop3': int_peak, 'fp64_dmul_dadd': fp64_peak},
            'hbm': {'algorithmic_bytes_per_member_step': 0.0,
                    'full_trajectory_bytes_per_member_step': 8.0 * C,
                    'peak_gbs': hbm_peak, 'peak_kind': peak_kind,
                    'full_trajectory_ceiling_member_steps_per_s':
                        (hbm_peak * 1e9 / (8.0 * C)) if hbm_peak else None},
            'note': 'neither HBM- nor tensor-bound: streaming-statistics mode moves ~0 B per '
                    'member-step, and even full-trajectory output (8*C B per member-step) has an '
                    'HBM ceiling ~10x above what the Poisson draws allow; the bound is '
                    'instruction issue (Philox rounds, fp64 compares, shared-memory bookkeeping)',
        }
        cpu_members, cores = sized_cpu_sample(args.cpu_seconds)
        cpu_value, cpu_dt, cores = cpu_arm(cpu_members)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'comp-steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'int%d+f64' % args.count_bits, 'data': 'synthetic',
            'members_per_sec': float(world) * M * args.steps / (total_ms * 1e-3),
            'config': {'workload': workload_name(M),
                       'members_per_gpu': M, 'steps_per_member': steps_per_member,
                       'parallelism': 'members sharded x%d, all-reduce of integer statistics'
                                      % world,
                       'l2': 'flushed between timed iterations (256 MiB fill)',
                       'block': info['block'], 'regs_per_thread': info['regs_per_thread'],
                       'blocks_per_sm': info['blocks_per_sm']},
            'roofline': roofline,
            'cpu_baseline': {'value': cpu_value, 'unit': 'comp-steps/s', 'cores': cores,
                             'kind': 'port',
                             'sample': '%d members x 1000 steps x 5 compartments (%.1f s)'
                                       % (cpu_members, cpu_dt)},
            'e2e': {'value': e2e_value, 'unit': 'comp-steps/s', 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': int(d2h), 'steps': e2e_steps},
            # per timed step: pd_tau_table_kernel, pd_tau_leap_kernel, pd_hist_moments_kernel,
            # pd_quantiles_kernel
            'gpu_launches': 4 * args.steps,
            'clocks': clocks, 'wall_s': wall,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--members', type=int, default=10**7, help='members per GPU')
    ap.add_argument('--count-bits', type=int, default=32, choices=[32, 64],
                    help='width of the integer counts (32: overflow raises the device error flag)')
    ap.add_argument('--min-blocks', type=int, default=4, help='__launch_bounds__ resident-block cap')
    ap.add_argument('--block', type=int, default=256)
    ap.add_argument('--moments-from', default='hist', choices=['hist', 'kernel'],
                    help="'hist': exact moments derived from the width-1 histograms after the "
                         "launch; 'kernel': reduced inside the integrator kernel")
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
