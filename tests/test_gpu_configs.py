"""GPU tests of the BASELINE.json configurations as bench.py runs them (examples/configs.py): the
same builders at sizes the oracle finishes in seconds, members embedded in the sweeps compared
with the reference's golden vectors and with the oracle, the chunked host-staging pipeline of the
C-ABI against a single resident launch, and the multi-rank path over NCCL when the box has two
GPUs.  Bit-exact everywhere (explicit fp64, integer counts)."""
import os
import socket

import numpy as np
import pytest
import torch

from examples import configs
from oracle import oracle
from popdynamics_b200 import compiler, distributed, ensemble

import helpers

pytestmark = pytest.mark.gpu


def _oracle_explicit_member(model, i):
    """Oracle run of member i of a sweep model (explicit)."""
    model.check_flow_transfers()
    cm = compiler.compile_model(model, 'explicit')
    om = oracle.OracleModel(cm)
    soln, _, err = om.explicit(model.get_init_list(), model.target_times,
                               params=om.member_params(i))
    assert err == 0
    return soln


# ------------------------------------------------------------------------------- config 2
def test_config2_seir_sweep_50_years_golden_member_and_oracle_members(golden):
    n = 4096
    model = configs.seir_sweep(0, n)
    res = model.integrate_ensemble('explicit', n_members=n, output='final')
    # member 0 = the measles base case: 18 250 days, 376 696 derivative evaluations
    assert np.array_equal(res.final[:, 0], golden[configs.SEIR_GOLDEN_KEY])
    assert res.info['n_steps'] == 376696 * n
    for i in (1, 77, n - 1):
        assert np.array_equal(res.final[:, i], _oracle_explicit_member(model, i)[-1]), i


def test_config2_trajectory_rows_of_the_golden_member(golden):
    model = configs.seir_sweep(0, 64)
    res = model.integrate_ensemble('explicit', n_members=64, output='full')
    assert np.array_equal(res.soln[::365, :, 0], golden['seird_explicit_18250_every365'])


def test_sweeps_do_not_depend_on_the_shard_they_are_built_for():
    whole = configs.rmit_sweep(0, 3 * configs.BLOCK + 17)
    part = configs.rmit_sweep(configs.BLOCK - 5, 2 * configs.BLOCK + 9)
    for key in ('beta', 'rho', 'sigma1', 'sigma2', 'sigma3'):
        assert np.array_equal(whole.params[key][configs.BLOCK - 5:2 * configs.BLOCK + 9],
                              part.params[key])


# ------------------------------------------------------------------------------- config 5
def test_config5_rmit_sweep_golden_points_and_chunked_staging(golden, monkeypatch):
    n = 20000
    model = configs.rmit_sweep(0, n)
    # one resident launch ...
    whole = model.integrate_ensemble('explicit', n_members=n, output='final')
    assert whole.info['n_launch'] == 1
    # ... against the chunk pipeline (host parameter columns staged 3000 members at a time,
    # ragged last chunk), final state kept on the device
    monkeypatch.setenv('POPDYN_CHUNK_MEMBERS', '3000')
    out = torch.empty((6, n), dtype=torch.float64, device='cuda')
    chunked = model.integrate_ensemble('explicit', n_members=n, output='final', out=out)
    assert chunked.info['n_launch'] == 7
    assert np.array_equal(out.cpu().numpy(), whole.final)
    # host output through the pipeline as well
    host = model.integrate_ensemble('explicit', n_members=n, output='final')
    assert host.info['n_launch'] == 7 and np.array_equal(host.final, whole.final)
    monkeypatch.delenv('POPDYN_CHUNK_MEMBERS')
    # vars['proportion'] of the final state, one scalar per member (rmit_tb_model.py:143)
    prop = np.empty((1, n))
    d = ensemble.diagnostics(model, chunked, want_vars=['proportion'], out=prop)
    assert d.var_labels == ['proportion'] and d.vars.shape == (1, n)
    assert np.array_equal(prop[0, :6], golden['rmit_fig2_proportion'])
    assert np.array_equal(whole.final[:, :6].T, golden['rmit_fig2_final'])
    every = ensemble.diagnostics(model, whole)
    assert np.array_equal(every.var('proportion'), prop[0])
    for i in (6, 4321, n - 1):
        assert np.array_equal(whole.final[:, i], _oracle_explicit_member(model, i)[-1]), i


def test_chunked_pipeline_full_trajectories_and_stochastic_outputs(monkeypatch):
    # FULL output (rows = T*C) and a stochastic run with swept parameters through the chunks
    n = 5000
    model = configs.seir_sweep(0, n)
    model.make_times(0, 30, 1)
    whole = model.integrate_ensemble('explicit', n_members=n, output='full')
    monkeypatch.setenv('POPDYN_CHUNK_MEMBERS', '1024')
    parts = model.integrate_ensemble('explicit', n_members=n, output='full')
    assert parts.info['n_launch'] == 5 and np.array_equal(parts.soln, whole.soln)
    monkeypatch.delenv('POPDYN_CHUNK_MEMBERS')
    rng = np.random.default_rng(3)
    strains = helpers.make_model('strains', params=[('rate_recover', rng.uniform(5e-3, 2e-2, n))])
    strains.make_times(0, 40, 1)
    a = strains.integrate_ensemble('discrete_time_stochastic', n_members=n, output='final',
                                   seed=4, member_offset=100)
    monkeypatch.setenv('POPDYN_CHUNK_MEMBERS', '1536')
    b = strains.integrate_ensemble('discrete_time_stochastic', n_members=n, output='final',
                                   seed=4, member_offset=100)
    assert b.info['n_launch'] == 4 and np.array_equal(a.final, b.final)
    s = strains.integrate_ensemble('discrete_time_stochastic', n_members=n, output='stats',
                                   seed=4, member_offset=100, hist_bins=4096, hist_max=4095)
    monkeypatch.delenv('POPDYN_CHUNK_MEMBERS')
    t = strains.integrate_ensemble('discrete_time_stochastic', n_members=n, output='stats',
                                   seed=4, member_offset=100, hist_bins=4096, hist_max=4095)
    assert s.info['n_launch'] == 4 and t.info['n_launch'] == 1
    assert np.array_equal(s.moments, t.moments) and np.array_equal(s.hist, t.hist)
    assert int(s.info['n_steps']) == int(t.info['n_steps']) == 40 * n


def test_error_member_is_global_across_chunks(monkeypatch):
    # checks() failing in a later chunk reports the GLOBAL member id
    n = 4000
    # explicit Euler at dt = .05 diverges to NaN for these RMIT parameters (see
    # test_default_checks_report_the_reference_failure_index); every other member is benign
    bad = dict(rho=0.4675096681435942, sigma1=0.4864215594357471, sigma2=0.40346161064467484,
               sigma3=0.40408850758699627, beta=462.931574029995)
    p = dict(helpers.RMIT)
    for key, value in bad.items():
        column = np.full(n, 50. if key == 'beta' else helpers.RMIT.get(key, 0.))
        column[3333] = value
        p[key] = column
    from examples import models
    model = models.RmitTbModel(p)
    model.make_times(0, 500., 1.)
    monkeypatch.setenv('POPDYN_CHUNK_MEMBERS', '1000')
    res = model.integrate_ensemble('explicit', n_members=n, output='final', check=False)
    assert res.info['n_launch'] == 4 and res.info['err_code'] == 1
    assert res.info['err_member'] == 3333


# ------------------------------------------------------------------------------- config 4
def test_config4_tb_gillespie_members_equal_the_oracle():
    model = configs.tb_gillespie()
    model.make_times(1950, 1953, 1.)
    method = 'continuous_time_stochastic'
    n, seed, offset = 96, 0, 12345
    res = model.integrate_ensemble(method, n_members=n, output='full', seed=seed,
                                   member_offset=offset, count_bits=32)
    cm = helpers.compiled(model, method)
    om = oracle.OracleModel(cm)
    for i in (0, 31, 95):
        want, n_event, err = om.gillespie(model.get_init_list(), model.target_times,
                                          oracle.Rng('philox', seed=seed, member=offset + i))
        assert err == 0 and np.array_equal(res.soln[:, :, i], want), i
    final = model.integrate_ensemble(method, n_members=n, output='final', seed=seed,
                                     member_offset=offset, count_bits=64)
    assert np.array_equal(final.final, res.soln[-1])


def test_initial_state_beyond_32_bit_counts_is_rejected():
    model = helpers.make_model('strains')
    model.make_times(0, 5, 1)
    model.set_compartment('susceptible', 3e9)
    with pytest.raises(ValueError, match='count_bits=32'):
        model.integrate_ensemble('discrete_time_stochastic', n_members=64, output='final',
                                 count_bits=32)
    res = model.integrate_ensemble('discrete_time_stochastic', n_members=64, output='final',
                                   count_bits=64)
    assert res.final.sum(axis=0).min() > 2.9e9      # the epidemic moves them, nothing saturates


# ------------------------------------------------------------------------------- config 5b
def test_config5b_hybrid_switching_over_50_days_matches_the_oracle_loop():
    n, seed = 2048, 5
    model = configs.mix_model()
    res = ensemble.integrate_hybrid(model, configs.MIX_SEGMENTS, n, configs.MIX_CUTOFF, seed=seed,
                                    member_offset=7 * n, output='final')
    final = res.final.cpu().numpy()
    modes = res.modes.cpu().numpy()
    assert 0 < modes[1:].sum() < modes[1:].size
    for i in (0, 999, n - 1):
        rows, want_modes = helpers.oracle_hybrid(
            configs.mix_model, configs.MIX_SEGMENTS, configs.MIX_CUTOFF,
            lambda k: oracle.Rng('philox', seed=seed, member=(k << 40) + 7 * n + i))
        assert [int(v) for v in modes[:, i]] == [int(w == 'explicit') for w in want_modes]
        assert np.array_equal(final[:, i], rows[-1]), i


# ------------------------------------------------------------------------------- multi-GPU
def _nccl_worker(rank, world, port, queue):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        def factory(start, stop):
            m = configs.strains()
            m.make_times(0, 60, 1)
            m.device = rank
            return m
        res, (start, stop) = distributed.integrate_sharded(
            factory, 'discrete_time_stochastic', 20001, reduce_device=dev, output='stats',
            seed=9, hist_bins=2048, hist_max=2047, statistics='device', moments_from='hist',
            count_bits=32)
        q = res.quantiles([0.25, 0.5, 0.975]).cpu().numpy()
        if rank == 0:
            queue.put((res.moments.cpu().numpy(), res.hist.cpu().numpy(), q, res.mean(),
                       res.n_total, (start, stop)))
        # a sweep whose failing member lives on the LAST rank: every rank must raise
        def bad_factory(start, stop):
            from examples import models
            bad = dict(rho=0.4675096681435942, sigma1=0.4864215594357471,
                       sigma2=0.40346161064467484, sigma3=0.40408850758699627,
                       beta=462.931574029995)
            p = dict(helpers.RMIT)
            for key, value in bad.items():
                column = np.full(stop - start, 50. if key == 'beta' else helpers.RMIT.get(key, 0.))
                if start <= 3999 < stop:
                    column[3999 - start] = value
                p[key] = column
            m = models.RmitTbModel(p)
            m.make_times(0, 500., 1.)
            m.device = rank
            return m
        try:
            distributed.integrate_sharded(bad_factory, 'explicit', 4000, reduce_device=dev,
                                          output='final')
            raised = None
        except AssertionError as exc:
            raised = str(exc)
        gathered = [None] * world
        dist.all_gather_object(gathered, raised)
        if rank == 0:
            queue.put(gathered)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpus_over_nccl_equal_one_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    moments, hist, q, mean, n_total, shard = queue.get(timeout=300)
    raised = queue.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert n_total == 20001 and shard == (0, 10001)
    m = configs.strains()
    m.make_times(0, 60, 1)
    one = m.integrate_ensemble('discrete_time_stochastic', n_members=20001, output='stats',
                               seed=9, hist_bins=2048, hist_max=2047, moments_from='hist',
                               count_bits=32)
    assert np.array_equal(hist, one.hist.view(np.int32))
    assert np.array_equal(moments, one.moments.view(np.int64))
    assert np.array_equal(q, one.quantiles([0.25, 0.5, 0.975]))
    assert np.array_equal(mean, one.mean())
    assert all(r is not None and 'member 3999' in r for r in raised), raised


def test_chunked_pipeline_with_per_member_initial_states_and_a_member_mask(monkeypatch):
    # y0 [C][M] from host arrays, FULL stochastic output, and a host mask, all through the chunks
    n = 3000
    rng = np.random.default_rng(11)
    def build():
        m = helpers.make_model('strains')
        m.set_compartment('susceptible', np.floor(rng.uniform(800., 1200., n)))
        m.set_compartment('infectious_invader', np.floor(rng.uniform(1., 5., n)))
        m.make_times(0, 25, 1)
        return m
    rng = np.random.default_rng(11)
    whole = build().integrate_ensemble('continuous_time_stochastic', n_members=n, output='full',
                                       seed=2, member_offset=50)
    monkeypatch.setenv('POPDYN_CHUNK_MEMBERS', '700')
    rng = np.random.default_rng(11)
    parts = build().integrate_ensemble('continuous_time_stochastic', n_members=n, output='full',
                                       seed=2, member_offset=50)
    assert parts.info['n_launch'] == 5 and np.array_equal(parts.soln, whole.soln)
    assert int(parts.info['n_steps']) == int(whole.info['n_steps'])
    # masked explicit launch over host buffers: untouched members keep what the buffer held
    rng = np.random.default_rng(11)
    m = build()
    mask = (np.arange(n) % 3 == 0).astype(np.uint8)
    out = np.full((5, n), -7.0)
    cm = helpers.compiled(m, 'explicit')
    res = ensemble.run_compiled(m, cm, 'explicit', m.get_init_list(), m.target_times, n,
                                output='final', out=out, mask=mask, mask_value=1)
    assert res.info['n_launch'] == 5
    assert np.all(out[:, mask == 0] == -7.0) and np.all(out[:, mask == 1] >= 0.0)
    monkeypatch.delenv('POPDYN_CHUNK_MEMBERS')
    rng = np.random.default_rng(11)
    ref = build().integrate_ensemble('explicit', n_members=n, output='final')
    assert np.array_equal(out[:, mask == 1], ref.final[:, mask == 1])


def test_two_threads_share_one_kernel_handle_without_sharing_an_error_flag():
    # the error block, events and staging scratch belong to a per-run slot: concurrent runs of the
    # same cached handle (ctypes releases the GIL) must neither lose nor misattribute a failure
    import threading
    from examples import models
    n = 4096
    bad = dict(rho=0.4675096681435942, sigma1=0.4864215594357471, sigma2=0.40346161064467484,
               sigma3=0.40408850758699627, beta=462.931574029995)

    def build(fail):
        p = dict(helpers.RMIT)
        for key, value in bad.items():
            column = np.full(n, 50. if key == 'beta' else helpers.RMIT.get(key, 0.))
            if fail:
                column[n // 2] = value
            p[key] = column
        m = models.RmitTbModel(p)
        m.make_times(0, 500., 1.)
        return m

    good, failing = build(False), build(True)
    want = good.integrate_ensemble('explicit', n_members=n, output='final').final
    outcomes = {'good': [], 'failing': []}

    def work(name, model):
        for _ in range(6):
            res = model.integrate_ensemble('explicit', n_members=n, output='final', check=False)
            outcomes[name].append((res.info['err_code'], res.info['err_member'],
                                   np.array_equal(res.final, want) if name == 'good' else True))

    threads = [threading.Thread(target=work, args=('good', good)),
               threading.Thread(target=work, args=('failing', failing))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert outcomes['good'] == [(0, 0, True)] * 6
    assert [o[:2] for o in outcomes['failing']] == [(1, n // 2)] * 6


def test_final_vars_fused_into_the_ode_kernels_equal_the_diagnostics_pass(golden):
    # config 5's output: model.vars['proportion'] of the final state, straight from the kernel
    n = 6000
    model = configs.rmit_sweep(0, n)
    fused = model.integrate_ensemble('explicit', n_members=n, final_vars=['proportion'])
    assert fused.output == 'final_vars' and fused.final_vars.shape == (1, n)
    assert np.array_equal(fused.final_var('proportion')[:6], golden['rmit_fig2_proportion'])
    state = model.integrate_ensemble('explicit', n_members=n, output='final')
    two_pass = ensemble.diagnostics(model, state, want_vars=['proportion'])
    assert np.array_equal(fused.final_vars, two_pass.vars)
    # several vars, the SEIR-demography sweep, through the chunk pipeline
    seir = configs.seir_sweep(0, 3000)
    seir.make_times(0, 400, 1)
    import os
    os.environ['POPDYN_CHUNK_MEMBERS'] = '1024'
    try:
        fused = seir.integrate_ensemble('explicit', n_members=3000,
                                        final_vars=['prevalence', 'population', 'incidence'])
    finally:
        del os.environ['POPDYN_CHUNK_MEMBERS']
    assert fused.info['n_launch'] == 3
    state = seir.integrate_ensemble('explicit', n_members=3000, output='final')
    every = ensemble.diagnostics(seir, state)
    for label in ('prevalence', 'population', 'incidence'):
        assert np.array_equal(fused.final_var(label), every.var(label)), label
    # scale-up vars keep the value of the last derivative call (TB model, drop-in var_array)
    tb = helpers.make_model('tb')
    tb.make_times(1900, 2050, .05)
    tb.integrate('explicit')
    assert np.array_equal(tb.var_array[::50], golden['tb_var_array_every50'])
    var_last, var_labels = tb.var_array[-1].copy(), list(tb.var_labels)
    labels = ['prevalence', 'incidence', 'mortality', 'program_rate_detect']
    fused = tb.integrate_ensemble('explicit', n_members=3, final_vars=labels)   # clears var_array
    for label in labels:
        assert np.all(fused.final_var(label) == var_last[var_labels.index(label)]), label
    # the adaptive method takes the same output
    a = seir.integrate_ensemble('scipy', n_members=3000, final_vars=['prevalence'])
    b = seir.integrate_ensemble('scipy', n_members=3000, output='final')
    pop = b.final.sum(axis=0)
    assert np.allclose(a.final_var('prevalence'), b.final[2] / pop, rtol=1e-14, atol=0)
