"""CPU tests of the round-2 host logic: shard-independent sweep builders, the want_vars subset of
the diagnostics pass, the collective error / empty-shard handling of integrate_sharded (gloo,
world size 2, the per-shard integration replaced by a stub), the reference runner built into
oracle/_ref, and the generated kernel-source header being in sync."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from examples import configs, models
from popdynamics_b200 import build, compiler, distributed, ensemble

import helpers


def test_sweep_values_are_a_function_of_the_global_member_id():
    a = configs.sweep_uniform(0, 3, 2., 18., 0, 2 * configs.BLOCK + 5)
    b = configs.sweep_uniform(0, 3, 2., 18., configs.BLOCK - 3, configs.BLOCK + 9)
    assert np.array_equal(a[configs.BLOCK - 3:configs.BLOCK + 9], b)
    assert a.min() >= 2. and a.max() < 18.
    assert not np.array_equal(configs.sweep_uniform(0, 4, 2., 18., 0, 100), a[:100])
    m = configs.seir_sweep(0, 10)
    assert m.params['r0'][0] == 13. and m.params['infection_rate_recover'][0] == 1. / 7.
    late = configs.seir_sweep(5, 10)
    assert np.array_equal(late.params['r0'], m.params['r0'][5:])
    r = configs.rmit_sweep(2, 9)
    assert list(r.params['beta'][:4]) == [20., 99., 250., 500.] and r.params['sigma1'][0] == 0.


def test_diagnostics_want_vars_keeps_the_reference_order_and_rejects_unknown_labels():
    m = helpers.make_model('rmit')
    m.check_flow_transfers()
    full = compiler.compile_model(m, 'explicit', diagnostics=True)
    assert 'proportion' in full.diag_labels and len(full.diag_labels) > 3
    one = compiler.compile_model(m, 'explicit', diagnostics=True, want_vars=['proportion'])
    assert one.diag_labels == ['proportion']
    assert '#define PD_NDIAG 1' in one.cuda_source()
    two = compiler.compile_model(m, 'explicit', diagnostics=True,
                                 want_vars=['proportion', 'population'])
    assert two.diag_labels == [l for l in full.diag_labels if l in ('proportion', 'population')]
    with pytest.raises(KeyError):
        compiler.compile_model(m, 'explicit', diagnostics=True, want_vars=['nope'])


def test_embedded_kernel_sources_are_in_sync():
    """csrc/pd_embedded.h is generated from csrc/kernels/* at build time (and not tracked): a
    stale copy would make NVRTC compile something other than the sources under review."""
    build.write_embedded()
    text = open(build.EMBEDDED).read()
    for name in build.KERNEL_FILES:
        assert open(os.path.join(build.KERNEL_DIR, name)).read() in text
    assert build.source_digest() in text


def _stub_integrate(model, method, n_members=None, member_offset=0, check=True, **kw):
    """Stands in for the GPU launch: statistics that are a pure function of the global ids."""
    res = ensemble.EnsembleResult()
    res.method, res.output, res.labels = method, kw.get('output'), list(model.labels)
    res.n_members, res.member_offset = n_members, member_offset
    ids = np.arange(member_offset, member_offset + n_members, dtype=np.uint64)
    T, C = len(model.target_times), len(model.labels)
    res.moments = np.zeros((T, C, 2), np.uint64)
    res.moments[0, 0, 0] = ids.sum()
    res.moments[0, 0, 1] = (ids * ids).sum()
    res.hist = np.zeros((T, C, 8), np.uint32)
    np.add.at(res.hist[0, 0], (ids % 8).astype(int), 1)
    res.hist_times, res.hist_shift, res.hist_bins = np.arange(T), np.zeros(C, np.int32), 8
    bad = model.params.get('bad_member', -1)
    res.info = dict(err_code=1 if member_offset <= bad < member_offset + n_members else 0,
                    err_member=bad, err_where=3)
    res.moments_from = None
    return res


def _sharded_worker(rank, world, port, n_member, bad, queue):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ensemble.integrate_ensemble = _stub_integrate

        def factory(start, stop):
            m = models.SirModel({})
            m.make_times(0, 1, 1)
            m.params['bad_member'] = bad
            return m
        try:
            res, shard = distributed.integrate_sharded(factory, 'discrete_time_stochastic',
                                                       n_member, output='stats', hist_bins=8,
                                                       hist_max=7)
            queue.put((rank, 'ok', res.moments.copy(), res.hist.copy(), res.n_total,
                       float(res.mean()[0, 0]), shard))
        except AssertionError as exc:
            queue.put((rank, 'raised', str(exc)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run_sharded(n_member, bad):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, n_member, bad, queue))
             for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([queue.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


def test_integrate_sharded_reduces_and_reports_the_global_member_count():
    got = _run_sharded(11, -1)
    ids = np.arange(11, dtype=np.uint64)
    for rank, status, moments, hist, n_total, mean, shard in got:
        assert status == 'ok' and n_total == 11
        assert moments[0, 0, 0] == ids.sum() and moments[0, 0, 1] == (ids * ids).sum()
        assert hist.sum() == 11 and mean == ids.sum() / 11.0
    assert got[0][6] == (0, 6) and got[1][6] == (6, 11)


def test_integrate_sharded_with_an_empty_shard_still_joins_the_collectives():
    got = _run_sharded(1, -1)                     # rank 1 owns no member
    for rank, status, moments, hist, n_total, mean, shard in got:
        assert status == 'ok' and n_total == 1 and hist.sum() == 1
    assert got[1][6] == (1, 1)


def test_a_failing_shard_raises_on_every_rank_instead_of_hanging():
    got = _run_sharded(10, 7)                     # member 7 lives on rank 1
    assert [g[1] for g in got] == ['raised', 'raised']
    assert all('member 7' in g[2] for g in got)


def test_reference_runner_loads_the_compiled_reference():
    from oracle import build_ref, ref_runner
    if not ref_runner.available() and not build_ref.build():
        pytest.skip('no /root/reference here and no prebuilt oracle/_ref')
    ref = ref_runner.basepop()
    cls = ref_runner.example_classes('stochastic_strains_model')['StrainsModel']
    assert issubclass(cls, ref.BaseModel)
    import random
    random.seed(1)
    np.random.seed(1)
    m = cls()
    m.make_times(0, 30, 1)
    m.integrate('discrete_time_stochastic')
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_vectors.npz'))
    assert np.array_equal(m.soln_array, g['strains_dts_seed1'][:31])


def test_model_size_limits_raise_clear_errors_before_any_compilation():
    from popdynamics_b200.basepop import BaseModel

    class Wide(BaseModel):
        def __init__(self, n):
            BaseModel.__init__(self)
            self.n = n
            self.set_param('k', 0.1)
            for i in range(n):
                self.set_compartment('c%d' % i, 10.)

        def set_flows(self):
            for i in range(self.n):
                self.set_fixed_transfer_rate_flow('c%d' % i, 'c%d' % ((i + 1) % self.n), 'k')

    m = Wide(70)
    m.make_times(0, 2, 1)
    with pytest.raises(ValueError, match='at most 64 flow-table rows'):
        m.integrate_ensemble('discrete_time_stochastic', n_members=8, output='final')
    with pytest.raises(ValueError, match='at most 64 compartments'):
        m.integrate_ensemble('continuous_time_stochastic', n_members=8, output='stats')


def test_final_vars_are_traced_the_way_the_diagnostics_pass_runs_them():
    from popdynamics_b200 import capi
    m = helpers.make_model('rmit')
    m.check_flow_transfers()
    cm = compiler.compile_model(m, 'explicit', final_vars=['proportion', 'population'])
    assert cm.final_var_labels == ['proportion', 'population']
    src = cm.cuda_source()
    assert '#define PD_NFVAR 2' in src and 'pd_final_vars(' in src
    kernel = capi.Model(src, capi.EXPLICIT, capi.FINAL_VARS, capi.PHILOX, 0, cm.n_comp,
                        cm.n_param, 0, cm.n_su, cm.n_flow)
    assert kernel.cubin()[:4] == b'\x7fELF'
    kernel.close()
    # same nodes as the separate diagnostics compile reaches (shared leaves, plain sum fold)
    diag = compiler.compile_model(m, 'explicit', diagnostics=True, want_vars=['proportion'])
    assert '#define PD_NDIAG 1' in diag.cuda_source()
    with pytest.raises(KeyError):
        compiler.compile_model(m, 'explicit', final_vars=['nope'])
    with pytest.raises(ValueError, match='ODE methods'):
        s = helpers.make_model('strains')
        s.check_flow_transfers()
        compiler.compile_model(s, 'discrete_time_stochastic', final_vars=['population'])
    with pytest.raises(capi.PopdynError):              # PD_FINAL_VARS is for the ODE kernels
        s = helpers.make_model('strains')
        s.check_flow_transfers()
        cs = compiler.compile_model(s, 'discrete_time_stochastic')
        capi.Model(cs.cuda_source(), capi.TAU_LEAP, capi.FINAL_VARS, capi.PHILOX, 0, cs.n_comp,
                   cs.n_param, 0, 0, cs.n_flow)
