"""Multi-rank host logic on CPU: world_size-2 gloo processes shard an ensemble by global member
id and all-reduce the integer statistics.  The per-shard work is done by the oracle here (no GPU
in this container); the sharding / reduction code is the product's (popdynamics_b200.distributed).
Philox streams are keyed by the global member id, so the reduced statistics must equal the
single-process ones bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from popdynamics_b200 import distributed
from popdynamics_b200.ensemble import EnsembleResult

import helpers


def test_shard_range_partitions_exactly():
    for n, world in ((10, 3), (7, 8), (10**7, 8), (5, 1)):
        ranges = [distributed.shard_range(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
    arr = np.arange(20).reshape(2, 10)
    assert np.array_equal(distributed.shard_rows(arr, 1, 3), arr[:, 4:7])


def _shard_stats(rank, world, n_member, seed, bins):
    m = helpers.make_model('strains')
    m.make_times(0, 40, 1)
    om = oracle.OracleModel(helpers.compiled(m, 'discrete_time_stochastic'))
    start, stop = distributed.shard_range(n_member, rank, world)
    out = om.stochastic_batch('discrete_time_stochastic', stop - start, m.get_init_list(),
                              m.target_times, seed, member0=start, want_soln=True,
                              want_moments=True, n_thread=2)
    res = EnsembleResult()
    res.labels = list(m.labels)
    res.moments = np.stack([out['sum'], out['sumsq']], axis=-1).astype(np.uint64)
    counts = out['soln'].astype(np.int64)                      # [M][T][C]
    hist = np.zeros((counts.shape[1], counts.shape[2], bins), dtype=np.uint32)
    for t in range(counts.shape[1]):
        for c in range(counts.shape[2]):
            hist[t, c] = np.bincount(np.minimum(counts[:, t, c], bins - 1), minlength=bins)
    res.hist = hist
    return res


def _worker(rank, world, port, n_member, seed, bins, queue):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        res = _shard_stats(rank, world, n_member, seed, bins)
        distributed.allreduce_stats(res)
        if rank == 0:
            queue.put((res.moments.copy(), res.hist.copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_reduction_equals_single_process():
    n_member, seed, bins = 96, 2024, 4096
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_member, seed, bins, queue))
             for r in range(2)]
    for p in procs:
        p.start()
    moments, hist = queue.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _shard_stats(0, 1, n_member, seed, bins)
    assert np.array_equal(moments, single.moments)
    assert np.array_equal(hist, single.hist)
    assert int(hist[3, 0].sum()) == n_member
