# -*- coding: utf-8 -*-
"""
Multi-GPU sharding of an ensemble: one process per GPU (``torch.distributed``, NCCL over NVLink).

Members are independent (no cross-member term anywhere in /root/reference/basepop.py:658-852), so
the path shards embarrassingly: rank ``g`` of ``G`` integrates the contiguous global member range
``shard_range(M, g, G)`` with no data-path collective.  Philox streams are keyed by the GLOBAL
member id, hence every statistic is bit-identical for any ``G``.  The only exchange step is the
final reduction of the ensemble statistics: exact integer moments ``[T][C][2]`` and histograms
``[rows][C][bins]`` are summed with one all-reduce each; mean / variance / quantiles are then
derived on every rank from the reduced integers.
"""

import numpy


def shard_range(n_members, rank, world_size):
    """Contiguous block partition [start, stop) of the global member ids; the first
    ``n_members % world_size`` ranks take one extra member."""
    base, extra = divmod(int(n_members), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_rows(array, rank, world_size, axis=-1):
    """Slice of a per-member array (parameter sweep, initial state) owned by ``rank``."""
    start, stop = shard_range(array.shape[axis], rank, world_size)
    index = [slice(None)] * array.ndim
    index[axis] = slice(start, stop)
    return array[tuple(index)]


def _as_signed_tensor(buf, device):
    """uint64 / uint32 accumulators viewed as int64 / int32 torch tensors (two's-complement sums
    wrap identically, and NCCL has no unsigned 64-bit type in torch)."""
    import torch
    if hasattr(buf, 'data_ptr'):
        return buf, None
    signed = {numpy.dtype('uint64'): numpy.int64, numpy.dtype('uint32'): numpy.int32}
    view = buf.view(signed.get(buf.dtype, buf.dtype))
    tensor = torch.from_numpy(view)
    if device is not None:
        return tensor.to(device), buf
    return tensor, buf


def allreduce_stats(result, group=None, device=None):
    """Sum ``result.moments`` and ``result.hist`` over all ranks, in place.  ``device`` is the
    torch device used for the collective (a CUDA device for the NCCL backend, None for gloo)."""
    import torch.distributed as dist
    for name in ('moments', 'hist'):
        buf = getattr(result, name)
        if buf is None:
            continue
        tensor, host = _as_signed_tensor(buf, device)
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
        if host is not None:
            reduced = tensor.cpu().numpy()
            host[...] = reduced.view(host.dtype)
    return result


def allgather_error(info, group=None, device=None):
    """Make the device error flag of any rank known to every rank: returns the ``info`` triple
    (err_code, err_member, err_where) of the failing member with the smallest global id, or
    zeros.  Every rank must call it (it is a collective)."""
    import torch
    import torch.distributed as dist
    mine = torch.tensor([int(info.get('err_code', 0)), int(info.get('err_member', -1)),
                         int(info.get('err_where', 0))], dtype=torch.int64, device=device)
    world = dist.get_world_size(group)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    failures = [tuple(int(v) for v in t.cpu()) for t in gathered if int(t[0]) != 0]
    if not failures:
        return dict(err_code=0, err_member=-1, err_where=0)
    code, member, where = min(failures, key=lambda f: f[1])
    return dict(err_code=code, err_member=member, err_where=where)


def integrate_sharded(model_factory, method, n_members, rank=None, world_size=None, group=None,
                      reduce_device=None, check=True, **kwargs):
    """Integrate this rank's shard of an ``n_members`` ensemble and all-reduce the statistics.

    ``model_factory(start, stop)`` builds the model for the global member range [start, stop)
    (so that per-member parameter arrays are created for the shard only; it is also called for an
    EMPTY range, ``start == stop``, when there are more ranks than members -- that rank launches
    nothing but still joins the collectives).  Keyword arguments go to
    ``ensemble.integrate_ensemble``.  Returns ``(result, (start, stop))``; with ``output='stats'``
    the moments / histograms in ``result`` are the ensemble totals and ``result.n_total`` is the
    global member count, so ``mean()`` / ``var()`` / ``quantiles()`` describe the whole ensemble.

    Errors are collective too: the device error flag of every rank is gathered BEFORE the
    statistics are reduced and, with ``check=True``, the reference's exception is raised on every
    rank -- a failing shard can never leave the other ranks waiting in an all-reduce."""
    import torch.distributed as dist
    from . import ensemble
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    start, stop = shard_range(n_members, rank, world_size)
    model = model_factory(start, stop)
    if stop > start:
        result = ensemble.integrate_ensemble(model, method, n_members=stop - start,
                                             member_offset=start, check=False, **kwargs)
    else:
        result = ensemble.empty_result(model, method, member_offset=start, **kwargs)
    info = dict(result.info)
    if world_size > 1 and dist.is_initialized():
        info = allgather_error(info, group=group, device=reduce_device)
        result.info.update(info)
        if kwargs.get('output') == 'stats' and not info['err_code']:
            allreduce_stats(result, group=group, device=reduce_device)
            if getattr(result, 'moments_from', None) == 'hist':
                result.derive_moments()
    result.n_total = int(n_members)
    if check:
        ensemble.raise_device_error(info, method)
    return result, (start, stop)
