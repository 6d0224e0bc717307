"""Leaf -> rank map and query slicing for one-process-per-GPU runs.

The fit shards by leaves (contiguous ranges = contiguous sub-trees of the BSP, so every rank's factor spans are
contiguous in the packed buffers); the query shards by query index with no data-path collective; results are
gathered once at the end.  Pure index arithmetic + torch.distributed calls, backend-agnostic (NCCL on GPUs,
gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple


def leaf_range(rank: int, world: int, n_leaves: int) -> Tuple[int, int]:
    """(first_leaf 0-based, count) factorised by `rank`."""
    a = (n_leaves * rank) // world
    b = (n_leaves * (rank + 1)) // world
    return a, b - a


def query_slice(rank: int, world: int, nq: int) -> Tuple[int, int]:
    """[first, last) of the queries answered by `rank`."""
    return (nq * rank) // world, (nq * (rank + 1)) // world


def all_ranges(world: int, n: int, fn) -> List[Tuple[int, int]]:
    return [fn(r, world, n) for r in range(world)]


def gather_slices(local, nq: int, group=None):
    """all-gather the ranks' result slices (possibly of unequal length) into the full length-nq tensor."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    a, b = query_slice(rank, world, nq)
    assert local.shape[0] == b - a
    if nq % world == 0:
        out = torch.empty(nq, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # uneven slices (they differ by at most one element): pad to the longest, gather, trim
    sizes = [query_slice(r, world, nq)[1] - query_slice(r, world, nq)[0] for r in range(world)]
    m = max(sizes)
    padded = torch.zeros(m, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty(m * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * m:r * m + sizes[r]] for r in range(world)])


def exchange_spans(span_of, n_leaves: int, buffers, group=None):
    """After a sharded fit: every rank sends the device spans it factorised to every peer and receives theirs.
    span_of(which, first_leaf, count) -> 1-D tensor aliasing that span of the local model (may be empty).
    All transfers of one buffer go out as ONE batch of point-to-point operations (a single NCCL group), so they
    run concurrently over NVLink / NVSwitch in both directions instead of as serialised broadcasts."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return
    for which in buffers:
        a, n = leaf_range(rank, world, n_leaves)
        mine = span_of(which, a, n)
        ops = []
        for r in range(world):
            if r == rank:
                continue
            ar, nr = leaf_range(r, world, n_leaves)
            theirs = span_of(which, ar, nr)
            if theirs.numel():
                ops.append(dist.P2POp(dist.irecv, theirs, r, group))
            if mine.numel():
                ops.append(dist.P2POp(dist.isend, mine, r, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
