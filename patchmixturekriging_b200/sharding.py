"""Sub-tree ownership: the leaf -> rank map, the query slices and the routing bookkeeping of a sharded query.

Inside one box this is done by the library itself (pmk_multi_*, csrc/pmk_multi.cu: host threads + CUDA peer copies).  This
module restates the same index arithmetic in Python -- pinned to the library's host-only exports by tests/test_multi.py --
and carries the SAME flow over torch.distributed point-to-point operations (NCCL between boxes, gloo in the CPU tests) for a
host layer that runs one process per GPU on top of the single-GPU building blocks (pmk_set_leaf_base,
pmk_query_plan_segments / _pack_dev, pmk_query_pairs_routed_dev, pmk_query_plan_unpack_dev):

  fit   : rank r fits leaves leaf_range(r) (equal counts; balanced_first_leaf gives pmk_multi's cost-balanced ranges); no exchange.
  query : rank r plans queries query_slice(r); its (query, leaf) pairs, sorted by leaf, form one contiguous segment per
          owner (segments); the segments travel to the owners (exchange_segments), the owners answer them in the order
          received, the answers travel back (return_segments), the planner combines in the reference's slot order.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def leaf_range(rank: int, world: int, n_leaves: int) -> Tuple[int, int]:
    """(first_leaf 0-based, count) owned by `rank` -- pmk_multi_leaf_range."""
    a = (n_leaves * rank) // world
    b = (n_leaves * (rank + 1)) // world
    return a, b - a


def balanced_first_leaf(world: int, leaf_sizes: Sequence[int]) -> np.ndarray:
    """world + 1 ascending 0-based leaf ids: contiguous ranges of nearly equal cost sum(n^3) -- pmk_multi's own leaf -> rank
    map (pmk_multi_balanced_ranges / pmk_multi_owned_range): boundary i is the leaf whose cost prefix is nearest to i / world
    of the total, every rank keeping at least one leaf."""
    n = np.asarray(leaf_sizes, dtype=np.float64)
    n_leaves = len(n)
    assert n_leaves >= world >= 1
    pre = np.concatenate([[0.0], np.cumsum(n * n * n)])       # same left-to-right sum as the library
    bnd = np.zeros(world + 1, dtype=np.int64)
    bnd[world] = n_leaves
    for i in range(1, world):
        target = pre[n_leaves] * float(i) / float(world)
        k = int(np.searchsorted(pre, target, side="left"))
        if k > 0 and target - pre[k - 1] < pre[k] - target:
            k -= 1
        k = max(k, int(bnd[i - 1]) + 1)
        k = min(k, n_leaves - (world - i))
        bnd[i] = k
    return bnd


def query_slice(rank: int, world: int, nq: int) -> Tuple[int, int]:
    """[first, last) of the queries planned by `rank` -- pmk_multi_query_range."""
    return (nq * rank) // world, (nq * (rank + 1)) // world


def all_ranges(world: int, n: int, fn) -> List[Tuple[int, int]]:
    return [fn(r, world, n) for r in range(world)]


def owner_first_leaf(world: int, n_leaves: int) -> np.ndarray:
    """world + 1 ascending 0-based leaf ids: owner o holds leaves [f[o], f[o+1])."""
    return np.array([leaf_range(r, world, n_leaves)[0] for r in range(world)] + [n_leaves], dtype=np.int64)


def segments(sorted_leaf: np.ndarray, first_leaf: Sequence[int]) -> np.ndarray:
    """Offsets (len(first_leaf) entries) of every owner's segment in a pair list sorted by 1-based leaf id --
    pmk_query_plan_segments."""
    return np.searchsorted(np.asarray(sorted_leaf), np.asarray(first_leaf) + 1, side="left").astype(np.int64)


def rx_offsets(seg_all: np.ndarray, owner: int) -> np.ndarray:
    """seg_all[s] = segments of planner s.  Offsets (world + 1) at which `owner` stores what each planner sends it."""
    seg_all = np.asarray(seg_all)
    counts = seg_all[:, owner + 1] - seg_all[:, owner]
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)


# --- the same flow over torch.distributed (one process per rank) ------------------------------------------------------------
def _all_gather_rows(row, group=None):
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [torch.empty_like(row) for _ in range(world)]
    dist.all_gather(out, row, group=group)
    return torch.stack(out)


def exchange_segments(send, seg, group=None):
    """Planner -> owner.  `send`: tensor whose first dimension is the planner's leaf-sorted pair list, seg: its world + 1
    segment offsets.  Returns (received, rx_off, seg_all): what the planners sent THIS rank, concatenated in planner
    order, the offsets of every planner's share, and all ranks' segment tables."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    seg_t = torch.as_tensor(np.asarray(seg, dtype=np.int64), device=send.device)
    seg_all = _all_gather_rows(seg_t, group).cpu().numpy()
    rx_off = rx_offsets(seg_all, rank)
    recv = torch.empty((int(rx_off[-1]),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    ops = []
    for peer in range(world):
        mine = send[int(seg[peer]):int(seg[peer + 1])]
        theirs = recv[int(rx_off[peer]):int(rx_off[peer + 1])]
        if peer == rank:
            theirs.copy_(mine)
            continue
        if theirs.shape[0]:
            ops.append(dist.P2POp(dist.irecv, theirs, peer, group))
        if mine.shape[0]:
            ops.append(dist.P2POp(dist.isend, mine.contiguous(), peer, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return recv, rx_off, seg_all


def return_segments(answers, rx_off, seg_all, group=None):
    """Owner -> planner: the inverse of exchange_segments.  `answers` is in the order received; the result is in the
    planner's leaf-sorted pair order."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    seg = seg_all[rank]
    out = torch.empty((int(seg[-1]),) + tuple(answers.shape[1:]), dtype=answers.dtype, device=answers.device)
    ops = []
    for peer in range(world):
        mine = answers[int(rx_off[peer]):int(rx_off[peer + 1])]
        theirs = out[int(seg[peer]):int(seg[peer + 1])]
        if peer == rank:
            theirs.copy_(mine)
            continue
        if theirs.shape[0]:
            ops.append(dist.P2POp(dist.irecv, theirs, peer, group))
        if mine.shape[0]:
            ops.append(dist.P2POp(dist.isend, mine.contiguous(), peer, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out
