"""Sub-tree ownership on CPU: the leaf -> rank map and query slices (Python mirror == the library's host-only exports),
the routing bookkeeping (segments / receive offsets, simulated for several rank counts), and the whole sharded query flow
between two real processes (gloo, world_size 2): every rank owns half of the leaves of an oracle-fitted model, plans half of
the queries, the pairs travel to the owners and back, and the result equals the single-process reference loop bit for bit.
The per-pair compute is stood in for by the oracle (test infrastructure); the GPU form of the same flow is pmk_multi
(tests/test_gpu_multi.py, bench.py --gpus N)."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np
import pytest

from patchmixturekriging_b200 import _lib, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n", [1, 7, 64, 4096, 10_000_001])
def test_ranges_partition_everything(world, n):
    lr = sharding.all_ranges(world, n, sharding.leaf_range)
    assert lr[0][0] == 0 and sum(c for _, c in lr) == n
    for (a, c), (a2, _) in zip(lr[:-1], lr[1:]):
        assert a + c == a2
    qs = sharding.all_ranges(world, n, sharding.query_slice)
    assert qs[0][0] == 0 and qs[-1][1] == n
    for (_, b), (a2, _) in zip(qs[:-1], qs[1:]):
        assert b == a2
    sizes = [b - a for a, b in qs]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("n", [8, 64, 4096, 8192, 10_000_001])
def test_python_mirror_equals_library_map(built_lib, world, n):
    """pmk_multi_leaf_range / pmk_multi_query_range are host-only (no GPU needed): the map the library shards by is the
    one this package documents; with a power-of-two rank count the ranges are the sub-trees below the top log2(world) levels."""
    L = _lib.lib()
    a, c = C.c_int64(0), C.c_int64(0)
    for r in range(world):
        assert L.pmk_multi_leaf_range(world, n, r, C.byref(a), C.byref(c)) == _lib.PMK_OK
        assert (a.value, c.value) == sharding.leaf_range(r, world, n)
        assert L.pmk_multi_query_range(world, n, r, C.byref(a), C.byref(c)) == _lib.PMK_OK
        assert (a.value, a.value + c.value) == sharding.query_slice(r, world, n)
    assert L.pmk_multi_leaf_range(world, n, world, C.byref(a), C.byref(c)) == _lib.PMK_ERR_ARG
    if world in (2, 8) and n in (64, 4096, 8192):
        f = sharding.owner_first_leaf(world, n)
        assert np.array_equal(f, np.arange(world + 1) * (n // world))       # aligned sub-trees


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_balanced_ranges_host_only(built_lib, world):
    """pmk_multi_balanced_ranges (host only): pmk_multi's cost-balanced leaf -> rank map -- contiguous ranges that cover the
    leaves once, every rank at least one leaf, every range within one leaf of its share of sum(n^3); the Python mirror gives the
    same boundaries; equal leaves give the equal-count split (the sub-trees of the BSP)."""
    L = _lib.lib()
    rng = np.random.default_rng(5 + world)
    for sizes in (rng.integers(300, 1500, 4096), rng.integers(1, 2049, 37), np.full(64, 512), np.arange(1, world + 1)):
        sizes = np.asarray(sizes, dtype=np.int64)
        leaf_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        first = np.zeros(world + 1, dtype=np.int64)
        assert L.pmk_multi_balanced_ranges(world, len(sizes), _lib.ptr(leaf_off), _lib.ptr(first)) == _lib.PMK_OK
        assert first[0] == 0 and first[-1] == len(sizes) and np.all(np.diff(first) >= 1)
        assert np.array_equal(first, sharding.balanced_first_leaf(world, sizes))
        cost = sizes.astype(float) ** 3
        worst = max(cost[a:b].sum() for a, b in zip(first[:-1], first[1:]))
        assert worst <= cost.sum() / world + cost.max() * (1 + 1e-12)
        if np.all(sizes == sizes[0]) and len(sizes) % world == 0:
            assert np.array_equal(first, sharding.owner_first_leaf(world, len(sizes)))
    bad = np.zeros(world + 1, dtype=np.int64)
    assert L.pmk_multi_balanced_ranges(world + 1, world, _lib.ptr(np.arange(world + 1, dtype=np.int64)), _lib.ptr(bad)) == _lib.PMK_ERR_ARG


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_routing_round_trip_in_numpy(world):
    """Every pair reaches the owner of its leaf and its answer comes back to its own slot: segments() / rx_offsets() with the
    copies of pmk_multi_query_staged (owner pulls planner segments, planner pulls answers) simulated in numpy."""
    rng = np.random.default_rng(world)
    n_leaves = 37
    first = sharding.owner_first_leaf(world, n_leaves)
    plans = []
    for s in range(world):
        npairs = int(rng.integers(0, 400)) if s != 1 else 0                  # one planner without pairs
        leaf = np.sort(rng.integers(1, n_leaves + 1, size=npairs)).astype(np.int32)
        payload = rng.standard_normal(npairs)
        plans.append((leaf, payload, sharding.segments(leaf, first)))
    seg_all = np.stack([p[2] for p in plans])
    assert np.array_equal(seg_all[:, 0], np.zeros(world)) and np.array_equal(seg_all[:, -1], [len(p[0]) for p in plans])
    answers = []
    for o in range(world):                                                   # owner side
        off = sharding.rx_offsets(seg_all, o)
        rx_leaf = np.concatenate([plans[s][0][seg_all[s, o]:seg_all[s, o + 1]] for s in range(world)])
        rx_val = np.concatenate([plans[s][1][seg_all[s, o]:seg_all[s, o + 1]] for s in range(world)])
        assert rx_leaf.shape[0] == off[-1]
        assert np.all((rx_leaf - 1 >= first[o]) & (rx_leaf - 1 < first[o + 1]))      # only leaves this owner holds
        answers.append((off, 2.0 * rx_val + rx_leaf))                        # "u" of every received pair
    for s in range(world):                                                   # planner side
        leaf, payload, seg = plans[s]
        back = np.concatenate([answers[o][1][answers[o][0][s]:answers[o][0][s + 1]] for o in range(world)])
        assert np.array_equal(back, 2.0 * payload + leaf)


def _worker(rank, world, port, nq, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import cases
    import helpers
    from oracle import pmk_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = cases.mixgp_file()
    Xq = case["Xq"][::7][:nq]
    th, _ = helpers.kernels(case["kernel"])
    wth, _ = helpers.kernels(case["wkernel"])
    m = helpers.oracle_model(case)
    eta = m["eta"]
    n_leaves = len(eta.X_parts)
    a, c = sharding.leaf_range(rank, world, n_leaves)
    for p in range(n_leaves):             # this rank OWNS leaves [a, a+c): the others' state is simply not there
        if not (a <= p < a + c):
            eta.c_set[p] = None
            eta.L_set[p] = None
    # plan my slice of the queries (needs the tree only)
    q0, q1 = sharding.query_slice(rank, world, nq)
    Xs = Xq[q0:q1]
    home, pq, ph, pl, pt = O.query_structure_vec(Xs, m["hv"], m["hc"], case["levels"], case["radius"], case["delta"])
    nloc = Xs.shape[0]
    allq = np.concatenate([pq, np.arange(nloc)])
    alll = np.concatenate([pl, home])
    allw = np.concatenate([O.evalkernel_tau(np.abs(pt), wth), np.ones(nloc)])
    slot = np.concatenate([np.zeros(pq.shape[0], dtype=np.int64), np.ones(nloc, dtype=np.int64)])
    o = np.lexsort((np.concatenate([ph, np.zeros(nloc, dtype=np.int64)]), slot, allq))     # reference slot order
    allq, alll, allw = allq[o], alll[o], allw[o]
    by_leaf = np.argsort(alll, kind="stable")                                              # pair ids sorted by leaf
    seg = sharding.segments(alll[by_leaf], sharding.owner_first_leaf(world, n_leaves))
    send = torch.from_numpy(np.column_stack([Xs[allq[by_leaf]], alll[by_leaf].astype(np.float64)]))
    recv, rx_off, seg_all = sharding.exchange_segments(send, seg)
    # owner: answer what arrived, in the order it arrived
    ans = np.empty((recv.shape[0], 2))
    rn = recv.numpy()
    for k in range(rn.shape[0]):
        leaf = int(rn[k, -1])
        assert a <= leaf - 1 < a + c
        ans[k] = O.queryinner(rn[k, :-1], eta.X_parts[leaf - 1], th, eta.c_set[leaf - 1], eta.L_set[leaf - 1])
    back = sharding.return_segments(torch.from_numpy(ans), rx_off, seg_all).numpy()
    u, v = np.empty(allq.shape[0]), np.empty(allq.shape[0])
    u[by_leaf], v[by_leaf] = back[:, 0], back[:, 1]
    # combine in the reference's order (mixtureGP.jl:263-272)
    Y, V = np.empty(nloc), np.empty(nloc)
    counts = np.bincount(allq, minlength=nloc)
    off = np.concatenate([[0], np.cumsum(counts)])
    for j in range(nloc):
        sw = 0.0
        for s in range(off[j], off[j + 1]):
            sw = sw + allw[s]
        y = vv = 0.0
        for s in range(off[j], off[j + 1]):
            w = allw[s] / sw
            y = y + w * u[s]
            vv = vv + w * (v[s] * w)
        Y[j], V[j] = y, vv
    np.save(os.path.join(out_dir, f"Y{rank}.npy"), Y)
    np.save(os.path.join(out_dir, f"V{rank}.npy"), V)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nq", [400, 401])
def test_two_rank_ownership_flow_matches_single_process(tmp_path, nq):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    import helpers
    from oracle import pmk_oracle as O
    port = 29500 + (os.getpid() % 2000) + nq % 7
    mp.spawn(_worker, args=(2, port, nq, str(tmp_path)), nprocs=2, join=True)
    case = cases.mixgp_file()
    m = helpers.oracle_model(case)
    th, _ = helpers.kernels(case["kernel"])
    wth, _ = helpers.kernels(case["wkernel"])
    Xq = case["Xq"][::7][:nq]
    Y, V, _ = O.querymixtureGP(Xq, m["eta"], m["root"], case["levels"], case["radius"], case["delta"], th, case["sigma2"], wth)
    Yg = np.concatenate([np.load(tmp_path / "Y0.npy"), np.load(tmp_path / "Y1.npy")])
    Vg = np.concatenate([np.load(tmp_path / "V0.npy"), np.load(tmp_path / "V1.npy")])
    # a pair's u, v depend on its leaf and point only, the combine order is the reference's: identical bits
    assert np.array_equal(Yg, Y) and np.array_equal(Vg, V)
