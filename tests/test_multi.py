"""Multi-rank host logic on CPU (gloo, world_size 2): leaf -> rank map, query slicing, span exchange and the final
gather.  The per-rank compute is stood in for by the oracle (this is test infrastructure; the GPU version of the same
flow is tests/test_gpu_parity.py::test_sharded_fit_equals_single_fit and bench.py --gpus N)."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

from patchmixturekriging_b200 import sharding

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
    Xq = case["Xq"][:nq]
    th, _ = helpers.kernels(case["kernel"])
    wth, _ = helpers.kernels(case["wkernel"])
    # every rank lays the model out identically; it "factorises" only its leaves, peers' spans start as NaN
    m = helpers.oracle_model(case)
    eta = m["eta"]
    n_leaves = len(eta.X_parts)
    a, c = sharding.leaf_range(rank, world, n_leaves)
    sizes = [len(x) for x in eta.X_parts]
    off = np.concatenate([[0], np.cumsum(sizes)])
    off2 = np.concatenate([[0], np.cumsum(np.square(sizes))])
    alpha = torch.full((int(off[-1]),), float("nan"), dtype=torch.float64)
    Lbuf = torch.full((int(off2[-1]),), float("nan"), dtype=torch.float64)
    for p in range(a, a + c):
        alpha[off[p]:off[p + 1]] = torch.from_numpy(eta.c_set[p])
        Lbuf[off2[p]:off2[p + 1]] = torch.from_numpy(np.ascontiguousarray(eta.L_set[p]).ravel())

    def span_of(which, first, count):
        o, buf = (off, alpha) if which == "alpha" else (off2, Lbuf)
        return buf[int(o[first]):int(o[first + count])]

    sharding.exchange_spans(span_of, n_leaves, ["alpha", "L"])
    assert not torch.isnan(alpha).any() and not torch.isnan(Lbuf).any()
    for p in range(n_leaves):          # the replicated model equals the single-process one, bit for bit
        assert np.array_equal(alpha[off[p]:off[p + 1]].numpy(), eta.c_set[p])
        eta.c_set[p] = alpha[off[p]:off[p + 1]].numpy().copy()
        eta.L_set[p] = Lbuf[off2[p]:off2[p + 1]].numpy().reshape(sizes[p], sizes[p]).copy()
    q0, q1 = sharding.query_slice(rank, world, nq)
    Y, V, _ = O.querymixtureGP_vec(Xq[q0:q1], eta, case["levels"], case["radius"], case["delta"], th, wth)
    Yall = sharding.gather_slices(torch.from_numpy(Y), nq)
    Vall = sharding.gather_slices(torch.from_numpy(V), nq)
    if rank == 0:
        np.save(os.path.join(out_dir, "Y.npy"), Yall.numpy())
        np.save(os.path.join(out_dir, "V.npy"), Vall.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nq", [600, 601])
def test_two_rank_flow_matches_single_process(tmp_path, nq):
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
    Y, V, _ = O.querymixtureGP_vec(case["Xq"][:nq], m["eta"], case["levels"], case["radius"], case["delta"], th, wth)
    Yg, Vg = np.load(tmp_path / "Y.npy"), np.load(tmp_path / "V.npy")
    # slices are answered independently, so the batched BLAS calls see different shapes: equal to rounding
    assert np.abs(Yg - Y).max() <= 1e-10 * np.abs(Y).max() and np.abs(Vg - V).max() <= 1e-8 * np.abs(V).max()
