"""BASELINE.json configs[3] (C4: 3-D, N = 4M, 8192 leaves of ~1024 points) and configs[4] (C5: dev/image_upscale.jl-style
kriging of a 2048^2 image onto the 4096^2 grid, leaves of up to 2048 points) at FULL size on one B200, in the style of
tests/test_gpu_fullsize.py: bit-exact indexing (device == host mirror), the oracle on the leaves a few dozen queries touch
(structure bit-exact, u and v within 1e-9), and size-independent properties (permutation invariance, variance bounds).
One configuration is resident at a time (C5's factors + query operand are 82 GB)."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

import helpers
import patchmixturekriging_b200 as P
from patchmixturekriging_b200 import _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _build(name):
    import bench
    w = bench.workload(name)
    X, y = w["X"], w["y"]
    root, X_parts, X_parts_inds = P.setuppartition(X, w["levels"])
    X_set, X_set_inds, _, _ = P.organizetrainingsets_device(root, w["levels"], X, w["eps"])
    θ = w["theta"]()
    wθ = P.Spline34KernelType(1.0 / w["radius"])
    η = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
    P.fitmixtureGP_(η, [y[i - 1] for i in X_set_inds], θ, w["sigma2"])
    return dict(w=w, root=root, X_set=X_set, X_set_inds=X_set_inds, X_parts_inds=X_parts_inds, θ=θ, wθ=wθ, η=η, bench=bench)


@pytest.fixture(scope="module")
def c4(built_lib):
    m = _build("c4")
    yield m
    m["η"].close()


@pytest.fixture(scope="module")
def c5(built_lib, c4):          # after c4 (its fixture is torn down at module end; close it early to make room)
    c4["η"].close()
    m = _build("c5")
    yield m
    m["η"].close()


def _indexing(m, n_leaves, D):
    w, root = m["w"], m["root"]
    assert len(m["X_set"]) == n_leaves and len(root.hps_c) == n_leaves - 1
    # device eps-overlap sets == host mirror (itself bit-exact against the oracle at small sizes), every leaf, every index
    _, inds_h, _, _ = P.organizetrainingsets(root, w["levels"], w["X"], w["eps"])
    assert all(np.array_equal(a, b) for a, b in zip(m["X_set_inds"], inds_h))
    # level-stepped device tree build == host mirror: hyperplanes and base partition
    droot, _, dinds = P.setuppartition_device(w["X"], w["levels"])
    assert np.array_equal(droot.hps_v, root.hps_v) and np.array_equal(droot.hps_c, root.hps_c)
    assert all(np.array_equal(a, b) for a, b in zip(dinds, m["X_parts_inds"]))
    for leaf in (0, n_leaves // 3, n_leaves - 1):
        assert np.isin(m["X_parts_inds"][leaf], m["X_set_inds"][leaf]).all()
    # device findpartition == host findpartition
    Xq = m["bench"].gen_queries(w, 1_000_000, 1_000_000)
    out = np.empty(len(Xq), dtype=np.int32)
    h = m["η"].handle
    hv, hc = np.ascontiguousarray(root.hps_v), np.ascontiguousarray(root.hps_c)
    h.check(_lib.lib().pmk_set_tree(h.raw, D, w["levels"], _lib.ptr(hv), _lib.ptr(hc)))
    h.check(_lib.lib().pmk_find_partition(h.raw, len(Xq), _lib.ptr(Xq), _lib.ptr(out)))
    assert np.array_equal(out, P.findpartition(Xq, root))


def _oracle_spot_check(m, kernel, first, count, tag):
    """the oracle on the leaves the queries touch: home leaf, neighbour leaves, hyperplane ids, t bit-exact; u, v to 1e-9"""
    from oracle import pmk_oracle as O
    w, root, η, θ, wθ = m["w"], m["root"], m["η"], m["θ"], m["wθ"]
    Xq = m["bench"].gen_queries(w, first, count)
    Y, V, dv = P.querymixtureGP(Xq, η, root, w["levels"], w["radius"], w["delta"], θ, w["sigma2"], wθ, debug_flag=True)
    f = dv._flat
    hv, hc = np.ascontiguousarray(root.hps_v), np.ascontiguousarray(root.hps_c)
    home, pq, ph, pl, pt = O.query_structure_vec(Xq, hv, hc, w["levels"], w["radius"], w["delta"])
    assert np.array_equal(f["home"], home)
    nb = f["pair_hp"] != 0
    assert np.array_equal(f["pair_leaf"][nb], pl) and np.array_equal(f["pair_hp"][nb], ph + 1) and np.array_equal(f["pair_t"][nb], pt)
    assert np.array_equal(np.diff(f["pair_off"]), np.bincount(pq, minlength=len(Xq)) + 1)
    oth = O.Kernel(*kernel)
    cache = {}
    eu = ev = 0.0
    us, vs, uo, vo = [], [], [], []
    for j in range(len(Xq)):
        for s in range(f["pair_off"][j], f["pair_off"][j + 1]):
            leaf = int(f["pair_leaf"][s])
            if leaf not in cache:
                Xl = m["X_set"][leaf - 1]
                U = O.constructkernelmatrix(Xl, oth) + w["sigma2"] * np.eye(len(Xl))
                cache[leaf] = (Xl, O.backslash(U, w["y"][m["X_set_inds"][leaf - 1] - 1]), O.cholesky_L(U))
            Xl, c, L = cache[leaf]
            u, v = O.queryinner(Xq[j], Xl, oth, c, L)
            us.append(f["pair_u"][s]); vs.append(f["pair_v"][s]); uo.append(u); vo.append(v)
    us, vs, uo, vo = map(np.asarray, (us, vs, uo, vo))
    su, sv = np.sqrt(np.mean(uo ** 2)), np.sqrt(np.mean(vo ** 2))
    eu = float((np.abs(us - uo) / np.maximum(np.abs(uo), su)).max())
    ev = float((np.abs(vs - vo) / np.maximum(np.abs(vo), sv)).max())
    helpers.record_parity(f"fullsize/{tag}", pairs=int(len(us)), leaves=len(cache), u_floored_rel=eu, v_floored_rel=ev,
                          v_pointwise_rel=float((np.abs(vs - vo) / vo).max()), tol=1e-9)
    assert eu <= 1e-9 and ev <= 1e-9, (eu, ev)


def _properties(m, first, count):
    w, root, η, θ, wθ = m["w"], m["root"], m["η"], m["θ"], m["wθ"]
    Xq = m["bench"].gen_queries(w, first, count)
    args = (root, w["levels"], w["radius"], w["delta"], θ, w["sigma2"], wθ)
    Y, V, dv = P.querymixtureGP(Xq, η, *args, debug_flag=True)
    f = dv._flat
    assert np.isfinite(Y).all() and np.all(V >= 1e-12) and np.all(V <= 1.0 + 1e-12)
    assert np.array_equal(f["pair_leaf"][f["pair_off"][1:] - 1], f["home"])          # home slot last
    perm = np.random.default_rng(1).permutation(len(Xq))
    Yp, Vp, _ = P.querymixtureGP(Xq[perm], η, *args)
    assert np.array_equal(Yp, Y[perm]) and np.array_equal(Vp, V[perm])               # a pair never depends on its tile mates
    return Xq, Y, V, np.diff(f["pair_off"])


# ---- C4 -------------------------------------------------------------------------------------------------------------------
def test_c4_indexing_bit_exact(c4):
    sizes = np.array([len(i) for i in c4["X_set_inds"]])
    assert 600 < sizes.min() and sizes.max() < 1700 and abs(sizes.mean() - 1026) < 30        # "8192 leaves of ~1024 pts"
    _indexing(c4, 8192, 3)


def test_c4_oracle_spot_check(c4):
    _oracle_spot_check(c4, (0, c4["w"]["eps_sq"]), 9_000_000, 32, "c4")


def test_c4_query_properties(c4):
    from patchmixturekriging_b200 import synth
    Xq, Y, V, npairs = _properties(c4, 3_000_000, 300_000)
    assert npairs.min() >= 1 and 1.5 < npairs.mean() < 4.0
    assert np.median(np.abs(Y - synth.f_mixgp(Xq))) < 5e-3                           # the fit tracks the generating function


def test_c4_two_virtual_ranks_equal_one(c4):
    """C4 is the configuration the north star shards: sub-tree ownership (two ranks on this GPU) == the single handle."""
    w, root, θ, wθ = c4["w"], c4["root"], c4["θ"], c4["wθ"]
    args = (root, w["levels"], w["radius"], w["delta"], θ, w["sigma2"], wθ)
    Xq = c4["bench"].gen_queries(w, 5_000_000, 200_000)
    Y0, V0, _ = P.querymixtureGP(Xq, c4["η"], *args)
    c4["η"].close()                      # make room: the sharded twin holds the same 78 GB
    em = P.MixtureGPType(c4["X_set"], P.fetchhyperplanes(root), devices=[0, 0])
    P.fitmixtureGP_(em, [w["y"][i - 1] for i in c4["X_set_inds"]], θ, w["sigma2"])
    Y1, V1, _ = P.querymixtureGP(Xq, em, *args)
    em.close()
    assert np.array_equal(Y0, Y1) and np.array_equal(V0, V1)


# ---- C5 -------------------------------------------------------------------------------------------------------------------
def test_c5_indexing_bit_exact(c5):
    sizes = np.array([len(i) for i in c5["X_set_inds"]])
    assert sizes.max() <= 2048 and sizes.max() > 1536 and sizes.min() > 1000         # the two largest size classes
    _indexing(c5, 4096, 2)


def test_c5_oracle_spot_check(c5):
    w = c5["w"]
    _oracle_spot_check(c5, w["kernel_oracle"], 4096 * 1777 + 1500, 24, "c5")


def test_c5_query_properties(c5):
    w = c5["w"]
    Xq, Y, V, npairs = _properties(c5, 4096 * 2000, 400_000)
    truth = np.sin(7.0 * Xq[:, 0]) * np.cos(5.0 * Xq[:, 1]) + 0.5 * np.exp(-8.0 * ((Xq[:, 0] - 0.6) ** 2 + (Xq[:, 1] - 0.3) ** 2))
    err = np.abs(Y - truth)                                                          # upscaling a smooth image
    assert np.median(err) < 1e-3 and err.max() < 5e-2
    # every second grid line of the 2x grid is (up to rounding of linspace) a training pixel: small variance there
    cond, solver = __import__("patchmixturekriging_b200.mixturegp", fromlist=["x"]).condition_estimate(c5["η"])
    helpers.record_parity("fullsize/c5", cond_lower_bound=cond, solver_in_use=int(solver), pairs_per_query=float(npairs.mean()))
