"""BASELINE.json's full-size configuration (C3: 2-D, N = 1M, 4096 leaves of ~512 points, SqExp) on the GPU, checked through
size-independent properties (the oracle cannot run this size in seconds): bit-exact indexing against the host mirror,
factor residuals on sampled leaves, linearity of the mean in y, y-independence of the variance, invariance under query
permutation, pruned == full hyperplane scan, and an oracle spot check on the leaves around a handful of queries."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np
import pytest

import patchmixturekriging_b200 as P
from patchmixturekriging_b200 import _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def c3(built_lib):
    import bench
    w = bench.workload("c3")
    X, y = w["X"], w["y"]
    root, X_parts, X_parts_inds = P.setuppartition(X, w["levels"])
    X_set, X_set_inds, rl, _ = P.organizetrainingsets_device(root, w["levels"], X, w["eps"])
    θ = P.GaussianKernel1DType(w["eps_sq"])
    wθ = P.Spline34KernelType(1.0 / w["radius"])
    η = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
    P.fitmixtureGP_(η, [y[i - 1] for i in X_set_inds], θ, w["sigma2"])
    return dict(w=w, root=root, X_set=X_set, X_set_inds=X_set_inds, X_parts_inds=X_parts_inds, θ=θ, wθ=wθ, η=η, bench=bench)


def test_c3_indexing_bit_exact(c3):
    w, root = c3["w"], c3["root"]
    assert len(c3["X_set"]) == 4096 and len(root.hps_c) == 4095
    sizes = np.array([len(i) for i in c3["X_set_inds"]])
    assert 300 < sizes.min() and sizes.max() < 800 and abs(sizes.mean() - 512) < 16      # "4096 leaves of ~512 pts with overlap"
    # device ε-overlap sets == host mirror (itself bit-exact against the oracle at small sizes)
    _, inds_h, _, _ = P.organizetrainingsets(root, w["levels"], w["X"], w["eps"])
    assert all(np.array_equal(a, b) for a, b in zip(c3["X_set_inds"], inds_h))
    # every base-partition point is in its own leaf's overlapped set
    for leaf in (0, 1000, 4095):
        assert np.isin(c3["X_parts_inds"][leaf], c3["X_set_inds"][leaf]).all()
    # device findpartition == host findpartition on 2M query points
    Xq = c3["bench"].gen_queries(w, 0, 2_000_000)
    out = np.empty(len(Xq), dtype=np.int32)
    h = c3["η"].handle
    hv, hc = np.ascontiguousarray(root.hps_v), np.ascontiguousarray(root.hps_c)
    h.check(_lib.lib().pmk_set_tree(h.raw, 2, w["levels"], _lib.ptr(hv), _lib.ptr(hc)))
    h.check(_lib.lib().pmk_find_partition(h.raw, len(Xq), _lib.ptr(Xq), _lib.ptr(out)))
    assert np.array_equal(out, P.findpartition(Xq, root))
    assert np.bincount(out, minlength=4097)[1:].min() > 0


def test_c3_factor_residuals(c3):
    w, η = c3["w"], c3["η"]
    for leaf in (0, 777, 2048, 4095):
        K = η.U_set[leaf]
        L = η.L_set[leaf]
        c = η.c_set[leaf]
        U = K + w["sigma2"] * np.eye(len(K))
        y = w["y"][c3["X_set_inds"][leaf] - 1]
        assert np.abs(L @ L.T - U).max() < 5e-14
        assert np.abs(U @ c - y).max() < 1e-10 * max(1.0, np.abs(c).max())
        assert np.array_equal(np.triu(L, 1), np.zeros_like(L)) and np.all(np.diag(L) > 0)


def test_c3_query_properties(c3):
    w, root, η, θ, wθ = c3["w"], c3["root"], c3["η"], c3["θ"], c3["wθ"]
    Xq = c3["bench"].gen_queries(w, 3_000_000, 500_000)
    args = (root, w["levels"], w["radius"], w["delta"], θ, w["sigma2"], wθ)
    Y, V, dv = P.querymixtureGP(Xq, η, *args, debug_flag=True)
    f = dv._flat
    assert np.isfinite(Y).all() and np.all(V >= 1e-12) and np.all(V <= 1.0 + 1e-12)
    assert np.all((f["pair_w"] >= 0) & (f["pair_w"] <= 1.0))
    npairs = np.diff(f["pair_off"])
    assert npairs.min() >= 1 and 1.5 < npairs.mean() < 2.3                 # SURVEY: ~1.9 leaves per query at radius = eps
    assert np.array_equal(f["pair_leaf"][f["pair_off"][1:] - 1], f["home"])  # home slot last
    # the fit is good where there is data: predictions track the generating function
    from patchmixturekriging_b200 import synth
    err = np.abs(Y - synth.f_mixgp(Xq))
    assert np.median(err) < 1e-3
    # permutation invariance, bit for bit (a pair's result must not depend on its tile mates)
    perm = np.random.default_rng(0).permutation(len(Xq))
    Yp, Vp, _ = P.querymixtureGP(Xq[perm], η, *args)
    assert np.array_equal(Yp, Y[perm]) and np.array_equal(Vp, V[perm])
    # pruned neighbour search == scan over all 4095 hyperplanes (on a 60k slice)
    L = _lib.lib()
    sl = slice(0, 60_000)
    η.handle.check(L.pmk_set_option(η.handle.raw, _lib.OPT_FULL_HYPERPLANE_SCAN, 1))
    Yf, Vf, dvf = P.querymixtureGP(Xq[sl], η, *args, debug_flag=True)
    η.handle.check(L.pmk_set_option(η.handle.raw, _lib.OPT_FULL_HYPERPLANE_SCAN, 0))
    Ys, Vs, dvs = P.querymixtureGP(Xq[sl], η, *args, debug_flag=True)
    for key in ("home", "pair_off", "pair_leaf", "pair_hp", "pair_t", "pair_w"):
        assert np.array_equal(dvf._flat[key], dvs._flat[key]), key
    assert np.array_equal(Yf, Ys) and np.array_equal(Vf, Vs) and np.array_equal(Ys, Y[sl])


def test_c3_linearity_and_variance_independent_of_y(c3):
    """posterior mean is linear in y; posterior variance does not depend on y at all."""
    w, root, θ, wθ = c3["w"], c3["root"], c3["θ"], c3["wθ"]
    X = w["X"]
    y1 = w["y"]
    y2 = np.cos(0.7 * X[:, 0]) * np.sin(0.3 * X[:, 1])
    Xq = c3["bench"].gen_queries(w, 7_000_000, 200_000)
    args = (root, w["levels"], w["radius"], w["delta"], θ, w["sigma2"], wθ)
    res = []
    for y in (y1, y2, 2.0 * y1 - 3.0 * y2):
        η = P.MixtureGPType(c3["X_set"], P.fetchhyperplanes(root))
        P.fitmixtureGP_(η, [y[i - 1] for i in c3["X_set_inds"]], θ, w["sigma2"])
        res.append(P.querymixtureGP(Xq, η, *args)[:2])
        η.close()
    (Y1, V1), (Y2, V2), (Y3, V3) = res
    scale = np.sqrt(np.mean(Y3 ** 2))
    assert np.abs(Y3 - (2.0 * Y1 - 3.0 * Y2)).max() < 1e-9 * scale
    assert np.array_equal(V1, V2) and np.array_equal(V1, V3)


def test_c3_oracle_spot_check(c3):
    """the oracle on the handful of leaves a few queries touch, at full C3 size."""
    from oracle import pmk_oracle as O
    w, root, η, θ, wθ = c3["w"], c3["root"], c3["η"], c3["θ"], c3["wθ"]
    Xq = c3["bench"].gen_queries(w, 9_000_000, 40)
    Y, V, dv = P.querymixtureGP(Xq, η, root, w["levels"], w["radius"], w["delta"], θ, w["sigma2"], wθ, debug_flag=True)
    oth = O.Kernel(O.SQEXP, w["eps_sq"])
    f = dv._flat
    cache = {}
    for j in range(len(Xq)):
        us, vs = [], []
        for s in range(f["pair_off"][j], f["pair_off"][j + 1]):
            leaf = int(f["pair_leaf"][s])
            if leaf not in cache:
                Xl = c3["X_set"][leaf - 1]
                U = O.constructkernelmatrix(Xl, oth) + w["sigma2"] * np.eye(len(Xl))
                cache[leaf] = (Xl, O.backslash(U, w["y"][c3["X_set_inds"][leaf - 1] - 1]), O.cholesky_L(U))
            Xl, c, L = cache[leaf]
            u, v = O.queryinner(Xq[j], Xl, oth, c, L)
            assert abs(f["pair_u"][s] - u) <= 1e-9 * max(abs(u), 1e-3)
            assert abs(f["pair_v"][s] - v) <= 1e-9 * v
