"""CPU tests of the oracle: the known answers and invariants that pin it, and agreement of
its vectorised routines with its scalar (reference-shaped) loops."""
from __future__ import annotations

import numpy as np
import pytest

import cases
import helpers
from oracle import pmk_oracle as O


def test_kernel_known_answers():
    assert abs(O.evalkernel([0.3], [0.7], O.Kernel(O.BB10)) - 0.09) < 1e-16             # min(p,q) - p*q
    assert O.evalkernel([0.0, 0.0], [0.0, 0.0], O.Kernel(O.SPLINE34, 2.0)) == 1.0       # Spline34(0) = 1
    assert O.evalkernel([0.0, 0.0], [0.5, 0.0], O.Kernel(O.SPLINE34, 2.0)) == 0.0       # r = 1 -> 0
    assert O.evalkernel([0.0, 0.0], [3.0, 4.0], O.Kernel(O.SPLINE34, 1.0)) == 0.0       # r >= 1 -> 0
    assert O.evalkernel([1.0, 2.0], [1.5, 2.0], O.Kernel(O.SQEXP, 3.0)) == np.exp(-3.0 * 0.25)
    x, z = 0.3, 0.6                                                                      # BB20, x < z branch
    assert O.evalkernel([x], [z], O.Kernel(O.BB20)) == (((-1 / 6) * x) * (1 - z)) * ((x * x + z * z) - 2 * z)
    assert abs(O.evalkernel([x], [z], O.Kernel(O.BB20)) - O.evalkernel([z], [x], O.Kernel(O.BB20))) < 1e-17
    # BB1eps -> BB10 as eps -> 0
    assert abs(O.evalkernel([x], [z], O.Kernel(O.BB1EPS, 1e-4)) - O.evalkernel([x], [z], O.Kernel(O.BB10))) < 1e-9
    # tensor product over dimensions (kernel.jl:196-198)
    a = O.evalkernel([0.2, 0.9], [0.5, 0.4], O.Kernel(O.BB10))
    assert a == O.evalkernel([0.2], [0.5], O.Kernel(O.BB10)) * O.evalkernel([0.9], [0.4], O.Kernel(O.BB10))


def test_gram_is_mirrored_and_posdef():
    case = cases.ibb1d(15, 100, "BB10")
    K = O.constructkernelmatrix(case["X"], O.Kernel(O.BB10))
    assert np.array_equal(K, K.T)
    assert np.linalg.matrix_rank(K) == 15 and np.all(np.linalg.eigvalsh(K) > 0)          # examples/IBB1D.jl:38-41


def test_traversal_orders():
    """dev/btree_easy.jl:63-69: PreOrder [0,1,3,2], Leaves [3,2] on the same node type."""
    n0, n1, n2, n3 = O.Node(), O.Node(), O.Node(), O.Node()
    n0.index, n1.index, n2.index, n3.index = 0, 1, 2, 3
    n0.left, n0.right, n1.left = n1, n2, n3
    assert [n.index for n in O.preorder(n0)] == [0, 1, 3, 2]
    assert [n.index for n in O.leaves(n0)] == [3, 2]


def test_partition_invariants():
    case = cases.mixgp_driver(N=5000, levels=5)
    X = case["X"]
    root, X_parts, X_parts_inds = O.setuppartition(X, 5)
    hv, hc = O.fetchhyperplanes(root)
    assert len(hc) == len(X_parts) - 1                              # examples/patchGP_partitioning.jl:198
    assert len(X_parts) == 16
    allinds = np.sort(np.concatenate(X_parts_inds))
    assert np.array_equal(allinds, np.arange(1, 5001))              # a partition of 1..N
    sizes = [len(i) for i in X_parts_inds]
    assert max(sizes) - min(sizes) <= 1                             # median splits
    for inds, Xp in zip(X_parts_inds, X_parts):
        assert np.all(np.diff(inds) > 0)                            # global order preserved
        assert np.array_equal(X[inds - 1], Xp)                      # partition.jl:151-153
        assert len({O.findpartition(x, root, 5) for x in Xp[:20]}) == 1
    assert np.allclose(np.linalg.norm(hv, axis=1), 1.0, atol=1e-15)
    # leaf id = 1 + path bits; vectorised descent agrees with the scalar one
    lv = O._descend_vec(X[:500], hv, hc, 5)
    assert [O.findpartition(x, root, 5) for x in X[:500]] == list(lv)
    for leaf, inds in enumerate(X_parts_inds, start=1):
        assert np.all(O._descend_vec(X[inds - 1], hv, hc, 5) == leaf)
    # |t| is the distance to the projection (examples/patchGP_partitioning.jl:214-215)
    p = X[17]
    reg, ts, zs, keep = O.findneighbourpartitions(p, 0.5, root, 5, hv, hc, O.findpartition(p, root, 5), 1e-5)
    assert np.linalg.norm(np.linalg.norm(zs - p, axis=1) - np.abs(ts)) < 1e-10


def test_split_direction_forms():
    z = np.array([0.3, -1.2, 0.5])
    vc = O.split_direction(z, "column")
    vr = O.split_direction(z, "row")
    assert np.allclose(vc, z / np.linalg.norm(z), atol=1e-15)       # +z/|z|
    assert np.allclose(vr, -np.sign(z[0]) * z / np.linalg.norm(z), atol=1e-15)


def test_mean_and_median_restatements():
    rng = np.random.default_rng(0)
    for n in (1, 2, 15, 1024, 1025, 3000, 5001):
        X = rng.normal(size=(n, 2))
        m = O.mean_pairwise(X)
        assert np.allclose(m, X.mean(0), rtol=1e-13, atol=1e-15)
        f = rng.normal(size=n)
        assert abs(O.median_julia(f) - np.median(f)) < 1e-15
    f = np.array([4.0, 1.0, 3.0, 2.0])
    assert O.median_julia(f) == 2.0 / 2 + 3.0 / 2


def test_organize_scalar_vs_vectorised():
    case = cases.mixgp_file()
    root, _, _ = O.setuppartition(case["X"], 3)
    hv, hc = O.fetchhyperplanes(root)
    X_set, X_set_inds, rl, prob = O.organizetrainingsets(root, 3, case["X"], 1.5)
    v = O.organizetrainingsets_vec(hv, hc, 3, case["X"], 1.5)
    assert prob == [] and all(np.array_equal(a, b) for a, b in zip(X_set_inds, v))
    home = O._descend_vec(case["X"], hv, hc, 3)
    assert all(h in r for h, r in zip(home, rl))                    # partition.jl:325-326 sanity check
    assert all(np.array_equal(case["X"][i - 1], Xs) for i, Xs in zip(X_set_inds, X_set))


def test_query_scalar_vs_vectorised_and_debug_lengths():
    case = cases.mixgp_file()
    m = helpers.oracle_model(case)
    wth, _ = helpers.kernels(case["wkernel"])
    Xq = case["Xq"][::41]
    Y, V, dbg = O.querymixtureGP(Xq, m["eta"], m["root"], 3, 0.3, 1e-5, m["th"], 1e-5, wth, debug=True)
    Y2, V2, d2 = O.querymixtureGP_vec(Xq, m["eta"], 3, 0.3, 1e-5, m["th"], wth)
    for u, w, r in zip(dbg["u"], dbg["w_tilde"], dbg["region_inds"]):
        assert len(u) == len(w) == len(r) + 1                       # examples/helpers/visualization.jl:163
    assert np.array_equal(np.array(dbg["p_region_ind"]), d2["home"])
    assert np.abs(Y - Y2).max() <= 1e-9 * np.abs(Y).max()
    assert np.abs(V - V2).max() <= 2e-8 * np.abs(V).max()           # dtrsv vs dtrsm at cond ~1e6
    assert np.all(V >= 1e-12)


def test_fit_matches_definition():
    case = cases.mixgp_file()
    m = helpers.oracle_model(case)
    eta = m["eta"]
    for X, c, L, inds in zip(eta.X_parts, eta.c_set, eta.L_set, m["X_set_inds"]):
        U = O.constructkernelmatrix(X, m["th"]) + 1e-5 * np.eye(len(X))
        assert np.abs(L @ L.T - U).max() < 1e-13
        assert np.abs(U @ c - case["y"][inds - 1]).max() < 1e-9
    with pytest.raises(O.PosDefException):
        O.cholesky_L(np.array([[1.0, 2.0], [2.0, 1.0]]))


def test_single_gp_ibb1d():
    case = cases.ibb1d(15, 100, "BB10")
    th = O.Kernel(O.BB10)
    c = O.fitRKHS(case["X"], case["y"], th, 1e-5)
    yq = O.query_rkhs(case["X"], case["X"], c, th)
    assert np.abs(yq - case["y"]).max() < 1e-3                      # near-interpolation at sigma2 = 1e-5


def test_c_restatement_matches_numpy_oracle():
    """oracle/pmk_oracle.c (the timed CPU baseline) against oracle/pmk_oracle.py on the mixGP.jl workload."""
    from oracle import c_oracle
    case = cases.mixgp_file()
    m = helpers.oracle_model(case)
    eta = m["eta"]
    wth, _ = helpers.kernels(case["wkernel"])
    y_set = [case["y"][i - 1] for i in m["X_set_inds"]]
    alpha, L, off, Xp = c_oracle.fit(eta.X_parts, y_set, m["th"].kind, m["th"].param, case["sigma2"])
    loff = np.concatenate([[0], np.cumsum(np.diff(off) ** 2)])
    for p in range(len(eta.X_parts)):
        n = off[p + 1] - off[p]
        Lc = L[loff[p]:loff[p + 1]].reshape(n, n, order="F")
        assert np.abs(Lc - eta.L_set[p]).max() < 1e-10 * np.abs(eta.L_set[p]).max()
        K = c_oracle.gram(eta.X_parts[p], m["th"].kind, m["th"].param)
        np.testing.assert_allclose(K, O.constructkernelmatrix(eta.X_parts[p], m["th"]), rtol=2e-15, atol=0)   # pow(t,6): libm vs numpy, 1-2 ulp
    Xq = case["Xq"][::7]
    Yc, Vc, home, npairs, absent = c_oracle.query(m["hv"], m["hc"], 3, off, Xp, alpha, L, m["th"].kind, m["th"].param, Xq, 0.3, 1e-5,
                                          wth.kind, wth.param)
    Yo, Vo, od = O.querymixtureGP_vec(Xq, eta, 3, 0.3, 1e-5, m["th"], wth)
    assert not absent.any()
    assert np.array_equal(home, od["home"])
    assert np.array_equal(npairs, np.diff(od["pair_off"]))
    assert np.abs(Yc - Yo).max() <= 2e-9 * np.abs(Yo).max()
    assert np.abs(Vc - Vo).max() <= 2e-8 * np.abs(Vo).max()
