"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the
same seeded inputs.  Bit-exact for leaf assignment / point indexing / neighbour lists; floating point
within the tolerances written next to each assert (north star: 1e-9 relative on mean and variance)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

import cases
import helpers
import patchmixturekriging_b200 as P
from oracle import pmk_oracle as O
from patchmixturekriging_b200 import _lib

pytestmark = pytest.mark.gpu

# tolerance on posterior mean / variance: |gpu - oracle| <= TOL * max(|oracle|, rms(oracle))
TOL = 1e-9


def assert_close(name, a, b, tol=TOL, record=None):
    st = helpers.bound_stats(a, b)
    print(f"{name}: {st}")
    if record:
        helpers.record_parity(record, **{name: dict(st, tol=tol)})
    scale = np.sqrt(np.mean(np.asarray(b) ** 2))
    bound = tol * np.maximum(np.abs(b), scale)
    bad = np.abs(np.asarray(a) - np.asarray(b)) > bound
    assert not bad.any(), f"{name}: {int(bad.sum())} of {bad.size} exceed {tol}: {st}"


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kname,param", [("SQEXP", 3.0), ("SPLINE34", 0.4), ("SPLINE12", 0.4), ("SPLINE32", 0.4), ("RQ", 1.7)])
@pytest.mark.parametrize("D,n", [(1, 77), (2, 130), (3, 257), (2, 1)])
def test_gram_stationary(built_lib, kname, param, D, n):
    X = P.synth.uniform_points(3 + D + n, n, [-1.0] * D, [1.0] * D) if hasattr(P, "synth") else None
    from patchmixturekriging_b200 import synth
    X = synth.uniform_points(3 + D + n, n, [-1.0] * D, [1.0] * D)
    ok, pk = helpers.kernels((kname, param))
    K = P.constructkernelmatrix(X, pk)
    Kref = O.constructkernelmatrix(X, ok)
    assert K.shape == (n, n)
    assert np.array_equal(K, K.T)                       # mirrored, exactly symmetric (RKHS.jl:27-31)
    np.testing.assert_allclose(K, Kref, rtol=5e-15, atol=1e-300)   # a few ulp: exp / t^6 implementations differ
    Ks = P.constructkernelmatrix(X, pk, σ2=0.25)
    np.testing.assert_allclose(np.diag(Ks), np.diag(Kref) + 0.25, rtol=5e-15)


@pytest.mark.parametrize("kname,param", [("BB10", 1.0), ("BB20", 1.0), ("BB1EPS", 4.5), ("BB2EPS", 2.5)])
@pytest.mark.parametrize("D,n", [(1, 15), (2, 100)])
def test_gram_brownian_bridge(built_lib, kname, param, D, n):
    from patchmixturekriging_b200 import synth
    X = synth.uniform_points(17 + n, n, [1e-3] * D, [1.0 - 1e-3] * D)
    ok, pk = helpers.kernels((kname, param))
    K = P.constructkernelmatrix(X, pk)
    Kref = O.constructkernelmatrix(X, ok)
    rtol = 5e-15 if kname in ("BB10", "BB20") else 1e-11    # eps-kernels: cancellation-prone exponent sums (kernel.jl:176-193)
    np.testing.assert_allclose(K, Kref, rtol=rtol, atol=1e-18 if rtol > 1e-14 else 0)


@pytest.mark.parametrize("D,n", [(1, 300), (2, 1000), (3, 777)])
def test_gram_fast_exp_option(built_lib, D, n):
    """PMK_OPT_GRAM_FAST_EXP: exp(-eps_sq |x - z|^2) with the table-driven exp.  The exp itself is within 1.3 ulp; its argument
    (sum of squares instead of the reference's sqrt + re-square) differs by a few ulp (measured: up to 3.3), which the exponential turns into that many |arg| ulp of
    the value: entry by entry |fast - oracle| <= (3 + 6 |log K|) ulp, i.e. < 1e-13 relative wherever K > 1e-30.  Exactly
    symmetric, diagonal exactly 1 (+ sigma2)."""
    from patchmixturekriging_b200 import synth
    X = synth.uniform_points(9 + n, n, [-1.0] * D, [1.0] * D)
    ok, pk = helpers.kernels(("SQEXP", 37.0))
    Kref = O.constructkernelmatrix(X, ok)
    Kf = P.constructkernelmatrix(X, pk, fast_exp=True)
    Ke = P.constructkernelmatrix(X, pk)
    assert np.array_equal(Kf, Kf.T) and np.all(np.diag(Kf) == 1.0)
    big = Kref > 1e-300
    ulp = np.abs(Kf - Kref)[big] / np.spacing(Kref[big])
    bound = 3.0 + 6.0 * np.abs(np.log(Kref[big]))
    helpers.record_parity(f"gram_fast_exp/D{D}_n{n}", max_ulp_vs_oracle=float(ulp.max()), max_ulp_over_bound=float((ulp / bound).max()),
                          exact_path_max_rel=float((np.abs(Ke - Kref)[big] / Kref[big]).max()))
    assert np.all(ulp <= bound), float((ulp / bound).max())
    near = Kref > 1e-30
    np.testing.assert_allclose(Kf[near], Kref[near], rtol=1e-13)
    assert np.array_equal(np.diag(P.constructkernelmatrix(X, pk, σ2=0.25, fast_exp=True)), np.full(n, 1.25))


def test_cross_gram_and_evalkernel(built_lib):
    from patchmixturekriging_b200 import synth
    X = synth.uniform_points(1, 70, [-1.0, -1.0], [1.0, 1.0])
    Z = synth.uniform_points(2, 45, [-1.0, -1.0], [1.0, 1.0])
    ok, pk = helpers.kernels(("SQEXP", 2.0))
    K = P.constructkernelmatrix(X, pk, Z)
    np.testing.assert_allclose(K, O.kernel_cross(X, Z, ok), rtol=5e-15)
    assert P.evalkernel(X[0], Z[0], pk) == K[0, 0]
    # known answers: BB10(0.3,0.7) = 0.09, Spline34(0) = 1, Spline34(r>=1) = 0
    assert abs(P.evalkernel([0.3], [0.7], P.BrownianBridge10()) - 0.09) < 1e-16
    assert P.evalkernel([0.0, 0.0], [0.0, 0.0], P.Spline34KernelType(2.0)) == 1.0
    assert P.evalkernel([0.0, 0.0], [0.6, 0.0], P.Spline34KernelType(2.0)) == 0.0


# ---------------------------------------------------------------------------------------------
def _fit_product(case, m):
    _, pk = helpers.kernels(case["kernel"])
    X, y = case["X"], case["y"]
    root, X_parts, X_parts_inds = P.setuppartition(X, case["levels"])
    X_set, X_set_inds, _, _ = P.organizetrainingsets(root, case["levels"], X, case["eps"])
    for a, b in zip(X_set_inds, m["X_set_inds"]):
        assert np.array_equal(a, b)                      # bit-exact point indexing
    eta = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
    P.fitmixtureGP_(eta, [y[i - 1] for i in X_set_inds], pk, case["sigma2"])
    return root, eta, pk


CASES = {"mixgp_file": cases.mixgp_file, "c3_mini": cases.c3_mini, "c4_mini": cases.c4_mini, "mixgp_driver": cases.mixgp_driver,
         "c5_mini": cases.c5_mini, "c3_mini_ill": cases.c3_mini_ill}
_cache = {}


def _setup(name):
    if name not in _cache:
        case = CASES[name]()
        m = helpers.oracle_model(case)
        root, eta, pk = _fit_product(case, m)
        _cache[name] = (case, m, root, eta, pk)
    return _cache[name]


@pytest.mark.parametrize("name", list(CASES))
def test_fit_factor_and_weights(built_lib, name):
    case, m, root, eta, pk = _setup(name)
    oeta = m["eta"]
    worst_L, worst_res = 0.0, 0.0
    for leaf in sorted({0, 1, len(oeta.L_set) // 2, len(oeta.L_set) - 1}):
        L = eta.L_set[leaf]
        Lref = oeta.L_set[leaf]
        assert np.array_equal(np.triu(L, 1), np.zeros_like(L))
        worst_L = max(worst_L, np.abs(L - Lref).max() / np.abs(Lref).max())
        # U_set = Gram without sigma2 (mixtureGP.jl:99)
        K = eta.U_set[leaf]
        np.testing.assert_allclose(K, O.constructkernelmatrix(oeta.X_parts[leaf], m["th"]), rtol=5e-15, atol=1e-300)
        # alpha: residual of (K + s2 I) alpha = y, and agreement with the oracle's LU solution
        c = eta.c_set[leaf]
        U = K + case["sigma2"] * np.eye(K.shape[0])
        y = case["y"][m["X_set_inds"][leaf] - 1]
        worst_res = max(worst_res, np.abs(U @ c - y).max() / (np.abs(U).sum(1).max() * np.abs(c).max()))
    print(f"{name}: L max rel err {worst_L:.3e}, alpha scaled residual {worst_res:.3e}")
    assert worst_L < 1e-9       # blocked DMMA Cholesky vs LAPACK dpotrf, relative to max|L|
    assert worst_res < 1e-14    # backward-stable solve


@pytest.mark.parametrize("builder", [0, 1])
@pytest.mark.parametrize("name", list(CASES))
def test_explicit_inverse_operand(built_lib, name, builder):
    """P = inv(L), the operand of the default (explicit-inverse) query solver: formed on the device by recursive doubling
    on the packed tiles (builder 0, default) or by blocked substitution on identity right-hand sides (builder 1).
    Checked as a residual, P L = I, and for exact triangularity."""
    from patchmixturekriging_b200 import mixturegp
    from patchmixturekriging_b200.mixturegp import _LazyLeafList
    case, m, root, eta, pk = _setup(name)
    mixturegp.set_inverse_builder(eta, builder)
    P_set = _LazyLeafList(eta, "Linv")
    n_leaves = len(eta.X_parts)
    worst = 0.0
    for leaf in sorted({0, 1, n_leaves // 2, n_leaves - 1}):
        L = eta.L_set[leaf]
        Pm = P_set[leaf]
        assert np.array_equal(np.triu(Pm, 1), np.zeros_like(Pm))
        R = Pm @ L - np.eye(L.shape[0])
        worst = max(worst, np.abs(R).max() / (np.abs(Pm) @ np.abs(L)).max())
    mixturegp.set_inverse_builder(eta, 0)
    print(f"{name} builder {builder}: |P L - I| / (|P||L|) = {worst:.3e}")
    assert worst < 1e-14


@pytest.mark.parametrize("name", list(CASES))
def test_query_structure_bit_exact(built_lib, name):
    case, m, root, eta, pk = _setup(name)
    _, wk = helpers.kernels(case["wkernel"])
    Xq = case["Xq"]
    Yq, Vq, dv = P.querymixtureGP(Xq, eta, root, case["levels"], case["radius"], case["delta"], pk, case["sigma2"], wk,
                                  debug_flag=True)
    home, pq, ph, pl, pt = O.query_structure_vec(Xq, m["hv"], m["hc"], case["levels"], case["radius"], case["delta"])
    f = dv._flat
    assert np.array_equal(f["home"], home)                                   # p_region_ind: bit-exact
    counts = np.bincount(pq, minlength=Xq.shape[0]) + 1
    assert np.array_equal(np.diff(f["pair_off"]), counts)
    nb = f["pair_hp"] != 0
    assert np.array_equal(f["pair_leaf"][nb], pl)                            # region_inds, in hyperplane order
    assert np.array_equal(f["pair_hp"][nb], ph + 1)
    assert np.array_equal(f["pair_t"][nb], pt)                               # ts[keep]: same operations, same bits
    assert np.array_equal(f["pair_leaf"][~nb], home)
    wth, _ = helpers.kernels(case["wkernel"])
    np.testing.assert_allclose(f["pair_w"][nb], O.evalkernel_tau(np.abs(pt), wth), rtol=5e-15, atol=1e-300)
    assert np.all(f["pair_w"][~nb] == 1.0)
    # device findpartition entry point
    out = np.empty(Xq.shape[0], dtype=np.int32)
    Xqc = np.ascontiguousarray(Xq)            # bound to a name: ptr() hands out a bare address
    eta.handle.check(_lib.lib().pmk_find_partition(eta.handle.raw, Xqc.shape[0], _lib.ptr(Xqc), _lib.ptr(out)))
    assert np.array_equal(out, home)
    assert np.array_equal(P.findpartition(Xq, root), home)


# Every solver on every configuration is held to TOL = 1e-9, the ill-conditioned ones (sigma2 = 1e-5, cond(K + sigma2 I) 3e6 ..
# 2e7) included.  Measured (profiles/parity_r02.json, floored relative error): well-conditioned cases <= 1.1e-11; mixgp_file
# 2.7e-10 (default = substitution) / 7.1e-10 (explicit inverse); c3_mini at sigma2 = 1e-5 6.5e-11 / 2.2e-10.
SOLVER_NAME = {_lib.SOLVER_AUTO: "auto", _lib.SOLVER_INVERSE: "inverse", _lib.SOLVER_SUBSTITUTION: "substitution"}


@pytest.mark.parametrize("solver", [_lib.SOLVER_AUTO, _lib.SOLVER_INVERSE, _lib.SOLVER_SUBSTITUTION])
@pytest.mark.parametrize("name", list(CASES))
def test_query_mean_variance(built_lib, name, solver):
    """Every query solver against the oracle's dtrsv path: the default (chosen by conditioning), s = inv(L) kq with the
    explicit inverse (row-panel kernel) and blocked forward substitution.  Measured errors go to gpurun_out/parity_r02.json."""
    from patchmixturekriging_b200 import mixturegp
    case, m, root, eta, pk = _setup(name)
    wth, wk = helpers.kernels(case["wkernel"])
    Xq = case["Xq"]
    mixturegp.set_query_solver(eta, solver)
    try:
        Yq, Vq, dv = P.querymixtureGP(Xq, eta, root, case["levels"], case["radius"], case["delta"], pk, case["sigma2"], wk,
                                      debug_flag=True)
        cond, used = mixturegp.condition_estimate(eta)
    finally:
        mixturegp.set_query_solver(eta, _lib.SOLVER_AUTO)
    Yo, Vo, od = O.querymixtureGP_vec(Xq, m["eta"], case["levels"], case["radius"], case["delta"], m["th"], wth)
    f = dv._flat
    assert np.array_equal(f["pair_leaf"], od["pair_leaf"])
    tol = TOL
    rec = f"query/{name}/{SOLVER_NAME[solver]}"
    helpers.record_parity(rec, cond_lower_bound=cond, solver_used=SOLVER_NAME[used], sigma2=case["sigma2"])
    if solver == _lib.SOLVER_AUTO:
        assert used == (_lib.SOLVER_SUBSTITUTION if cond >= 1e4 else _lib.SOLVER_INVERSE)
    assert_close("pair_u", f["pair_u"], od["pair_u"], tol, rec)
    assert_close("pair_v", f["pair_v"], od["pair_v"], tol, rec)
    assert_close("Yq", Yq, Yo, tol, rec)
    assert_close("Vq", Vq, Vo, tol, rec)
    assert np.all(Vq >= 1e-12 * 0.999)


@pytest.mark.parametrize("name", ["mixgp_file", "c3_mini_ill"])
def test_alpha_refinement_effect(built_lib, name):
    """alpha = (K + sigma2 I)^-1 y: the reference solves by LU (mixtureGP.jl:106), the fit by Cholesky (+ one step of iterative
    refinement on models its conditioning estimate flags, PMK_OPT_ALPHA_REFINE).  Both variants against the LU oracle on the
    ill-conditioned configurations; the measured distance is recorded (it is at the level at which LU itself differs from the
    extended-precision solution: profiles/parity_floor_r02.json)."""
    case, m, root, eta, pk = _setup(name)
    wth, wk = helpers.kernels(case["wkernel"])
    Xq = case["Xq"][:6000]
    Yo, Vo, od = O.querymixtureGP_vec(Xq, m["eta"], case["levels"], case["radius"], case["delta"], m["th"], wth)
    X_set, X_set_inds, _, _ = P.organizetrainingsets(root, case["levels"], case["X"], case["eps"])
    y_set = [case["y"][i - 1] for i in X_set_inds]
    for refine in (0, 1):
        e2 = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
        e2.handle.check(_lib.lib().pmk_set_option(e2.handle.raw, _lib.OPT_ALPHA_REFINE, refine))
        P.fitmixtureGP_(e2, y_set, pk, case["sigma2"])
        Yq, Vq, dv = P.querymixtureGP(Xq, e2, root, case["levels"], case["radius"], case["delta"], pk, case["sigma2"], wk, debug_flag=True)
        assert_close("pair_u", dv._flat["pair_u"], od["pair_u"], TOL, f"alpha_refine/{name}/{'on' if refine else 'off'}")
        assert_close("Yq", Yq, Yo, TOL, f"alpha_refine/{name}/{'on' if refine else 'off'}")
        e2.close()


def test_dense_debug_outputs(built_lib):
    """debug_flag=true: hps_keep_flags_set, zs_set, ts_set over ALL hyperplanes (mixtureGP.jl:17-19,256-258) against the scalar
    oracle's findneighbourpartitions -- same operations, same bits."""
    case, m, root, eta, pk = _setup("c3_mini")
    _, wk = helpers.kernels(case["wkernel"])
    Xq = case["Xq"][::501][:40]
    Yq, Vq, dv = P.querymixtureGP(Xq, eta, root, case["levels"], case["radius"], case["delta"], pk, case["sigma2"], wk, debug_flag=True)
    n_hp = m["hc"].shape[0]
    assert len(dv.ts_set) == len(dv.zs_set) == len(dv.hps_keep_flags_set) == Xq.shape[0]
    for j in range(Xq.shape[0]):
        home = O.findpartition(Xq[j], m["root"], case["levels"])
        region, ts, zs, keep = O.findneighbourpartitions(Xq[j], case["radius"], m["root"], case["levels"], m["hv"], m["hc"], home, case["delta"])
        assert dv.ts_set[j].shape == (n_hp,) and dv.zs_set[j].shape == (n_hp, 2)
        assert np.array_equal(dv.ts_set[j], ts)
        assert np.array_equal(dv.zs_set[j], np.asarray(zs))
        assert np.array_equal(dv.hps_keep_flags_set[j], np.asarray(keep, dtype=bool))
        assert np.array_equal(dv.t_kept_set[j], ts[np.asarray(keep, dtype=bool)])
        assert list(dv.region_inds_set[j]) == list(region)


def test_single_query_and_scalar_oracle(built_lib):
    """querymixtureGP(xq::Vector, ...) wrapper (mixtureGP.jl:120-135) and the scalar-loop oracle."""
    case, m, root, eta, pk = _setup("mixgp_file")
    wth, wk = helpers.kernels(case["wkernel"])
    Xq = case["Xq"][::97][:60]
    Yo, Vo, dbg = O.querymixtureGP(Xq, m["eta"], m["root"], case["levels"], case["radius"], case["delta"], m["th"],
                                   case["sigma2"], wth, debug=True)
    Yq, Vq, dv = P.querymixtureGP(Xq, eta, root, case["levels"], case["radius"], case["delta"], pk, case["sigma2"], wk, debug_flag=True)
    assert list(dv.p_region_ind_set) == dbg["p_region_ind"]
    for a, b in zip(dv.region_inds_set, dbg["region_inds"]):
        assert list(a) == list(b)
    assert_close("Yq", Yq, Yo, TOL, "query/mixgp_file/scalar_oracle")
    assert_close("Vq", Vq, Vo, TOL, "query/mixgp_file/scalar_oracle")
    y1, v1, _ = P.querymixtureGP(Xq[3], eta, root, case["levels"], case["radius"], case["delta"], pk, case["sigma2"], wk)
    assert y1.shape == (1,) and y1[0] == Yq[3] and v1[0] == Vq[3]


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["BB10", "BB20"])
@pytest.mark.parametrize("N,Nq", [(15, 100), (1000, 100)])
def test_single_gp_ibb1d(built_lib, kind, N, Nq):
    """examples/IBB1D.jl: fitRKHS! + query! (mean only)."""
    case = cases.ibb1d(N, Nq, kind)
    ok, pk = helpers.kernels(case["kernel"])
    eta = P.RKHSProblemType(np.zeros(N), case["X"], pk, case["sigma2"])
    P.fitRKHS_(eta, case["y"])
    c_ref = O.fitRKHS(case["X"], case["y"], ok, case["sigma2"])
    yq = np.empty(Nq)
    P.query_(yq, case["Xq"], eta)
    yref = O.query_rkhs(case["Xq"], case["X"], c_ref, ok)
    # N=1000 Brownian-bridge Gram with sigma2=1e-5 is ill-conditioned (cond ~5e6): compare predictions, not weights
    assert_close("yq", yq, yref, TOL, f"single_gp/ibb1d_{kind}_{N}")
    if N == 15:
        np.testing.assert_allclose(eta.c, c_ref, rtol=1e-8)


def test_not_positive_definite(built_lib):
    """cholesky(U) throws PosDefException(info) (mixtureGP.jl:109): duplicated point, sigma2 = 0."""
    from patchmixturekriging_b200 import synth
    X = synth.uniform_points(4, 40, [-1.0, -1.0], [1.0, 1.0])
    X[29] = X[7]
    ok, pk = helpers.kernels(("SQEXP", 1.0))
    eta = P.MixtureGPType([X[:20], X], (np.zeros((1, 2)), np.zeros(1)))
    with pytest.raises(P.PosDefException) as ei:
        P.fitmixtureGP_(eta, [np.zeros(20), np.zeros(40)], pk, 0.0)
    with pytest.raises(O.PosDefException) as eo:
        U = O.constructkernelmatrix(X, ok)
        O.cholesky_L(U, 2)
    print("gpu info", ei.value.info, "leaf", ei.value.leaf, "| lapack info", eo.value.info)
    assert ei.value.leaf == 2
    assert ei.value.info == eo.value.info == 30


def test_argument_errors(built_lib):
    eta = P.MixtureGPType([np.zeros((4, 2))], (np.zeros((0, 2)), np.zeros(0)))
    with pytest.raises(P.PMKError):
        P.fitmixtureGP_(eta, [np.zeros(3)], P.GaussianKernel1DType(1.0), 1e-3)       # length(y) != length(X)
    with pytest.raises(P.PMKError):
        P.querymixtureGP(np.zeros((1, 2)), eta, None, 1, 0.1, 1e-5, P.GaussianKernel1DType(1.0), 1e-3, P.Spline34KernelType(1.0))


def test_leaf_size_edges(built_lib):
    """ragged leaves around the padding / size-class boundaries: 1, 31, 32, 33, 512, 513 points."""
    from patchmixturekriging_b200 import synth
    sizes = [1, 31, 32, 33, 512, 513]
    ok, pk = helpers.kernels(("SQEXP", 30.0))
    Xs = [synth.uniform_points(100 + s, s, [0.0, 0.0], [1.0, 1.0]) for s in sizes]
    ys = [np.sin(4 * X[:, 0]) + X[:, 1] for X in Xs]
    eta = P.MixtureGPType(Xs, (np.zeros((0, 2)), np.zeros(0)))
    P.fitmixtureGP_(eta, ys, pk, 1e-3)
    for i, (X, y) in enumerate(zip(Xs, ys)):
        U = O.constructkernelmatrix(X, ok) + 1e-3 * np.eye(len(X))
        Lref = O.cholesky_L(U)
        assert np.abs(eta.L_set[i] - Lref).max() < 1e-11
        cref = O.backslash(U, y)
        assert np.abs(U @ eta.c_set[i] - y).max() < 1e-12 * max(1.0, np.abs(cref).max())


@pytest.mark.parametrize("name", ["c3_mini", "c4_mini", "mixgp_file"])
def test_pruned_neighbour_search_equals_full_scan(built_lib, name):
    """The per-leaf candidate lists must reproduce the reference's scan over ALL hyperplanes bit for bit:
    same kept hyperplanes, same order, same t, same weights (and therefore the same Yq, Vq)."""
    case, m, root, eta, pk = _setup(name)
    _, wk = helpers.kernels(case["wkernel"])
    Xq = case["Xq"]
    L = _lib.lib()
    out = {}
    for full in (1, 0):
        eta.handle.check(L.pmk_set_option(eta.handle.raw, _lib.OPT_FULL_HYPERPLANE_SCAN, full))
        Yq, Vq, dv = P.querymixtureGP(Xq, eta, root, case["levels"], case["radius"], case["delta"], pk, case["sigma2"], wk,
                                      debug_flag=True)
        out[full] = (Yq, Vq, dv._flat)
    for key in ("home", "pair_off", "pair_leaf", "pair_hp", "pair_t", "pair_w", "pair_u", "pair_v"):
        assert np.array_equal(out[0][2][key], out[1][2][key]), key
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_single_gp_variance_query(built_lib):
    """setupGPquery / evalqueryGP! (src/RKHS/querying.jl:43-79): mean from the caller's c, unclamped variance k** - kᵀA⁻¹k."""
    from patchmixturekriging_b200 import synth
    X = synth.uniform_points(31, 300, [-1.0, -1.0], [1.0, 1.0])
    y = np.sin(3 * X[:, 0]) * X[:, 1]
    ok, pk = helpers.kernels(("SQEXP", 6.0))
    s2 = 1e-3
    A = O.constructkernelmatrix(X, ok) + s2 * np.eye(len(X))
    c = O.backslash(A, y)
    Xq = synth.uniform_points(32, 64, [-1.2, -1.2], [1.2, 1.2])
    kq = O.kernel_cross(Xq, X, ok)
    mean_ref = kq @ c
    var_ref = 1.0 - np.einsum("ij,ij->i", kq, np.linalg.solve(A, kq.T).T)      # evalqueryGP!: A\\k by LU, no clamp
    fq = P.setupGPquery(c, X, pk, s2)
    m, v = fq(Xq)
    assert_close("evalqueryGP mean", m, mean_ref)
    assert_close("evalqueryGP var", v, var_ref, 1e-8)
    m0, v0 = fq(Xq[5])
    assert m0 == m[5] and v0 == v[5]
    np.testing.assert_allclose(P.evalquery(Xq, c, X, pk), mean_ref, rtol=1e-12, atol=1e-13)
    # a query on top of a training point: variance ~ sigma2-level, may dip below the 1e-12 clamp of queryinner! but is not clamped here
    mm, vv = fq(X[:8])
    assert np.all(np.abs(vv) < 5 * s2)


@pytest.mark.parametrize("solver", [_lib.SOLVER_INVERSE, _lib.SOLVER_SUBSTITUTION])
@pytest.mark.parametrize("D,n", [(2, 700), (3, 1000), (1, 1100), (2, 1500), (3, 1700), (2, 2048)])
def test_large_leaf_size_classes(built_lib, D, n, solver):
    """One leaf of every large size class (n_pad <= 768, 1024, 1536, 2048; D = 1, 2, 3) through the fit, the inverse
    builder and the pair kernels (their chunk shapes and shared-memory fallbacks differ per class): posterior mean and
    latent variance against a dense numpy solve."""
    from patchmixturekriging_b200 import synth, mixturegp
    lo, hi = [-1.0] * D, [1.0] * D
    X = synth.uniform_points(41 + n, n, lo, hi)
    y = np.sin(3 * X[:, 0]) + (X[:, -1] if D > 1 else 0.0)
    spacing = (2.0 ** D / n) ** (1.0 / D)
    ok, pk = helpers.kernels(("SQEXP", 1.0 / (3.0 * spacing) ** 2))
    s2 = 1e-3
    A = O.constructkernelmatrix(X, ok) + s2 * np.eye(n)
    Xq = synth.uniform_points(43, 200, lo, hi)
    kq = O.kernel_cross(Xq, X, ok)
    mean_ref = kq @ np.linalg.solve(A, y)
    var_ref = np.maximum(1.0 - np.einsum("ij,ij->i", kq, np.linalg.solve(A, kq.T).T), 1e-12)
    eta = P.MixtureGPType([X], (np.zeros((0, D)), np.zeros(0)))
    P.fitmixtureGP_(eta, [y], pk, s2)
    mixturegp.set_query_solver(eta, solver)
    Yq, Vq, _ = P.querymixtureGP(Xq, eta, None, 1, 0.1, 1e-5, pk, s2, P.Spline34KernelType(1.0))
    # against a dense numpy LU solve of the full system (not the oracle's dpotrf + dtrsv): two algorithms apart
    rec = f"large_leaf/D{D}_n{n}/{SOLVER_NAME[solver]}"
    assert_close("mean", Yq, mean_ref, TOL, rec)
    assert_close("var", Vq, var_ref, TOL, rec)


def test_cholesky_variants_mixed_batch(built_lib):
    """PMK_OPT_CHOL_VARIANT: a batch with leaves on both sides of the size threshold (768 padded rows) through the default
    (by leaf size), the level-synchronous kernels only and the one-CTA-per-leaf kernel only -- every factor against numpy's
    dpotrf, alpha against a dense solve; and the default's choice depends on the leaf alone: a leaf fitted alone has the
    bits it has in the batch (what makes pmk_multi independent of the rank count)."""
    from patchmixturekriging_b200 import synth
    sizes = [40, 300, 767, 768, 800, 1300, 96, 1025]
    Xs, ys = [], []
    for k, n in enumerate(sizes):
        X = synth.uniform_points(91 + k, n, [-1.0, -1.0], [1.0, 1.0])
        Xs.append(X)
        ys.append(np.sin(3 * X[:, 0]) + X[:, 1])
    spacing = (4.0 / 800) ** 0.5
    ok, pk = helpers.kernels(("SQEXP", 1.0 / (3.0 * spacing) ** 2))
    s2 = 1e-3
    hps = (np.zeros((0, 2)), np.zeros(0))
    ref = []
    for X, y in zip(Xs, ys):
        A = O.constructkernelmatrix(X, ok) + s2 * np.eye(len(X))
        ref.append((np.linalg.cholesky(A), np.linalg.solve(A, y)))
    factors = {}
    for variant in (-1, 0, 1):
        eta = P.MixtureGPType(Xs, hps)
        eta.handle.check(_lib.lib().pmk_set_option(eta.handle.raw, _lib.OPT_CHOL_VARIANT, variant))
        P.fitmixtureGP_(eta, ys, pk, s2)
        for p, (Lr, ar) in enumerate(ref):
            L, a = eta.L_set[p], eta.c_set[p]
            assert np.abs(L - Lr).max() <= 3e-12 * np.abs(Lr).max(), (variant, sizes[p])
            assert np.abs(a - ar).max() <= 1e-9 * np.abs(ar).max(), (variant, sizes[p])
            factors[(variant, p)] = (L.copy(), a.copy())
        eta.close()
    for p in (2, 3, 5):          # alone, default variant: same bits as inside the batch
        eta = P.MixtureGPType([Xs[p]], hps)
        P.fitmixtureGP_(eta, [ys[p]], pk, s2)
        assert np.array_equal(eta.L_set[0], factors[(-1, p)][0]) and np.array_equal(eta.c_set[0], factors[(-1, p)][1])
        eta.close()


@pytest.mark.parametrize("name,eps", [("mixgp_file", 1.5), ("c3_mini", 0.31), ("c4_mini", 0.35), ("c3_mini", 0.0)])
def test_organizetrainingsets_device_bit_exact(built_lib, name, eps):
    """ε-overlap training sets on the GPU, bit-exact X_set_inds / regions_list_set."""
    case = CASES[name]()
    X = case["X"]
    root, _, _ = P.setuppartition(X, case["levels"])
    Xs_h, inds_h, rl_h, _ = P.organizetrainingsets(root, case["levels"], X, eps)
    Xs_d, inds_d, rl_d, prob = P.organizetrainingsets_device(root, case["levels"], X, eps)
    assert prob == [] and len(inds_d) == len(inds_h)
    for a, b, xa in zip(inds_d, inds_h, Xs_d):
        assert np.array_equal(a, b) and np.array_equal(xa, X[b - 1])
    for i in range(0, len(X), 53):
        assert list(rl_d[i]) == list(rl_h[i])


@pytest.mark.parametrize("name", ["mixgp_file", "c3_mini"])
def test_checkpoint_round_trip(built_lib, name, tmp_path):
    """savemixtureGP -> loadmixtureGP into a fresh handle: X_parts, c_set, L_set, hyperplanes and the query results
    (mean, variance, debug structure) are bit-identical to the handle that fitted; error paths keep their codes."""
    case, m, root, eta, pk = _setup(name)
    _, wk = helpers.kernels(case["wkernel"])
    path = tmp_path / "model.pmk"
    P.savemixtureGP(eta, path, root, case["levels"])
    eta2, root2, levels2 = P.loadmixtureGP(path)
    assert levels2 == case["levels"]
    assert np.array_equal(root2.hps_v, root.hps_v) and np.array_equal(root2.hps_c, root.hps_c)
    assert len(eta2.X_parts) == len(eta.X_parts)
    assert eta2.θ == pk and eta2.σ2_set[0] == case["sigma2"]
    for leaf in sorted({0, len(eta.X_parts) // 2, len(eta.X_parts) - 1}):
        assert np.array_equal(eta2.X_parts[leaf], eta.X_parts[leaf])
        assert np.array_equal(eta2.c_set[leaf], eta.c_set[leaf])
        assert np.array_equal(eta2.L_set[leaf], eta.L_set[leaf])
    Xq = case["Xq"][:4000]
    args = (case["levels"], case["radius"], case["delta"], pk, case["sigma2"], wk)
    Y0, V0, d0 = P.querymixtureGP(Xq, eta, root, *args, debug_flag=True)
    Y1, V1, d1 = P.querymixtureGP(Xq, eta2, root2, *args, debug_flag=True)
    assert np.array_equal(Y0, Y1) and np.array_equal(V0, V1)
    for key in ("home", "pair_off", "pair_leaf", "pair_hp", "pair_t", "pair_w", "pair_u", "pair_v"):
        assert np.array_equal(d0._flat[key], d1._flat[key]), key
    eta2.close()
    # error behaviour: not a model file / truncated file / save before fit
    bad = tmp_path / "bad.pmk"
    bad.write_bytes(b"not a model")
    with pytest.raises(P.PMKError) as e:
        P.loadmixtureGP(bad)
    assert e.value.code == _lib.PMK_ERR_ARG
    trunc = tmp_path / "trunc.pmk"
    trunc.write_bytes(path.read_bytes()[:4096])
    with pytest.raises(P.PMKError) as e:
        P.loadmixtureGP(trunc)
    assert e.value.code == _lib.PMK_ERR_ARG
    h = P.Handle(0)
    assert _lib.lib().pmk_save_model(h.raw, str(tmp_path / "x.pmk").encode()) == _lib.PMK_ERR_STATE
    h.close()


@pytest.mark.parametrize("N,levels,D", [(850, 3, 2), (5000, 6, 2), (4097, 4, 3), (3000, 2, 1), (16384, 7, 2), (12000, 5, 3),
                                        (300001, 10, 2)])
def test_setuppartition_device_bit_exact(built_lib, N, levels, D):
    """setuppartition with the per-level O(N) work on the GPU (pmk_partition_*): hyperplanes, X_parts_inds and X_parts are
    bit-identical to the oracle's (and the host mirror's) for every node, including the pairwise-summed means of nodes
    larger than 1024 points and the even/odd median rule."""
    from patchmixturekriging_b200 import synth
    X = synth.uniform_points(25, N, [-5.0, -10.0, -5.0][:D], [5.0, 10.0, 5.0][:D])
    root, X_parts, X_parts_inds = P.setuppartition_device(X, levels)
    if N <= 20000:
        oroot, _, oinds = O.setuppartition(X, levels)
        hv, hc = O.fetchhyperplanes(oroot)
    else:                                   # the oracle's recursion is slow here; the host mirror is pinned to it on CPU
        hroot, _, oinds = P.setuppartition(X, levels)
        hv, hc = hroot.hps_v, hroot.hps_c
    assert np.array_equal(root.hps_v, hv) and np.array_equal(root.hps_c, hc)
    assert len(X_parts_inds) == len(oinds) == 2 ** (levels - 1)
    for a, b, Xa in zip(X_parts_inds, oinds, X_parts):
        assert np.array_equal(a, b) and np.array_equal(Xa, X[b - 1])


def test_setuppartition_device_ties_and_errors(built_lib):
    """Regular grids give many equal projections (ties go right: f < c is strict, partition.jl:75); call-order and size
    errors keep their codes."""
    case = cases.c5_mini()
    X = case["X"]
    root, _, inds = P.setuppartition_device(X, case["levels"])
    oroot, _, oinds = O.setuppartition(X, case["levels"])
    hv, hc = O.fetchhyperplanes(oroot)
    assert np.array_equal(root.hps_v, hv) and np.array_equal(root.hps_c, hc)
    assert all(np.array_equal(a, b) for a, b in zip(inds, oinds))
    L = _lib.lib()
    h = P.Handle(0)
    z = np.empty((4, 2))
    assert L.pmk_partition_level_z(h.raw, 0, _lib.ptr(z)) == _lib.PMK_ERR_STATE
    Xs = np.ascontiguousarray(X[:3])
    assert L.pmk_partition_begin(h.raw, 2, 3, _lib.ptr(Xs), 4) == _lib.PMK_ERR_ARG          # fewer points than leaves
    assert L.pmk_partition_begin(h.raw, 2, 3, _lib.ptr(Xs), 1) == _lib.PMK_ERR_ARG          # levels must be > 1
    assert L.pmk_partition_begin(h.raw, 2, X.shape[0], _lib.ptr(X), 3) == _lib.PMK_OK
    c = np.empty(4)
    assert L.pmk_partition_level_split(h.raw, 0, _lib.ptr(z), _lib.ptr(c)) == _lib.PMK_ERR_STATE   # z of depth 0 not taken yet
    assert L.pmk_partition_level_z(h.raw, 1, _lib.ptr(z)) == _lib.PMK_ERR_STATE
    assert L.pmk_partition_fetch(h.raw, None, None) == _lib.PMK_ERR_STATE
    # identical points: every projection equals the median, the left child stays empty -> the reference fails in mean()
    Xc = np.ones((64, 2))
    assert L.pmk_partition_begin(h.raw, 2, 64, _lib.ptr(Xc), 3) == _lib.PMK_OK
    assert L.pmk_partition_level_z(h.raw, 0, _lib.ptr(z)) == _lib.PMK_OK
    v = np.array([[1.0, 0.0]])
    assert L.pmk_partition_level_split(h.raw, 0, _lib.ptr(v), _lib.ptr(c)) == _lib.PMK_ERR_ARG
    h.close()


@pytest.mark.parametrize("kname,param,D,n", [("SQEXP", 3.0, 2, 1000), ("SPLINE34", 0.4, 3, 777), ("BB20", 1.0, 1, 1300), ("RQ", 1.7, 2, 2049)])
def test_gram_mirrored_large(built_lib, kname, param, D, n):
    """constructkernelmatrix beyond a few tiles: the tiles above the diagonal are written by the tiles below them (each kernel
    value evaluated once, RKHS.jl:27-31 mirrors the lower triangle the same way) -- exact symmetry, oracle values, odd and even n."""
    from patchmixturekriging_b200 import synth
    lo, hi = ([0.0] * D, [1.0] * D) if kname == "BB20" else ([-1.0] * D, [1.0] * D)
    X = synth.uniform_points(4, n, lo, hi)
    ok, pk = helpers.kernels((kname, param))
    K = P.constructkernelmatrix(X, pk)
    assert np.array_equal(K, K.T)
    ref = O.constructkernelmatrix(X, ok)
    np.testing.assert_allclose(K, ref, rtol=5e-15, atol=1e-300)
