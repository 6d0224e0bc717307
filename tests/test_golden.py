"""Golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle on the reference's example
workloads): the oracle must keep reproducing them (CPU), and the CUDA path must match them (GPU)."""
from __future__ import annotations

import os

import numpy as np
import pytest

import cases
import helpers
from oracle import pmk_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MIX = {"mixgp_file": (cases.mixgp_file, 50), "mixgp_sqexp": (lambda: cases.mixgp_driver(N=3000, levels=4), 50)}


@pytest.mark.parametrize("name", list(MIX))
def test_oracle_reproduces_golden(name):
    mk, stride = MIX[name]
    case = mk()
    g = np.load(os.path.join(G, f"{name}.npz"))
    m = helpers.oracle_model(case)
    assert np.array_equal(m["hv"], g["hps_v"]) and np.array_equal(m["hc"], g["hps_c"])
    assert np.array_equal([len(i) for i in m["X_set_inds"]], g["set_sizes"])
    assert np.array_equal([int(i.sum()) for i in m["X_set_inds"]], g["set_inds_sum"])
    oth, _ = helpers.kernels(case["kernel"])
    owth, _ = helpers.kernels(case["wkernel"])
    Xq = case["Xq"][::stride]
    assert np.array_equal(Xq, g["Xq"])
    Y, V, d = O.querymixtureGP_vec(Xq, m["eta"], case["levels"], case["radius"], case["delta"], oth, owth)
    assert np.array_equal(d["home"], g["home"]) and np.array_equal(np.diff(d["pair_off"]), g["npairs"])
    assert np.abs(Y - g["Yq"]).max() <= 1e-9 * np.abs(g["Yq"]).max()
    assert np.abs(V - g["Vq"]).max() <= 2e-8 * np.abs(g["Vq"]).max()
    np.testing.assert_allclose(m["eta"].c_set[0][:16], g["alpha0_head"], rtol=1e-6)
    np.testing.assert_allclose(np.diag(m["eta"].L_set[0]), g["L0_diag"], rtol=1e-10)


@pytest.mark.parametrize("kind", ["BB10", "BB20"])
def test_oracle_reproduces_golden_ibb1d(kind):
    case = cases.ibb1d(15, 100, kind)
    g = np.load(os.path.join(G, f"ibb1d_{kind}.npz"))
    th, _ = helpers.kernels(case["kernel"])
    assert np.array_equal(O.constructkernelmatrix(case["X"], th), g["K"])
    c = O.fitRKHS(case["X"], case["y"], th, case["sigma2"])
    np.testing.assert_allclose(O.query_rkhs(case["Xq"], case["X"], c, th), g["yq"], rtol=1e-9, atol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MIX))
def test_gpu_matches_golden(built_lib, name):
    import patchmixturekriging_b200 as P
    mk, stride = MIX[name]
    case = mk()
    g = np.load(os.path.join(G, f"{name}.npz"))
    _, θ = helpers.kernels(case["kernel"])
    _, wθ = helpers.kernels(case["wkernel"])
    root, _, _ = P.setuppartition(case["X"], case["levels"])
    assert np.array_equal(root.hps_v, g["hps_v"]) and np.array_equal(root.hps_c, g["hps_c"])
    droot, _, _ = P.setuppartition_device(case["X"], case["levels"])          # the level-stepped device build, same golden tree
    assert np.array_equal(droot.hps_v, g["hps_v"]) and np.array_equal(droot.hps_c, g["hps_c"])
    X_set, X_set_inds, _, _ = P.organizetrainingsets_device(root, case["levels"], case["X"], case["eps"])
    assert np.array_equal([len(i) for i in X_set_inds], g["set_sizes"])
    assert np.array_equal([int(i.sum()) for i in X_set_inds], g["set_inds_sum"])
    assert np.array_equal(np.concatenate([i[:8] for i in X_set_inds]), g["set_inds_head"])
    η = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
    P.fitmixtureGP_(η, [case["y"][i - 1] for i in X_set_inds], θ, case["sigma2"])
    Yq, Vq, dv = P.querymixtureGP(g["Xq"], η, root, case["levels"], case["radius"], case["delta"], θ, case["sigma2"], wθ,
                                  debug_flag=True)
    assert np.array_equal(dv.p_region_ind_set, g["home"])
    assert np.array_equal(np.diff(dv.pair_off), g["npairs"])
    assert np.array_equal(np.concatenate([r for r in dv.region_inds_set] + [np.zeros(0, np.int32)]), g["region_inds"])
    tol = 1e-9          # every golden case, sigma2 = 1e-5 included: the default solver is chosen by conditioning
    sy, sv = np.sqrt(np.mean(g["Yq"] ** 2)), np.sqrt(np.mean(g["Vq"] ** 2))
    helpers.record_parity(f"golden/{name}", Yq=helpers.bound_stats(Yq, g["Yq"]), Vq=helpers.bound_stats(Vq, g["Vq"]), tol=tol)
    assert np.all(np.abs(Yq - g["Yq"]) <= tol * np.maximum(np.abs(g["Yq"]), sy))
    assert np.all(np.abs(Vq - g["Vq"]) <= tol * np.maximum(np.abs(g["Vq"]), sv))
    np.testing.assert_allclose(np.diag(η.L_set[0]), g["L0_diag"], rtol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["BB10", "BB20"])
def test_gpu_matches_golden_ibb1d(built_lib, kind):
    import patchmixturekriging_b200 as P
    case = cases.ibb1d(15, 100, kind)
    g = np.load(os.path.join(G, f"ibb1d_{kind}.npz"))
    _, θ = helpers.kernels(case["kernel"])
    np.testing.assert_allclose(P.constructkernelmatrix(case["X"], θ), g["K"], rtol=5e-15, atol=0)
    η = P.RKHSProblemType(np.zeros(15), case["X"], θ, case["sigma2"])
    P.fitRKHS_(η, case["y"])
    yq = np.empty(100)
    P.query_(yq, case["Xq"], η)
    np.testing.assert_allclose(yq, g["yq"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(η.c, g["c"], rtol=1e-7)
