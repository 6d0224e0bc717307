"""Sub-tree ownership on the GPU (pmk_multi, include/pmk.h): a model sharded over several ranks must answer exactly like one
handle that owns every leaf -- a pair's u, v depend on its leaf and its query point only, and the planner combines in the
reference's slot order.  On a one-GPU box the ranks are several handles on the same device (the library allows a device
ordinal to repeat; the peer copies become device-to-device copies); with two or more GPUs the same tests also run on
distinct devices."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

import cases
import helpers
import patchmixturekriging_b200 as P
from patchmixturekriging_b200 import _lib, sharding

pytestmark = pytest.mark.gpu

_cache = {}


def _model(name):
    """single-handle reference model of a case + everything needed to build sharded twins"""
    if name not in _cache:
        case = getattr(cases, name)()
        _, pk = helpers.kernels(case["kernel"])
        _, wk = helpers.kernels(case["wkernel"])
        root, _, _ = P.setuppartition(case["X"], case["levels"])
        X_set, X_set_inds, _, _ = P.organizetrainingsets(root, case["levels"], case["X"], case["eps"])
        y_set = [case["y"][i - 1] for i in X_set_inds]
        eta = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
        P.fitmixtureGP_(eta, y_set, pk, case["sigma2"])
        _cache[name] = (case, root, X_set, y_set, pk, wk, eta)
    return _cache[name]


def _device_sets():
    import torch
    sets = [[0], [0, 0], [0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        sets.append([0, 1])
    if torch.cuda.device_count() >= 4:
        sets.append([0, 1, 2, 3])
    return sets


@pytest.mark.parametrize("name", ["c3_mini", "c4_mini", "mixgp_file"])
def test_multi_equals_single_bit_for_bit(built_lib, name):
    case, root, X_set, y_set, pk, wk, eta = _model(name)
    args = (root, case["levels"], case["radius"], case["delta"], pk, case["sigma2"], wk)
    Xq = case["Xq"][:9001]
    Y0, V0, dv = P.querymixtureGP(Xq, eta, *args, debug_flag=True)
    per_leaf0 = np.bincount(dv._flat["pair_leaf"] - 1, minlength=len(X_set))
    for devs in _device_sets():
        if len(devs) > len(X_set):
            continue
        em = P.MixtureGPType(X_set, P.fetchhyperplanes(root), devices=devs)
        P.fitmixtureGP_(em, y_set, pk, case["sigma2"])
        Y1, V1, _ = P.querymixtureGP(Xq, em, *args)
        assert np.array_equal(Y0, Y1) and np.array_equal(V0, V1), f"devices {devs}"
        # every leaf lives on exactly one rank, with global ids: factors and weights come from the owner
        for leaf in sorted({0, len(X_set) // 2 - 1, len(X_set) // 2, len(X_set) - 1}):
            assert np.array_equal(em.c_set[leaf], eta.c_set[leaf])
            assert np.array_equal(em.L_set[leaf], eta.L_set[leaf])
        # the leaf -> rank map: contiguous ranges that cover the leaves once, every range within one leaf of its share of sum(n^3)
        ranges = [em.multi.owned_range(r) for r in range(len(devs))]
        assert ranges[0][0] == 0 and all(c >= 1 for _, c in ranges)
        assert all(ranges[r][0] + ranges[r][1] == (ranges[r + 1][0] if r + 1 < len(devs) else len(X_set)) for r in range(len(devs)))
        first = np.zeros(len(devs) + 1, dtype=np.int64)
        lo = np.concatenate([[0], np.cumsum([x.shape[0] for x in X_set])]).astype(np.int64)
        assert _lib.lib().pmk_multi_balanced_ranges(len(devs), len(X_set), _lib.ptr(lo), _lib.ptr(first)) == _lib.PMK_OK
        assert [a for a, _ in ranges] == list(first[:-1])          # the host-only export is the map pmk_multi uses
        cost = np.array([float(x.shape[0]) ** 3 for x in X_set])
        if len(X_set) >= 4 * len(devs):
            assert max(cost[a:a + c].sum() for a, c in ranges) <= cost.sum() / len(devs) + cost.max() * (1 + 1e-12)
        # pairs per leaf, summed over the planners
        pl = np.empty(len(X_set), dtype=np.int64)
        em.multi.check(_lib.lib().pmk_multi_leaf_pairs(em.multi.raw, _lib.ptr(pl)))
        assert np.array_equal(pl, per_leaf0)
        # a second query on the same model (buffers reused, other size), then a refit
        Y2, V2, _ = P.querymixtureGP(Xq[:777], em, *args)
        assert np.array_equal(Y2, Y0[:777]) and np.array_equal(V2, V0[:777])
        ms, per = em.multi.timings()
        assert ms[_lib.MT_QUERY] > 0 and ms[_lib.MT_FIT] > 0 and per.shape == (len(devs), _lib.T_COUNT)
        em.close()


def test_multi_staged_and_mean_only(built_lib):
    """the *_staged forms (inputs and results resident in HBM between calls) and the mean-only flag"""
    case, root, X_set, y_set, pk, wk, eta = _model("c3_mini")
    L = _lib.lib()
    Xq = np.ascontiguousarray(case["Xq"][:5000])
    args = (root, case["levels"], case["radius"], case["delta"], pk, case["sigma2"], wk)
    Y0, V0, _ = P.querymixtureGP(Xq, eta, *args)
    sizes = np.array([len(x) for x in X_set], dtype=np.int64)
    leaf_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    Xp = np.ascontiguousarray(np.concatenate(X_set))
    yp = np.ascontiguousarray(np.concatenate(y_set))
    hv, hc = np.ascontiguousarray(root.hps_v), np.ascontiguousarray(root.hps_c)
    kp, wp = pk.params, wk.params
    m = _lib.MultiHandle([0, 0])
    assert L.pmk_multi_query_staged(m.raw, 0.1, 1e-5, wk.kernel_id, _lib.ptr(wp), 1, 0) == _lib.PMK_ERR_STATE     # before fit
    m.check(L.pmk_multi_stage_training(m.raw, 2, len(sizes), _lib.ptr(leaf_off), _lib.ptr(Xp), _lib.ptr(yp)))
    bad, info = C.c_int64(0), C.c_int(0)
    for _ in range(2):      # fit twice from the same resident inputs
        m.check(L.pmk_multi_fit_staged(m.raw, pk.kernel_id, _lib.ptr(kp), kp.shape[0], case["sigma2"], C.byref(bad), C.byref(info)))
    m.check(L.pmk_multi_set_tree(m.raw, 2, case["levels"], _lib.ptr(hv), _lib.ptr(hc)))
    m.check(L.pmk_multi_stage_queries(m.raw, len(Xq), _lib.ptr(Xq)))
    Y1, V1 = np.empty(len(Xq)), np.empty(len(Xq))
    for _ in range(2):      # query twice from the same resident queries
        m.check(L.pmk_multi_query_staged(m.raw, case["radius"], case["delta"], wk.kernel_id, _lib.ptr(wp), wp.shape[0], 0))
        m.check(L.pmk_multi_fetch_results(m.raw, _lib.ptr(Y1), _lib.ptr(V1)))
        assert np.array_equal(Y1, Y0) and np.array_equal(V1, V0)
    Ym = np.empty(len(Xq))
    m.check(L.pmk_multi_query(m.raw, len(Xq), _lib.ptr(Xq), case["radius"], case["delta"], wk.kernel_id, _lib.ptr(wp), wp.shape[0], 1,
                              _lib.ptr(Ym), None))
    # mean only (flags bit 0, Vq untouched): the mean-only pair kernel, bit-identical to the single handle's mean-only query
    Ys = np.empty(len(Xq))
    h = eta.handle
    h.check(L.pmk_set_tree(h.raw, 2, case["levels"], _lib.ptr(hv), _lib.ptr(hc)))
    h.check(L.pmk_query(h.raw, len(Xq), _lib.ptr(Xq), case["radius"], case["delta"], wk.kernel_id, _lib.ptr(wp), wp.shape[0], 1, _lib.ptr(Ys), None))
    assert np.array_equal(Ym, Ys)
    np.testing.assert_allclose(Ym, Y0, rtol=0, atol=1e-12 * np.abs(Y0).max())
    assert m.launch_count() > 0
    m.close()


def test_routed_building_blocks(built_lib):
    """plan -> segments -> pack -> routed pair kernel -> unpack -> combine on ONE handle equals pmk_query_dev bit for bit, and
    the segments are the owner ranges of the leaf-sorted pair list; a handle that owns a sub-tree refuses pairs of foreign
    leaves."""
    import torch
    case, root, X_set, y_set, pk, wk, eta = _model("c3_mini")
    L = _lib.lib()
    h = eta.handle
    Xq = np.ascontiguousarray(case["Xq"][:6000])
    args = (root, case["levels"], case["radius"], case["delta"], pk, case["sigma2"], wk)
    Y0, V0, dv = P.querymixtureGP(Xq, eta, *args, debug_flag=True)
    n_leaves = len(X_set)
    wp = wk.params
    dXq = torch.from_numpy(Xq).cuda()
    npairs = C.c_int64(0)
    h.check(L.pmk_query_plan_dev(h.raw, len(Xq), dXq.data_ptr(), case["radius"], case["delta"], wk.kernel_id, _lib.ptr(wp), wp.shape[0],
                                 C.byref(npairs)))
    n = npairs.value
    first = sharding.owner_first_leaf(3, n_leaves)
    seg = np.empty(4, dtype=np.int64)
    h.check(L.pmk_query_plan_segments(h.raw, 3, _lib.ptr(first), _lib.ptr(seg)))
    sorted_leaf = np.sort(dv._flat["pair_leaf"], kind="stable")
    assert np.array_equal(seg, sharding.segments(sorted_leaf, first)) and seg[0] == 0 and seg[-1] == n
    dXs = torch.empty((n, 2), dtype=torch.float64, device="cuda")
    dls = torch.empty(n, dtype=torch.int32, device="cuda")
    h.check(L.pmk_query_plan_pack_dev(h.raw, dXs.data_ptr(), dls.data_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(dls.cpu().numpy(), sorted_leaf)
    us, vs = torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")
    # owner side, fed in REVERSED order: the answers must come back in the order the pairs were given
    idx = torch.arange(n - 1, -1, -1, device="cuda")
    ur, vr = torch.empty_like(us), torch.empty_like(vs)
    Xr, lr = dXs[idx].contiguous(), dls[idx].contiguous()
    h.check(L.pmk_query_pairs_routed_dev(h.raw, n, Xr.data_ptr(), lr.data_ptr(), 0, ur.data_ptr(), vr.data_ptr()))
    torch.cuda.synchronize()
    us[idx], vs[idx] = ur, vr
    pu, pv = torch.empty_like(us), torch.empty_like(vs)
    h.check(L.pmk_query_plan_unpack_dev(h.raw, us.data_ptr(), vs.data_ptr(), pu.data_ptr(), pv.data_ptr()))
    h.check(L.pmk_query_set_flags(h.raw, 0))
    dY, dV = torch.empty(len(Xq), dtype=torch.float64, device="cuda"), torch.empty(len(Xq), dtype=torch.float64, device="cuda")
    h.check(L.pmk_query_combine_dev(h.raw, pu.data_ptr(), pv.data_ptr(), dY.data_ptr(), dV.data_ptr()))
    h.synchronize()
    assert np.array_equal(pu.cpu().numpy(), dv._flat["pair_u"]) and np.array_equal(pv.cpu().numpy(), dv._flat["pair_v"])
    assert np.array_equal(dY.cpu().numpy(), Y0) and np.array_equal(dV.cpu().numpy(), V0)
    # a sub-tree owner: leaves [half, n_leaves) only
    half = n_leaves // 2
    ho = P.Handle(0)
    ho.check(L.pmk_set_leaf_base(ho.raw, half, n_leaves))
    sizes = np.array([len(x) for x in X_set[half:]], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    Xp, yp = np.ascontiguousarray(np.concatenate(X_set[half:])), np.ascontiguousarray(np.concatenate(y_set[half:]))
    kp = pk.params
    ho.check(L.pmk_fit(ho.raw, 2, len(sizes), _lib.ptr(off), _lib.ptr(Xp), _lib.ptr(yp), pk.kernel_id, _lib.ptr(kp), kp.shape[0],
                       case["sigma2"], None, None))
    a = np.empty(len(X_set[half]))
    ho.check(L.pmk_get_alpha(ho.raw, half + 1, _lib.ptr(a)))                  # global 1-based leaf ids
    assert np.array_equal(a, eta.c_set[half])
    assert L.pmk_get_alpha(ho.raw, half, _lib.ptr(a)) == _lib.PMK_ERR_ARG       # not owned
    lo, hi = int(seg[0]), int(sharding.segments(sorted_leaf, [half])[0])
    mine = slice(hi, n)
    k = n - hi
    ho.check(L.pmk_query_pairs_routed_dev(ho.raw, k, dXs[mine].contiguous().data_ptr(), dls[mine].contiguous().data_ptr(), 0,
                                          ur.data_ptr(), vr.data_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(ur[:k].cpu().numpy(), us[mine].cpu().numpy()) and np.array_equal(vr[:k].cpu().numpy(), vs[mine].cpu().numpy())
    if hi > lo:
        rc = L.pmk_query_pairs_routed_dev(ho.raw, hi - lo, dXs[lo:hi].contiguous().data_ptr(), dls[lo:hi].contiguous().data_ptr(), 0,
                                          ur.data_ptr(), vr.data_ptr())
        assert rc == _lib.PMK_ERR_ARG                                          # foreign leaves
    assert L.pmk_save_model(ho.raw, b"/tmp/pmk_partial.pmk") == _lib.PMK_ERR_STATE   # a model file holds a whole model
    ho.close()


def test_multi_errors(built_lib):
    from patchmixturekriging_b200 import synth
    L = _lib.lib()
    with pytest.raises(P.PMKError):
        _lib.MultiHandle([999])
    # a leaf that is not positive definite: the first failing leaf's GLOBAL id comes back (mixtureGP.jl:109), whichever rank owns it
    X = synth.uniform_points(4, 40, [-1.0, -1.0], [1.0, 1.0])
    Xd = X.copy()
    Xd[29] = Xd[7]
    pk = P.GaussianKernel1DType(1.0)
    em = P.MixtureGPType([X[:20], X, Xd, X[:30]], (np.zeros((3, 2)), np.zeros(3)), devices=[0, 0])
    with pytest.raises(P.PosDefException) as ei:
        P.fitmixtureGP_(em, [np.zeros(20), np.zeros(40), np.zeros(40), np.zeros(30)], pk, 0.0)
    assert ei.value.leaf == 3 and ei.value.info == 30
    # fewer leaves than ranks
    e1 = P.MixtureGPType([X], (np.zeros((0, 2)), np.zeros(0)), devices=[0, 0])
    with pytest.raises(P.PMKError) as e:
        P.fitmixtureGP_(e1, [np.zeros(40)], pk, 1e-3)
    assert e.value.code == _lib.PMK_ERR_ARG
    em.close()
    e1.close()
