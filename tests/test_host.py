"""CPU tests of the host layer: BSP mirror vs oracle (bit-exact), the C-ABI library loads and exports
every symbol include/pmk.h declares, and the product fails loudly without a GPU (no fallback)."""
from __future__ import annotations

import ctypes
import os
import re

import numpy as np
import pytest

import cases
import patchmixturekriging_b200 as P
from oracle import pmk_oracle as O
from patchmixturekriging_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("N,levels,D", [(850, 3, 2), (5000, 6, 2), (4097, 4, 3), (3000, 2, 1)])
def test_setuppartition_bit_exact(N, levels, D):
    from patchmixturekriging_b200 import synth
    X = synth.uniform_points(25, N, [-5.0, -10.0, -5.0][:D], [5.0, 10.0, 5.0][:D])
    root, X_parts, X_parts_inds = P.setuppartition(X, levels)
    oroot, oX_parts, oinds = O.setuppartition(X, levels)
    hv, hc = O.fetchhyperplanes(oroot)
    assert np.array_equal(root.hps_v, hv) and np.array_equal(root.hps_c, hc)
    assert len(X_parts_inds) == len(oinds) == 2 ** (levels - 1)
    for a, b, Xa in zip(X_parts_inds, oinds, X_parts):
        assert np.array_equal(a, b) and np.array_equal(Xa, X[b - 1])
    v2, c2 = P.fetchhyperplanes(root)
    assert len(c2) == len(X_parts) - 1
    q = synth.uniform_points(9, 700, [-6.0, -11.0, -6.0][:D], [6.0, 11.0, 6.0][:D])
    assert np.array_equal(P.findpartition(q, root), O._descend_vec(q, hv, hc, levels))
    assert P.findpartition(q[5], root) == O.findpartition(q[5], oroot, levels)


@pytest.mark.parametrize("eps", [0.0, 0.3, 1.5])
def test_organizetrainingsets_bit_exact(eps):
    case = cases.mixgp_driver(N=4000, levels=5)
    X = case["X"]
    root, _, _ = P.setuppartition(X, 5)
    oroot, _, _ = O.setuppartition(X, 5)
    X_set, X_set_inds, rl, prob = P.organizetrainingsets(root, 5, X, eps)
    oX_set, oinds, orl, _ = O.organizetrainingsets(oroot, 5, X, eps)
    assert prob == []
    for a, b, Xa in zip(X_set_inds, oinds, X_set):
        assert np.array_equal(a, b) and np.array_equal(Xa, X[b - 1])
    for i in range(0, 4000, 37):
        assert list(rl[i]) == orl[i]
    if eps == 0.0:
        assert sum(len(a) for a in X_set_inds) == 4000


def test_svd_row_form_switch():
    X = cases.mixgp_file()["X"]
    r1, _, _ = P.setuppartition(X, 3, svd_form="row")
    o1, _, _ = O.setuppartition(X, 3, svd_form="row")
    assert np.array_equal(r1.hps_v, O.fetchhyperplanes(o1)[0])


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "pmk.h")).read()
    declared = set(re.findall(r"\b(pmk_[a-z_A-Z0-9]+)\s*\(", hdr))
    declared -= {"pmk_handle"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = ctypes.CDLL(built_lib)
    for s in declared:
        assert hasattr(L, s), s
    assert L.pmk_version() >= 100


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product refuses to run (it never routes through the oracle)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(P.PMKError) as e:
        P.Handle(0)
    assert e.value.code == _lib.PMK_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(P.PMKError):
        P.constructkernelmatrix(np.zeros((3, 2)), P.GaussianKernel1DType(1.0))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "patchmixturekriging_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "pmk_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f
                assert "dlopen" not in src and "libpmk_oracle" not in src, f


def test_kernel_descriptors():
    assert P.Spline34KernelType(0.5).kernel_id == 1 and P.Spline34KernelType(0.5).params[0] == 0.5
    assert P.GaussianKernel1DType(400.0).kernel_id == 0 and P.GaussianKernel1DType(400.0).stationary
    assert P.BrownianBridge10().kernel_id == 2 and not P.BrownianBridge10().stationary
    assert P.BrownianBridge2ϵ(2.5).params[0] == 2.5


def test_exp_neg_tab_arithmetic():
    """The row-panel pair kernel's inline exp (pmk_common.cuh exp_neg_tab: 64-entry table of 2^(j/64), the integer
    64k + j read from the low word of t*64/ln2 + 1.5*2^52, degree-5 polynomial) restated in numpy: <= 1.5 ulp from the
    correctly rounded value on the argument range the squared-exponential cross-covariance produces."""
    from decimal import Decimal, getcontext
    getcontext().prec = 50
    ln2 = Decimal(2).ln()
    tab = np.array([float((ln2 * Decimal(j) / 64).exp()) for j in range(64)])
    # the table compiled into the library is this one, bit for bit
    import re, os
    src = open(os.path.join(os.path.dirname(__file__), "..", "patchmixturekriging_b200", "csrc", "pmk_common.cuh")).read()
    lits = re.findall(r"0x1\.[0-9a-f]{13}p\+0", src[src.index("c_exp2_64[64]"):src.index("};", src.index("c_exp2_64[64]"))])
    assert len(lits) == 64 and all(float.fromhex(a) == b for a, b in zip(lits, tab))
    C, HI, LO = float.fromhex("0x1.71547652b82fep+6"), float.fromhex("0x1.62e42fee00000p-7"), float.fromhex("0x1.a39ef35793c76p-39")
    MAGIC = float.fromhex("0x1.8p52")
    rng = np.random.default_rng(7)
    t = -np.concatenate([rng.uniform(0, 50, 400000), rng.uniform(0, 699, 400000), [0.0, 1e-300, 1e-9, 699.0]])
    kd = t * C + MAGIC
    ki = (kd.view(np.int64) & 0xFFFFFFFF).astype(np.int64)
    ki = np.where(ki >= 2 ** 31, ki - 2 ** 32, ki)
    kd = kd - MAGIC
    assert np.array_equal(kd, ki.astype(float))
    r = (t - kd * HI) - kd * LO
    assert np.abs(r).max() <= 0.0054152056
    T = tab[ki & 63]
    r2 = r * r
    pr = r + ((0.5 + r / 6) + (1 / 24 + r / 120) * r2) * r2
    got = np.ldexp(T + T * pr, (ki >> 6).astype(int))
    ref = np.exp(t.astype(np.longdouble))
    err = np.abs(got.astype(np.longdouble) - ref) / np.spacing(np.exp(t)).astype(np.longdouble)
    assert err.max() <= 1.5
    # exp_neg_tab_s (the pair kernel's phase E) clamps with one unsigned minimum on the high word instead of fmax(t, -699):
    # the same value for every t in [-699, -0.0], within 2^-11 below -699 for anything smaller (-inf included)
    tt = -np.concatenate([rng.uniform(0, 699, 100000), rng.uniform(699, 1e6, 1000), [0.0, 699.0, 699.0000001, 1e300, np.inf]])
    hi = (tt.view(np.uint64) >> np.uint64(32)).astype(np.uint64)
    lo = tt.view(np.uint64) & np.uint64(0xFFFFFFFF)
    clamped = ((np.minimum(hi, np.uint64(0xC085D800)) << np.uint64(32)) | lo).view(np.float64)
    inside = tt >= -699.0
    assert np.array_equal(clamped[inside], tt[inside])
    assert np.all(clamped[~inside] <= -699.0) and np.all(clamped[~inside] > -699.0005)      # one high-word step = 2^-11


def test_inverse_plan_covers_every_block_once(built_lib):
    """The recursive-doubling plan of P = inv(L) (pmk_invert.cu, host side): for every leaf shape the nodes' off-diagonal
    rectangles tile the strictly-lower block triangle exactly once, and a node's children are finished at a lower height."""
    import ctypes as C
    L = _lib.lib()
    for nb in list(range(1, 25)) + [31, 32, 33, 47, 48, 63, 64]:
        n = C.c_int(0)
        assert L.pmk_inverse_plan(nb, 0, None, C.byref(n)) == 0
        nodes = np.zeros((max(n.value, 1), 4), dtype=np.int16)
        assert L.pmk_inverse_plan(nb, n.value, nodes.ctypes.data_as(C.c_void_p), C.byref(n)) == 0
        nodes = nodes[:n.value]
        assert n.value == nb - 1                                    # a binary tree over nb blocks
        cover = np.zeros((nb, nb), dtype=int)
        done_at = {(b, b + 1): 0 for b in range(nb)}                # single blocks: the diagonal inverses exist (height 0)
        assert np.all(np.diff(nodes[:, 3]) >= 0)                    # launch order = ascending height
        for lo, mid, hi, h in nodes:
            assert 0 <= lo < mid < hi <= nb
            cover[mid:hi, lo:mid] += 1
            assert done_at[(lo, mid)] < h and done_at[(mid, hi)] < h
            done_at[(lo, hi)] = h
        assert np.array_equal(cover, np.tril(np.ones((nb, nb), dtype=int), -1))
        assert (0, nb) in done_at
    assert L.pmk_inverse_plan(0, 0, None, C.byref(n)) != 0 and L.pmk_inverse_plan(65, 0, None, C.byref(n)) != 0


def _sum_plan(n):
    L = ctypes.CDLL(_lib.LIB_PATH)
    L.pmk_partition_sum_plan.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.POINTER(ctypes.c_int64)]
    cap = n // 512 + 2
    st, ln, dp = np.zeros(cap, np.int64), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    nb = ctypes.c_int64(0)
    assert L.pmk_partition_sum_plan(n, cap, st.ctypes.data, ln.ctypes.data, dp.ctypes.data, ctypes.byref(nb)) == 0
    return st[:nb.value], ln[:nb.value], dp[:nb.value]


def _planned_sum(X):
    """What k_part_block_sums + k_part_node_z do (csrc/pmk_partition.cu), in numpy: sequential block sums, then the stack
    combination by depth."""
    st, ln, dp = _sum_plan(X.shape[0])
    assert st[0] == 0 and np.array_equal(st[1:], np.cumsum(ln)[:-1]) and ln.sum() == X.shape[0] and ln.max() <= 1024
    stack = []
    for s, l, d in zip(st, ln, dp):
        stack.append([int(d), np.add.accumulate(X[s:s + l], axis=0)[-1]])
        while len(stack) >= 2 and stack[-1][0] == stack[-2][0]:
            b = stack.pop()
            stack[-1] = [stack[-1][0] - 1, stack[-1][1] + b[1]]
    assert len(stack) == 1 and stack[0][0] == 0
    return stack[0][1]


@pytest.mark.parametrize("n", [1, 2, 15, 16, 1023, 1024, 1025, 2048, 2049, 3071, 4097, 10000, 65537, 300001])
def test_partition_sum_plan_is_base_pairwise_order(built_lib, n):
    """pmk_partition_sum_plan (host-only export) + the device's combination rule reproduce Statistics.mean's summation
    order (oracle mean_pairwise; partition.jl:89) bit for bit."""
    from patchmixturekriging_b200 import synth
    X = synth.uniform_points(3, n, [-5.0, -10.0, 1e3], [5.0, 10.0, 1e3 + 1.0])
    got = _planned_sum(X) / n
    # the oracle's recursion (forced: its short-cut for n <= 1024 is the same order)
    assert np.array_equal(got, O.mean_pairwise(X))


def test_partition_level_algorithm_in_numpy():
    """The level-by-level formulation of setuppartition the device follows (stable in-place splits of contiguous segments,
    left-to-right node order, preorder_index) yields the reference's tree: restated in numpy and compared with the oracle."""
    from patchmixturekriging_b200 import synth
    from patchmixturekriging_b200.partition import _split_direction, preorder_index
    for N, levels, D in [(850, 3, 2), (6000, 6, 3), (3000, 4, 1)]:
        X = synth.uniform_points(25, N, [-5.0, -10.0, -5.0][:D], [5.0, 10.0, 5.0][:D])
        perm = np.arange(N)
        seg = np.array([0, N])
        n_hp = (1 << (levels - 1)) - 1
        hv, hc = np.empty((n_hp, D)), np.empty(n_hp)
        for depth in range(levels - 1):
            nodes = 1 << depth
            node_id = np.searchsorted(seg, np.arange(N), side="right") - 1
            v = np.empty((nodes, D)); c = np.empty(nodes)
            f = np.empty(N)
            for j in range(nodes):
                Xj = X[perm[seg[j]:seg[j + 1]]]
                z = Xj[0] - _planned_sum(Xj) / Xj.shape[0]
                v[j] = _split_direction(z, "column")
                fj = O.dot_seq(v[j][None, :], Xj)
                f[seg[j]:seg[j + 1]] = fj
                fs = np.sort(fj)
                n = len(fj)
                c[j] = fs[(n - 1) // 2] if n & 1 else fs[n // 2 - 1] / 2.0 + fs[n // 2] / 2.0
                k = preorder_index(depth, j, levels)
                hv[k], hc[k] = v[j], c[j]
            flag = np.concatenate([(f < c[node_id]).astype(np.int64), [0]])
            scan = np.concatenate([[0], np.cumsum(flag)[:-1]])
            s, e = seg[node_id], seg[node_id + 1]
            lt, lb = scan[e] - scan[s], scan[:N] - scan[s]
            dst = np.where(flag[:N] == 1, s + lb, s + lt + (np.arange(N) - s) - lb)
            new = np.empty(N, dtype=np.int64)
            new[dst] = perm
            child = np.empty(2 * nodes + 1, dtype=np.int64)
            child[0:2 * nodes:2] = seg[:-1]
            child[1:2 * nodes:2] = seg[:-1] + (scan[seg[1:]] - scan[seg[:-1]])
            child[-1] = N
            perm, seg = new, child
        oroot, _, oinds = O.setuppartition(X, levels)
        ohv, ohc = O.fetchhyperplanes(oroot)
        assert np.array_equal(hv, ohv) and np.array_equal(hc, ohc)
        for p, b in enumerate(oinds):
            assert np.array_equal(perm[seg[p]:seg[p + 1]] + 1, b)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU restatement timed on the host cores) prints ONE JSON line with the keys the driver
    reads, on the product arm's metric / unit / config; small workload so that the CPU suite stays short.  Run the way torchrun
    runs it: OMP_NUM_THREADS=1 injected, WORLD_SIZE=2 -- rank 0 must still use every host core and cap the step count, rank 1
    must exit 0 without output."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="1", WORLD_SIZE="2", RANK="0", LOCAL_RANK="0")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "c2", "--steps", "20", "--warmup", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pts/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["value"] > 0 and d["fit_leaves_per_s"] > 0 and d["n_gpus"] == 2
    assert d["steps"] == 3 and d["steps_requested"] == 20 and d["warmup"] == 1          # bounded whatever was asked
    ncores = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == ncores   # not the injected OMP_NUM_THREADS=1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c2:")
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(env, RANK="1", LOCAL_RANK="1"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
