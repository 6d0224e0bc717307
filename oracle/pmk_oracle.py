"""CPU ORACLE (test infrastructure, NOT product code) -- PARITY UNPINNED.

A numpy/scipy restatement of the reference's local-GP fit + mixture-query path
(RoyCCWang/PatchMixtureKriging, pure Julia).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module; the product package ``patchmixturekriging_b200`` never does.

"Parity unpinned": the reference cannot run here (no Julia in the image) and its own
test-suite is empty (test/runtests.jl:4-6), so there are no golden vectors to pin this
restatement against.  What pins it instead (tests/test_oracle.py):
  * closed-form kernel known-answers (kernel.jl:156-158, :218-225, :299-313, :350-357),
  * the inline invariants the reference scripts assert
    (examples/patchGP_partitioning.jl:198,214-215; src/patchwork/partition.jl:151-153;
     dev/btree_easy.jl:63-69 traversal orders),
  * the same LAPACK routines Julia's LinearAlgebra dispatches to (dgesdd, dgetrf/dgetrs,
    dpotrf, dtrsv) reached through scipy.

Arithmetic that lives outside /root/reference (un-vendored Julia stdlib, Project.toml:13
julia = "1.7"; AbstractTrees = "0.3", Project.toml:12) is restated from its published
algorithm: pairwise ``sum`` with 1024-element sequential blocks, ``median`` = middle order
statistic(s) with ``a/2 + b/2``, ``generic_norm2`` = sequential sum of squares then sqrt,
``dot`` = sequential multiply-add WITHOUT fused contraction (the contract the device
code matches bit-for-bit wherever a comparison depends on it).

All indices returned are 1-based, exactly as the Julia code returns them.
Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg
import scipy.linalg.blas
import scipy.linalg.lapack

# ----------------------------------------------------------------------------------------
# kernels  (src/RKHS/kernel.jl, parameter structs src/misc/declarations.jl)
# ----------------------------------------------------------------------------------------
SQEXP, SPLINE34, BB10, BB20, BB1EPS, BB2EPS, SPLINE12, SPLINE32, RQ = range(9)
STATIONARY = (SQEXP, SPLINE34, SPLINE12, SPLINE32, RQ)
KERNEL_NAMES = {SQEXP: "GaussianKernel1DType", SPLINE34: "Spline34KernelType", BB10: "BrownianBridge10",
                BB20: "BrownianBridge20", BB1EPS: "BrownianBridge1ϵ", BB2EPS: "BrownianBridge2ϵ",
                SPLINE12: "Spline12KernelType", SPLINE32: "Spline32KernelType",
                RQ: "RationalQuadraticKernelType"}


@dataclass(frozen=True)
class Kernel:
    """kind + its single scalar parameter (ϵ_sq / a / ϵ), declarations.jl:25-100."""
    kind: int
    param: float = 1.0


def evalkernel_tau(tau, th: Kernel):
    """Stationary kernels on a distance τ (scalar or ndarray).  kernel.jl:299-374."""
    tau = np.asarray(tau, dtype=np.float64)
    a = np.float64(th.param)
    if th.kind == SQEXP:                       # kernel.jl:350-357  exp(-ϵ_sq*τ^2)
        return np.exp((-a) * (tau * tau))
    if th.kind in (SPLINE34, SPLINE12, SPLINE32):
        r = tau * a                            # kernel.jl:301
        t = 1.0 - r
        neg = np.sign(t) < 0.0                 # kernel.jl:304 (t == 0 is kept -> value 0 anyway)
        tc = np.where(neg, 0.0, t)
        if th.kind == SPLINE34:                # kernel.jl:308-309
            out = ((35.0 * (r * r) + 18.0 * r + 3.0) * tc ** 6) / 3.0
        elif th.kind == SPLINE12:              # kernel.jl:325-326
            out = (3.0 * r + 1.0) * tc ** 3
        else:                                  # kernel.jl:342-343
            out = (4.0 * r + 1.0) * tc ** 4
        return np.where(neg, 0.0, out)
    if th.kind == RQ:                          # kernel.jl:360-366
        return np.sqrt(a) ** 3 / np.sqrt(a + tau * tau) ** 3
    raise ValueError("not a stationary kernel")


def _bb_1d(x, z, th: Kernel):
    """One-dimensional Brownian-bridge kernels on [0,1].  kernel.jl:156-225."""
    x = np.asarray(x, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    if th.kind == BB10:                        # kernel.jl:156-158
        return np.minimum(x, z) - x * z
    if th.kind == BB20:                        # kernel.jl:218-225 (left-to-right products)
        a = (((-1.0 / 6.0) * z) * (1.0 - x)) * ((x * x + z * z) - 2.0 * x)
        b = (((-1.0 / 6.0) * x) * (1.0 - z)) * ((x * x + z * z) - 2.0 * z)
        return np.where(z < x, a, b)
    e = np.float64(th.param)
    if th.kind == BB1EPS:                      # kernel.jl:168-174
        den = e * np.sinh(e)
        num = np.sinh(e * np.minimum(x, z)) * np.sinh(e * (1.0 - np.maximum(x, z)))
        return num / den
    if th.kind == BB2EPS:                      # kernel.jl:176-193
        mn, mx, ad = np.minimum(x, z), np.maximum(x, z), np.abs(x - z)
        s = x + z
        mult = np.exp(-e * s) / (4.0 * e ** 3 * (np.exp(2.0 * e) - 1.0) ** 2)
        t1 = np.exp(2.0 * e) * (2.0 * e - e * s - 1.0)
        t2 = np.exp(4.0 * e) * (e * s + 1.0)
        t3 = np.exp(2.0 * e * (1.0 + s)) * (2.0 * e - e * s + 1.0)
        t4 = np.exp(2.0 * e * s) * (e * s - 1.0)
        t5 = np.exp(2.0 * e * (2.0 + mn)) * (-e * ad - 1.0)
        t6 = np.exp(2.0 * e * mx) * (-e * ad + 1.0)
        t7 = np.exp(2.0 * e * (1.0 + mn)) * (1.0 - 2.0 * e + e * ad)
        t8 = np.exp(2.0 * e * (1.0 + mx)) * (1.0 + 2.0 * e - e * ad)
        return mult * (((((((t1 + t2) + t3) + t4) + t5) + t6) + t7) + t8)
    raise ValueError("not a Brownian-bridge kernel")


def norm2_seq(d: np.ndarray) -> np.ndarray:
    """Julia generic_norm2 (vectors shorter than 32): sequential Σd² then sqrt.  d: (..., D)."""
    s = d[..., 0] * d[..., 0]
    for k in range(1, d.shape[-1]):
        s = s + d[..., k] * d[..., k]
    return np.sqrt(s)


def dot_seq(v: np.ndarray, x: np.ndarray) -> np.ndarray:
    """dot(v, x) over the last axis: sequential, every product and sum rounded (no FMA)."""
    s = v[..., 0] * x[..., 0]
    for k in range(1, x.shape[-1]):
        s = s + v[..., k] * x[..., k]
    return s


def kernel_cross(X: np.ndarray, Z: np.ndarray, th: Kernel) -> np.ndarray:
    """K[i,j] = evalkernel(X[i], Z[j], θ); X:(n,D) Z:(m,D).  kernel.jl:277-287 (τ=norm(x1-x2)),
    :196-198 (tensor product over dimensions)."""
    X = np.asarray(X, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    if th.kind in STATIONARY:
        d = X[:, None, :] - Z[None, :, :]
        return evalkernel_tau(norm2_seq(d), th)
    out = _bb_1d(X[:, None, 0], Z[None, :, 0], th)
    for k in range(1, X.shape[1]):
        out = out * _bb_1d(X[:, None, k], Z[None, :, k], th)
    return out


def evalkernel(x, z, th: Kernel) -> float:
    """Scalar evalkernel(x, z, θ) for two points (1-D arrays)."""
    return float(kernel_cross(np.atleast_2d(np.asarray(x, float)), np.atleast_2d(np.asarray(z, float)), th)[0, 0])


def constructkernelmatrix(X: np.ndarray, th: Kernel) -> np.ndarray:
    """RKHS.jl:4-34: lower triangle evaluated, then mirrored -> exactly symmetric."""
    K = kernel_cross(X, X, th)
    il = np.tril_indices(K.shape[0], -1)
    K[il[1], il[0]] = K[il]
    return K


# ----------------------------------------------------------------------------------------
# BSP partition  (src/patchwork/partition.jl)
# ----------------------------------------------------------------------------------------
def mean_pairwise(X: np.ndarray) -> np.ndarray:
    """Statistics.mean of a Vector{Vector}: Base pairwise sum (sequential blocks of < 1024,
    split at ifirst + (ilast-ifirst)>>1), then / n.  partition.jl:89."""
    def rec(lo: int, hi: int) -> np.ndarray:          # inclusive 0-based range
        if hi - lo < 1024:
            s = X[lo].copy()
            for i in range(lo + 1, hi + 1):
                s = s + X[i]
            return s
        mid = lo + ((hi - lo) >> 1)
        return rec(lo, mid) + rec(mid + 1, hi)
    n = X.shape[0]
    if n - 1 < 1024 and n > 64:
        # same order, vectorised: sequential accumulation == cumulative sum, per dimension
        s = np.add.accumulate(X, axis=0)[-1]
    else:
        s = rec(0, n - 1)
    return s / n


def median_julia(f: np.ndarray) -> float:
    """Statistics.median: odd n -> middle element; even n -> a/2 + b/2.  partition.jl:70."""
    n = f.shape[0]
    mid = (1 + n) // 2                                # div(first+last, 2), 1-based
    if n % 2 == 1:
        return float(np.partition(f, mid - 1)[mid - 1])
    p = np.partition(f, [mid - 1, mid])
    return float(p[mid - 1] / 2.0 + p[mid] / 2.0)


def split_direction(z: np.ndarray, svd_form: str = "column") -> np.ndarray:
    """v = V[:,1] of svd((array2matrix([X[1]-μ]))').  partition.jl:90-94.
    `size(X,2)` of a Vector is 1, so only X[1]-μ enters.  Julia >= 1.7 computes the svd of an
    Adjoint from its D×1 parent and swaps U/V ('column' form, the default); the 'row' form is
    LAPACK dgesdd on the materialised 1×D matrix (old Julia)."""
    z = np.asarray(z, dtype=np.float64)
    if svd_form == "column":
        U, _, _ = np.linalg.svd(z.reshape(-1, 1), full_matrices=False)
        return np.ascontiguousarray(U[:, 0])
    _, _, Vt = np.linalg.svd(z.reshape(1, -1), full_matrices=False)
    return np.ascontiguousarray(Vt[0, :])


@dataclass
class Node:
    """BinaryNode{PartitionDataType} (partition.jl:9-26)."""
    v: Optional[np.ndarray] = None        # hp.v (None at leaves: HyperplaneType() undefined)
    c: float = 0.0
    left: Optional["Node"] = None
    right: Optional["Node"] = None
    inds: Optional[np.ndarray] = None     # global_X_indices (1-based), kept at leaves only
    index: int = 0                        # leaf label, 1-based


def gethyperplane(X: np.ndarray, svd_form: str = "column"):
    """partition.jl:86-100 + splitpoints :64-83."""
    mu = mean_pairwise(X)
    v = split_direction(X[0] - mu, svd_form)
    f = dot_seq(v[None, :], X)
    c = median_julia(f)
    return v, c, f < c


def _createchildren(parent: Node, left_ind: np.ndarray, X: np.ndarray, inds: np.ndarray, level: int, svd_form: str):
    """partition.jl:166-217 (both directions)."""
    for side in ("left", "right"):
        m = left_ind if side == "left" else ~left_ind
        Xk, ik = X[m], inds[m]
        kid = Node(inds=ik)
        setattr(parent, side, kid)
        if level == 1:
            continue
        kid.inds = None
        kid.v, kid.c, li = gethyperplane(Xk, svd_form)
        _createchildren(kid, li, Xk, ik, level - 1, svd_form)


def leaves(node: Node):
    """AbstractTrees.Leaves order: left-first DFS (dev/btree_easy.jl:63-69)."""
    if node.left is None and node.right is None:
        yield node
        return
    if node.left is not None:
        yield from leaves(node.left)
    if node.right is not None:
        yield from leaves(node.right)


def preorder(node: Node):
    """AbstractTrees.PreOrderDFS."""
    yield node
    if node.left is not None:
        yield from preorder(node.left)
    if node.right is not None:
        yield from preorder(node.right)


def setuppartition(X: np.ndarray, levels: int, svd_form: str = "column"):
    """partition.jl:106-129 (+labelleafnodes :131-159).  X: (N, D).
    Returns root, X_parts (list of (n_p, D) arrays), X_parts_inds (list of 1-based int64 arrays)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    root = Node()
    root.v, root.c, li = gethyperplane(X, svd_form)
    _createchildren(root, li, X, np.arange(1, X.shape[0] + 1, dtype=np.int64), levels - 1, svd_form)
    X_parts, X_parts_inds = [], []
    for i, leaf in enumerate(leaves(root), start=1):
        leaf.index = i
        X_parts.append(X[leaf.inds - 1])
        X_parts_inds.append(leaf.inds)
    return root, X_parts, X_parts_inds


def fetchhyperplanes(root: Node):
    """mixtureGP.jl:322-334: hyperplanes of the internal nodes in PreOrderDFS order."""
    hv, hc = [], []
    for node in preorder(root):
        if node.v is not None:
            hv.append(node.v)
            hc.append(node.c)
    return np.array(hv, dtype=np.float64).reshape(len(hc), -1), np.array(hc, dtype=np.float64)


def findpartition(x: np.ndarray, root: Node, levels: int) -> int:
    """partition.jl:248-262."""
    node = root
    for _ in range(levels - 1):
        node = node.left if dot_seq(node.v, x) < node.c else node.right
    return node.index


def find_eps_partitions(out: List[int], x: np.ndarray, node: Node, eps: float) -> None:
    """partition.jl:269-298."""
    if node.left is None and node.right is None:
        out.append(node.index)
        return
    h = dot_seq(node.v, x)
    if h < node.c + eps:
        find_eps_partitions(out, x, node.left, eps)
    if h > node.c - eps:
        find_eps_partitions(out, x, node.right, eps)


def organizetrainingsets(root: Node, levels: int, X0: np.ndarray, eps: float):
    """partition.jl:301-357 (scalar loop; use organizetrainingsets_vec for big N)."""
    n_regions = sum(1 for _ in leaves(root))
    X_set_inds: List[List[int]] = [[] for _ in range(n_regions)]
    regions_list_set = []
    for n in range(X0.shape[0]):
        lst: List[int] = []
        find_eps_partitions(lst, X0[n], root, eps)
        for r in lst:
            X_set_inds[r - 1].append(n + 1)
        regions_list_set.append(lst)
    X_set_inds_a = [np.array(v, dtype=np.int64) for v in X_set_inds]
    X_set = [X0[v - 1] for v in X_set_inds_a]
    return X_set, X_set_inds_a, regions_list_set, []


def flatten_tree(root: Node, levels: int):
    """Pre-order hyperplane arrays + per-internal-node child links, for the vectorised routines.
    Node k (pre-order index among internal nodes) at depth d; its left child is k+1, its right
    child is k + 2^(levels-1-d) (complete tree)."""
    hv, hc = fetchhyperplanes(root)
    return hv, hc


def _descend_vec(X: np.ndarray, hv: np.ndarray, hc: np.ndarray, levels: int) -> np.ndarray:
    """Vectorised findpartition for many points; returns 1-based leaf ids."""
    L = levels - 1
    node = np.zeros(X.shape[0], dtype=np.int64)
    leaf = np.zeros(X.shape[0], dtype=np.int64)
    for d in range(L):
        h = dot_seq(hv[node], X)
        right = ~(h < hc[node])
        leaf = leaf * 2 + right
        if d < L - 1:
            node = node + 1 + right * ((1 << (L - 1 - d)) - 1)
    return leaf + 1


def organizetrainingsets_vec(hv: np.ndarray, hc: np.ndarray, levels: int, X0: np.ndarray, eps: float):
    """Same result as organizetrainingsets, level-by-level frontier expansion (for N ~ 1e6)."""
    L = levels - 1
    pt = np.arange(X0.shape[0], dtype=np.int64)
    node = np.zeros_like(pt)
    leaf = np.zeros_like(pt)
    for d in range(L):
        h = dot_seq(hv[node], X0[pt])
        go_l = h < hc[node] + eps
        go_r = h > hc[node] - eps
        step = (1 << (L - 1 - d)) - 1
        pt = np.concatenate([pt[go_l], pt[go_r]])
        leaf = np.concatenate([leaf[go_l] * 2, leaf[go_r] * 2 + 1])
        node = np.concatenate([node[go_l] + 1, node[go_r] + 1 + step])
    order = np.lexsort((pt, leaf))
    pt, leaf = pt[order], leaf[order]
    n_leaves = 1 << L
    counts = np.bincount(leaf, minlength=n_leaves)
    off = np.concatenate([[0], np.cumsum(counts)])
    return [pt[off[r]:off[r + 1]] + 1 for r in range(n_leaves)]


# ----------------------------------------------------------------------------------------
# fit  (src/RKHS/mixtureGP.jl:70-118, src/RKHS/RKHS.jl:182-217)
# ----------------------------------------------------------------------------------------
class PosDefException(Exception):
    def __init__(self, info: int, leaf: int = 0):
        super().__init__(f"matrix is not positive definite; Cholesky factorization failed (info={info}, leaf={leaf})")
        self.info, self.leaf = info, leaf


def backslash(U: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Julia `U\\y` for a dense square matrix: LU with partial pivoting (dgetrf + dgetrs)."""
    lu, piv = scipy.linalg.lu_factor(U, check_finite=False)
    return scipy.linalg.lu_solve((lu, piv), y, check_finite=False)


def cholesky_L(U: np.ndarray, leaf: int = 0) -> np.ndarray:
    """cholesky(U).L: LAPACK dpotrf('U') then L = Rᵀ.  mixtureGP.jl:109-112."""
    R, info = scipy.linalg.lapack.dpotrf(np.asfortranarray(U), lower=0, clean=1)
    if info != 0:
        raise PosDefException(int(info), leaf)
    return np.ascontiguousarray(R.T)


@dataclass
class MixtureGP:
    """MixtureGPType (mixtureGP.jl:38-66)."""
    X_parts: List[np.ndarray]
    hps_v: np.ndarray
    hps_c: np.ndarray
    c_set: List[np.ndarray] = field(default_factory=list)
    L_set: List[np.ndarray] = field(default_factory=list)
    U_set: List[np.ndarray] = field(default_factory=list)
    sigma2_set: List[float] = field(default_factory=list)


def fitmixtureGP(eta: MixtureGP, y_parts: Sequence[np.ndarray], th: Kernel, sigma2: float, keep_U: bool = True):
    """mixtureGP.jl:70-118."""
    eta.c_set, eta.L_set, eta.U_set, eta.sigma2_set = [], [], [], []
    for n, (X, y) in enumerate(zip(eta.X_parts, y_parts), start=1):
        K = constructkernelmatrix(X, th)
        if keep_U:
            eta.U_set.append(K.copy())
        U = K
        U[np.diag_indices_from(U)] += sigma2
        c = backslash(U, np.asarray(y, dtype=np.float64))
        L = cholesky_L(U, n)
        eta.L_set.append(L)
        eta.c_set.append(c)
        eta.sigma2_set.append(sigma2)
    return eta


def fitRKHS(X: np.ndarray, y: np.ndarray, th: Kernel, sigma2: float) -> np.ndarray:
    """RKHS.jl:195-217: c = (K + σ²I) \\ y."""
    U = constructkernelmatrix(X, th)
    U[np.diag_indices_from(U)] += sigma2
    return backslash(U, np.asarray(y, dtype=np.float64))


def query_rkhs(Xq: np.ndarray, X: np.ndarray, c: np.ndarray, th: Kernel) -> np.ndarray:
    """RKHS.jl:220-247: Yq = dot(kq, c), mean only."""
    return kernel_cross(Xq, X, th) @ c


# ----------------------------------------------------------------------------------------
# query  (src/RKHS/mixtureGP.jl:159-316, :339-405)
# ----------------------------------------------------------------------------------------
def findneighbourpartitions(p: np.ndarray, radius: float, root: Node, levels: int,
                            hv: np.ndarray, hc: np.ndarray, home: int, delta: float = 1e-10):
    """mixtureGP.jl:339-405.  Returns region_inds, ts, zs, keep_flags."""
    H = hc.shape[0]
    region, keep = [], np.zeros(H, dtype=bool)
    ts = np.empty(H)
    zs = np.empty((H, p.shape[0]))
    for i in range(H):
        u, c = hv[i], hc[i]
        t = -dot_seq(u, p) + c
        z = p + t * u
        zs[i], ts[i] = z, t
        if norm2_seq(z - p) < radius:
            r1 = findpartition(p + (t + delta) * u, root, levels)
            r2 = findpartition(p + (t - delta) * u, root, levels)
            if (r2 == home) != (r1 == home):
                keep[i] = True
                region.append(r2 if r1 == home else r1)
    return region, ts, zs, keep


def queryinner(xq: np.ndarray, X: np.ndarray, th: Kernel, c: np.ndarray, L: np.ndarray, min_v: float = 1e-12):
    """mixtureGP.jl:296-316."""
    kq = kernel_cross(xq[None, :], X, th)[0]
    mu = float(np.dot(kq, c))
    v = scipy.linalg.blas.dtrsv(L, kq, lower=1)
    kxx = evalkernel(xq, xq, th)
    return mu, float(min(max(kxx - float(np.dot(v, v)), min_v), np.inf))


def querymixtureGP(Xq: np.ndarray, eta: MixtureGP, root: Node, levels: int, radius: float, delta: float,
                   th: Kernel, sigma2: float, weight_th: Kernel, debug: bool = False):
    """mixtureGP.jl:159-294, scalar loop over queries.  Returns Yq, Vq, debug dict."""
    Nq = Xq.shape[0]
    Yq, Vq = np.empty(Nq), np.empty(Nq)
    dbg = dict(w_tilde=[], u=[], v=[], region_inds=[], p_region_ind=[], keep=[], ts=[], zs=[])
    for j in range(Nq):
        xq = Xq[j]
        home = findpartition(xq, root, levels)
        region, ts, zs, keep = findneighbourpartitions(xq, radius, root, levels, eta.hps_v, eta.hps_c, home, delta)
        t_kept = ts[keep]
        nr = len(region)
        w, u, v = np.empty(nr + 1), np.empty(nr + 1), np.empty(nr + 1)
        for i, r in enumerate(region):
            w[i] = float(evalkernel_tau(abs(t_kept[i]), weight_th))
            u[i], v[i] = queryinner(xq, eta.X_parts[r - 1], th, eta.c_set[r - 1], eta.L_set[r - 1])
        w[nr] = 1.0
        u[nr], v[nr] = queryinner(xq, eta.X_parts[home - 1], th, eta.c_set[home - 1], eta.L_set[home - 1])
        if debug:
            dbg["w_tilde"].append(w.copy()); dbg["u"].append(u.copy()); dbg["v"].append(v.copy())
            dbg["region_inds"].append(list(region)); dbg["p_region_ind"].append(home)
            dbg["keep"].append(keep); dbg["ts"].append(ts); dbg["zs"].append(zs)
        sw = 0.0
        for i in range(nr + 1):
            sw = sw + w[i]
        w = w / sw
        y = 0.0
        vv = 0.0
        for i in range(nr + 1):
            y = y + w[i] * u[i]
            vv = vv + w[i] * (v[i] * w[i])
        Yq[j], Vq[j] = y, vv
    return Yq, Vq, dbg


def query_structure_vec(Xq: np.ndarray, hv: np.ndarray, hc: np.ndarray, levels: int, radius: float, delta: float):
    """Vectorised home leaf + neighbour list for many queries (same arithmetic, same order as
    findneighbourpartitions).  Returns home (Nq,), and flat pair arrays sorted by (query, hp index):
    pair_q (0-based query), pair_hp (0-based hp), pair_leaf (1-based), pair_t."""
    home = _descend_vec(Xq, hv, hc, levels)
    pq, ph, pl, pt = [], [], [], []
    for i in range(hc.shape[0]):
        u, c = hv[i], hc[i]
        t = -dot_seq(u[None, :], Xq) + c
        z = Xq + t[:, None] * u[None, :]
        near = np.nonzero(norm2_seq(z - Xq) < radius)[0]
        if near.size == 0:
            continue
        P, tn = Xq[near], t[near]
        r1 = _descend_vec(P + (tn + delta)[:, None] * u[None, :], hv, hc, levels)
        r2 = _descend_vec(P + (tn - delta)[:, None] * u[None, :], hv, hc, levels)
        h = home[near]
        k = (r2 == h) != (r1 == h)
        pq.append(near[k]); ph.append(np.full(int(k.sum()), i)); pl.append(np.where(r1 == h, r2, r1)[k]); pt.append(tn[k])
    if pq:
        pq, ph, pl, pt = map(np.concatenate, (pq, ph, pl, pt))
        o = np.lexsort((ph, pq))
        return home, pq[o], ph[o], pl[o], pt[o]
    z = np.zeros(0, dtype=np.int64)
    return home, z, z, z, np.zeros(0)


def querymixtureGP_vec(Xq: np.ndarray, eta: MixtureGP, levels: int, radius: float, delta: float,
                       th: Kernel, weight_th: Kernel, min_v: float = 1e-12):
    """Vectorised querymixtureGP for large Nq: per leaf, one batched kernel_cross, a dtrsm with
    many right-hand sides (same substitution as dtrsv, batched), then the reference's
    normalise-and-combine in the reference's order (neighbours in hp order, home last)."""
    hv, hc = eta.hps_v, eta.hps_c
    home, pq, ph, pl, pt = query_structure_vec(Xq, hv, hc, levels, radius, delta)
    Nq = Xq.shape[0]
    # pair list: neighbours then home, per query
    allq = np.concatenate([pq, np.arange(Nq)])
    alll = np.concatenate([pl, home])
    allw = np.concatenate([evalkernel_tau(np.abs(pt), weight_th), np.ones(Nq)])
    slot = np.concatenate([np.zeros(pq.shape[0], dtype=np.int64), np.ones(Nq, dtype=np.int64)])  # home last
    o = np.lexsort((np.concatenate([ph, np.zeros(Nq, dtype=np.int64)]), slot, allq))
    allq, alll, allw = allq[o], alll[o], allw[o]
    U_, V_ = np.empty(allq.shape[0]), np.empty(allq.shape[0])
    for r in np.unique(alll):
        m = np.nonzero(alll == r)[0]
        X, c, L = eta.X_parts[r - 1], eta.c_set[r - 1], eta.L_set[r - 1]
        for s in range(0, m.shape[0], 4096):
            mm = m[s:s + 4096]
            xq = Xq[allq[mm]]
            Kq = kernel_cross(X, xq, th)                       # (n_p, m)
            U_[mm] = c @ Kq
            S = scipy.linalg.solve_triangular(L, Kq, lower=True, check_finite=False)
            if th.kind in STATIONARY:
                kxx = float(evalkernel_tau(0.0, th))
            else:
                kxx = np.array([evalkernel(x, x, th) for x in xq])
            V_[mm] = np.minimum(np.maximum(kxx - np.einsum("ij,ij->j", S, S), min_v), np.inf)
    counts = np.bincount(allq, minlength=Nq)
    off = np.concatenate([[0], np.cumsum(counts)])
    Yq, Vq = np.zeros(Nq), np.zeros(Nq)
    maxc = int(counts.max())
    # sequential sums in slot order, vectorised across queries
    sw = np.zeros(Nq)
    for k in range(maxc):
        has = counts > k
        idx = off[:-1][has] + k
        sw[has] = sw[has] + allw[idx]
    for k in range(maxc):
        has = counts > k
        idx = off[:-1][has] + k
        w = allw[idx] / sw[has]
        Yq[has] = Yq[has] + w * U_[idx]
        Vq[has] = Vq[has] + w * (V_[idx] * w)
    return Yq, Vq, dict(home=home, pair_off=off, pair_leaf=alll, pair_w=allw, pair_u=U_, pair_v=V_)
