"""Host-side BSP partition (the L2 layer the reference keeps on the host:
src/patchwork/partition.jl).  Same public names and return values as the Julia functions;
the tree is held flattened (pre-order hyperplane arrays) because that is what crosses the C ABI
(pmk_set_tree).  Indices are 1-based like the reference's.

Numerics follow the reference operation by operation where a comparison depends on them:
  mean     : Base pairwise sum (sequential blocks of <= 1024 elements), then / n   partition.jl:89
  direction: v = V[:,1] of svd((array2matrix([X[1]-mu]))')  -- `size(X,2)` of a Vector is 1, so only
             X[1]-mu enters (partition.jl:90-94); LAPACK dgesdd through numpy, 'column' form
  split    : f_n = dot(v, X[n]) sequential, c = median(f) (a/2 + b/2 for even n), left iff f_n < c
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class BSPTree:
    """Flattened BinaryNode{PartitionDataType} tree (partition.jl:3-29).
    hps_v[k], hps_c[k]: hyperplane of the k-th internal node in PreOrderDFS order
    (== fetchhyperplanes order, mixtureGP.jl:322-334).  Complete tree: node k at depth d has its
    left child at k+1 and its right child at k + 2^(levels-2-d); leaf id = 1 + path bits."""
    levels: int
    hps_v: np.ndarray                      # (n_hp, D)
    hps_c: np.ndarray                      # (n_hp,)
    leaf_inds: List[np.ndarray] = field(default_factory=list)   # global_X_indices per leaf, 1-based

    @property
    def n_leaves(self) -> int:
        return 1 << (self.levels - 1)

    @property
    def D(self) -> int:
        return self.hps_v.shape[1]


def _dot_seq(v: np.ndarray, X: np.ndarray) -> np.ndarray:
    s = v[..., 0] * X[..., 0]
    for k in range(1, X.shape[-1]):
        s = s + v[..., k] * X[..., k]
    return s


def _pairwise_sum(X: np.ndarray) -> np.ndarray:
    """Base.mapreduce_impl(+) over the rows of X: ranges of <= 1024 rows are summed left to right,
    longer ranges are split at lo + (hi-lo)>>1."""
    stack = [(0, X.shape[0] - 1)]
    # post-order evaluation without recursion: collect leaf blocks in order, then combine pairwise
    def rec(lo, hi):
        if hi - lo < 1024:
            return np.add.accumulate(X[lo:hi + 1], axis=0)[-1]
        mid = lo + ((hi - lo) >> 1)
        return rec(lo, mid) + rec(mid + 1, hi)
    del stack
    return rec(0, X.shape[0] - 1)


def _median(f: np.ndarray) -> float:
    n = f.shape[0]
    mid = (1 + n) // 2
    if n & 1:
        return float(np.partition(f, mid - 1)[mid - 1])
    p = np.partition(f, [mid - 1, mid])
    return float(p[mid - 1] / 2.0 + p[mid] / 2.0)


def gethyperplane(X: np.ndarray, svd_form: str = "column"):
    """partition.jl:86-100 (+ splitpoints :64-83).  Returns (v, c), left_indicators."""
    mu = _pairwise_sum(X) / X.shape[0]
    z = X[0] - mu
    if svd_form == "column":      # Julia >= 1.7: svd of the Adjoint's D x 1 parent, U and V swapped
        U, _, _ = np.linalg.svd(z.reshape(-1, 1), full_matrices=False)
        v = np.ascontiguousarray(U[:, 0])
    else:                          # dgesdd on the materialised 1 x D matrix
        _, _, Vt = np.linalg.svd(z.reshape(1, -1), full_matrices=False)
        v = np.ascontiguousarray(Vt[0, :])
    f = _dot_seq(v[None, :], X)
    c = _median(f)
    return (v, c), f < c


def setuppartition(X, levels: int, svd_form: str = "column"):
    """setuppartition(X, levels) (partition.jl:106-129): returns root, X_parts, X_parts_inds.
    X: (N, D) array (or a list of D-vectors).  `root` is a BSPTree."""
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    if X.ndim == 1:
        X = X[:, None]
    if levels < 2:
        raise ValueError("levels must be larger than 1 (examples/mixGP.jl:108)")
    N, D = X.shape
    n_hp = (1 << (levels - 1)) - 1
    hv = np.empty((n_hp, D))
    hc = np.empty(n_hp)
    leaf_inds: List[np.ndarray] = [None] * (1 << (levels - 1))
    Lv = levels - 1
    # explicit stack instead of the reference's recursion: (pre-order node id, depth, leaf prefix, indices)
    stack = [(0, 0, 0, np.arange(N, dtype=np.int64))]
    while stack:
        k, d, prefix, idx = stack.pop()
        (v, c), left = gethyperplane(X[idx], svd_form)
        hv[k], hc[k] = v, c
        il, ir = idx[left], idx[~left]
        if d == Lv - 1:            # children are leaves (createchildren with level == 1, partition.jl:193)
            leaf_inds[prefix * 2] = il + 1
            leaf_inds[prefix * 2 + 1] = ir + 1
        else:
            stack.append((k + (1 << (Lv - 1 - d)), d + 1, prefix * 2 + 1, ir))
            stack.append((k + 1, d + 1, prefix * 2, il))
    root = BSPTree(levels=levels, hps_v=hv, hps_c=hc, leaf_inds=leaf_inds)
    X_parts = [X[i - 1] for i in leaf_inds]
    return root, X_parts, leaf_inds


def fetchhyperplanes(root: BSPTree):
    """fetchhyperplanes(root) (mixtureGP.jl:322-334): hyperplanes in PreOrderDFS order, as (v, c) arrays."""
    return root.hps_v, root.hps_c


def findpartition(x, root: BSPTree, levels: int | None = None):
    """findpartition (partition.jl:248-262) on the host, for one point or an (n, D) array; 1-based."""
    X = np.atleast_2d(np.asarray(x, dtype=np.float64))
    Lv = root.levels - 1
    node = np.zeros(X.shape[0], dtype=np.int64)
    leaf = np.zeros(X.shape[0], dtype=np.int64)
    for d in range(Lv):
        right = ~(_dot_seq(root.hps_v[node], X) < root.hps_c[node])
        leaf = leaf * 2 + right
        if d < Lv - 1:
            node = node + np.where(right, 1 << (Lv - 1 - d), 1)
    out = leaf + 1
    return int(out[0]) if np.ndim(x) == 1 else out


def organizetrainingsets(root: BSPTree, levels: int, X0, ε: float):
    """organizetrainingsets(root, levels, X0, ε) (partition.jl:301-357, findεpartitions! :269-298).
    Returns X_set, X_set_inds (ascending 1-based global ids per leaf), regions_list_set
    (per point: its leaves in left-to-right order), problematic_inds."""
    X0 = np.ascontiguousarray(np.asarray(X0, dtype=np.float64))
    if X0.ndim == 1:
        X0 = X0[:, None]
    Lv = root.levels - 1
    pt = np.arange(X0.shape[0], dtype=np.int64)
    node = np.zeros_like(pt)
    leaf = np.zeros_like(pt)
    for d in range(Lv):
        h = _dot_seq(root.hps_v[node], X0[pt])
        c = root.hps_c[node]
        gl = h < c + ε          # partition.jl:287
        gr = h > c - ε          # partition.jl:292
        pt = np.concatenate([pt[gl], pt[gr]])
        leaf = np.concatenate([leaf[gl] * 2, leaf[gr] * 2 + 1])
        node = np.concatenate([node[gl] + 1, node[gr] + (1 << (Lv - 1 - d))])
    o = np.lexsort((pt, leaf))
    pt_l, leaf_l = pt[o], leaf[o]
    cnt = np.bincount(leaf_l, minlength=root.n_leaves)
    off = np.concatenate([[0], np.cumsum(cnt)])
    X_set_inds = [pt_l[off[r]:off[r + 1]] + 1 for r in range(root.n_leaves)]
    X_set = [X0[i - 1] for i in X_set_inds]
    o2 = np.lexsort((leaf, pt))
    cnt2 = np.bincount(pt, minlength=X0.shape[0])
    off2 = np.concatenate([[0], np.cumsum(cnt2)])
    leaf_p = leaf[o2] + 1
    regions_list_set = _RaggedView(leaf_p, off2)
    return X_set, X_set_inds, regions_list_set, []


class _RaggedView:
    """List-of-lists view over CSR data (avoids materialising 10^6 small Python lists)."""

    def __init__(self, data, off):
        self.data, self.off = data, off

    def __len__(self):
        return len(self.off) - 1

    def __getitem__(self, i):
        return self.data[self.off[i]:self.off[i + 1]]


def organizetrainingsets_device(root: BSPTree, levels: int, X0, ε: float, handle=None):
    """organizetrainingsets on the GPU (pmk_organize_training_sets): same return values as organizetrainingsets,
    bit-identical index lists; the reference's host loop (one Vector allocation per point, partition.jl:330) is the
    slowest part of model setup at 10^6 points."""
    import ctypes as C
    from ._lib import Handle, lib, ptr
    X0 = np.ascontiguousarray(np.asarray(X0, dtype=np.float64))
    if X0.ndim == 1:
        X0 = X0[:, None]
    h = handle or Handle(0)
    L = lib()
    hv = np.ascontiguousarray(root.hps_v, dtype=np.float64)
    hc = np.ascontiguousarray(root.hps_c, dtype=np.float64)
    h.check(L.pmk_set_tree(h.raw, X0.shape[1], levels, ptr(hv), ptr(hc)))
    leaf_off = np.empty(root.n_leaves + 1, dtype=np.int64)
    total = C.c_int64(0)
    h.check(L.pmk_organize_training_sets(h.raw, X0.shape[0], ptr(X0), float(ε), ptr(leaf_off), C.byref(total)))
    inds = np.empty(total.value, dtype=np.int32)
    poff = np.empty(X0.shape[0] + 1, dtype=np.int64)
    pleaves = np.empty(total.value, dtype=np.int32)
    h.check(L.pmk_organize_fetch(h.raw, ptr(inds), ptr(poff), ptr(pleaves)))
    inds64 = inds.astype(np.int64)
    X_set_inds = [inds64[leaf_off[r]:leaf_off[r + 1]] for r in range(root.n_leaves)]
    X_set = [X0[i - 1] for i in X_set_inds]
    return X_set, X_set_inds, _RaggedView(pleaves.astype(np.int64), poff), []


def _split_direction(z: np.ndarray, svd_form: str) -> np.ndarray:
    """v = V[:,1] of svd((array2matrix([z]))') (partition.jl:90-94), by the host LAPACK exactly as gethyperplane does it."""
    if svd_form == "column":
        U, _, _ = np.linalg.svd(z.reshape(-1, 1), full_matrices=False)
        return np.ascontiguousarray(U[:, 0])
    _, _, Vt = np.linalg.svd(z.reshape(1, -1), full_matrices=False)
    return np.ascontiguousarray(Vt[0, :])


def preorder_index(depth: int, j: int, levels: int) -> int:
    """Pre-order (fetchhyperplanes) index of the node at `depth` with left-to-right index j in a complete tree."""
    Lv = levels - 1
    k = 0
    for i in range(depth):
        k += (1 << (Lv - 1 - i)) if (j >> (depth - 1 - i)) & 1 else 1
    return k


def setuppartition_device(X, levels: int, svd_form: str = "column", handle=None):
    """setuppartition(X, levels) (partition.jl:106-129) with the O(N) work of every level on the GPU
    (pmk_partition_begin / _level_z / _level_split / _fetch): node means in Base's pairwise order, projections, medians and
    the stable splits.  The 1 x D svd of each node stays on the host (LAPACK, as in gethyperplane), so root, X_parts and
    X_parts_inds are bit-identical to setuppartition's.  Same return values."""
    from ._lib import Handle, lib, ptr
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    if X.ndim == 1:
        X = X[:, None]
    if levels < 2:
        raise ValueError("levels must be larger than 1 (examples/mixGP.jl:108)")
    N, D = X.shape
    h = handle or Handle(0)
    L = lib()
    h.check(L.pmk_partition_begin(h.raw, D, N, ptr(X), levels))
    n_hp = (1 << (levels - 1)) - 1
    hv, hc = np.empty((n_hp, D)), np.empty(n_hp)
    for depth in range(levels - 1):
        nodes = 1 << depth
        z = np.empty((nodes, D))
        h.check(L.pmk_partition_level_z(h.raw, depth, ptr(z)))
        v = np.ascontiguousarray(np.stack([_split_direction(z[j], svd_form) for j in range(nodes)]))
        c = np.empty(nodes)
        h.check(L.pmk_partition_level_split(h.raw, depth, ptr(v), ptr(c)))
        ks = [preorder_index(depth, j, levels) for j in range(nodes)]
        hv[ks], hc[ks] = v, c
    n_leaves = 1 << (levels - 1)
    leaf_off = np.empty(n_leaves + 1, dtype=np.int64)
    inds = np.empty(N, dtype=np.int32)
    h.check(L.pmk_partition_fetch(h.raw, ptr(leaf_off), ptr(inds)))
    inds64 = inds.astype(np.int64)
    leaf_inds = [inds64[leaf_off[p]:leaf_off[p + 1]] for p in range(n_leaves)]
    root = BSPTree(levels=levels, hps_v=hv, hps_c=hc, leaf_inds=leaf_inds)
    return root, [X[i - 1] for i in leaf_inds], leaf_inds
