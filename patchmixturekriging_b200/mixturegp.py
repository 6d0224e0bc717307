"""Host mirror of the reference's mixture-GP layer (src/RKHS/mixtureGP.jl): same names, argument
order and return values; the bodies call the CUDA library through the C ABI.

  MixtureGPType(X_set, hps)                                 mixtureGP.jl:54-66
  fitmixtureGP_(η, y_parts, θ, σ²) -> η                     fitmixtureGP!          :70-118
  querymixtureGP_(Yq, Vq, Xq, η, root, levels, radius, δ, θ, σ², weight_θ, debug_vars; debug_flag)
                                                            querymixtureGP!        :159-294
  querymixtureGP(Xq | xq, η, root, levels, radius, δ, θ, σ², weight_θ; debug_flag) -> Yq, Vq, debug_vars
                                                            :120-157
  fetchhyperplanes(root)                                    :322-334 (in partition.py)
Julia's `!` suffix is spelled `_` here.  Vector{Vector{Float64}} inputs are (n, D) arrays (or lists
of them): a C-contiguous (n, D) array IS array2matrix(X) (utilities.jl:25-36) in column-major terms.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Handle, MultiHandle, PMKError, PosDefException, lib, ptr
from .partition import BSPTree


def _as_points(X) -> np.ndarray:
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    if X.ndim == 1:
        X = X[:, None]
    return X


@dataclass
class MixtureGPDebugType:
    """mixtureGP.jl:5-35: per-query debug outputs (filled when debug_flag=True)."""
    w_tilde_set: List[np.ndarray] = field(default_factory=list)
    u_set: List[np.ndarray] = field(default_factory=list)
    v_set: List[np.ndarray] = field(default_factory=list)
    region_inds_set: List[np.ndarray] = field(default_factory=list)
    p_region_ind_set: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int32))
    # hyperplane bookkeeping, compact: the kept hyperplane ids and their t per query ...
    kept_hp_set: List[np.ndarray] = field(default_factory=list)
    t_kept_set: List[np.ndarray] = field(default_factory=list)
    # ... and the reference's dense arrays over ALL hyperplanes (mixtureGP.jl:17-19,256-258), filled when
    # Nq * n_hp <= DENSE_DEBUG_LIMIT (4095 values per query in the 1M-point configuration): per query j,
    # hps_keep_flags_set[j] (bool, n_hp), zs_set[j] (n_hp x D), ts_set[j] (n_hp)
    hps_keep_flags_set: List[np.ndarray] = field(default_factory=list)
    zs_set: List[np.ndarray] = field(default_factory=list)
    ts_set: List[np.ndarray] = field(default_factory=list)
    # flat CSR form of the same data
    pair_off: Optional[np.ndarray] = None


DENSE_DEBUG_LIMIT = 1 << 26     # Nq * n_hp entries up to which debug_flag also fills the dense per-hyperplane arrays


class _LazyLeafList:
    """c_set / L_set / U_set of MixtureGPType: fetched from the GPU on indexing (0-based Python index)."""

    def __init__(self, eta: "MixtureGPType", what: str):
        self._eta, self._what = eta, what

    def __len__(self):
        return len(self._eta.X_parts)

    def __getitem__(self, i: int) -> np.ndarray:
        eta = self._eta
        if not eta._fitted:
            raise PMKError(_lib.PMK_ERR_STATE, "MixtureGPType is not fitted (Julia: UndefRefError on c_set[n])")
        n = eta.X_parts[i].shape[0]
        leaf = i + 1
        L = lib()
        h = eta._leaf_handle(i)          # the handle (rank) that owns the leaf; leaf ids stay global
        if self._what == "c":
            out = np.empty(n)
            h.check(L.pmk_get_alpha(h.raw, leaf, ptr(out)))
            return out
        out = np.empty((n, n), order="F")
        fn = {"L": L.pmk_get_L, "Linv": L.pmk_get_Linv}.get(self._what, L.pmk_get_K)
        h.check(fn(h.raw, leaf, ptr(out)))
        return out


class MixtureGPType:
    """MixtureGPType{T} (mixtureGP.jl:38-66).  Fields: X_parts, c_set, σ²_set, U_set, L_set, hps.
    The fitted state lives in HBM; c_set / L_set / U_set materialise a leaf on the host on demand
    (U_set is the Gram matrix WITHOUT σ², mixtureGP.jl:99)."""

    def __init__(self, X_parts: Sequence[np.ndarray], hps, device: int = 0, devices=None):
        """devices: a list of CUDA ordinals (or their number) shards the model over several GPUs of the box by sub-tree
        ownership (pmk_multi): rank r owns a contiguous range of the leaves; fit and query keep their signatures."""
        self.X_parts = [_as_points(X) for X in X_parts]
        self.hps = hps                     # (hps_v, hps_c) as returned by fetchhyperplanes
        self.σ2_set: List[float] = []
        self.c_set = _LazyLeafList(self, "c")
        self.L_set = _LazyLeafList(self, "L")
        self.U_set = _LazyLeafList(self, "K")
        self._multi = MultiHandle(devices) if devices is not None else None
        self._h = Handle(device) if self._multi is None else None
        self._fitted = False
        self._tree_key = None
        self.θ = None

    # -- convenience
    @property
    def handle(self) -> Handle:
        return self._h

    @property
    def multi(self) -> Optional[MultiHandle]:
        return self._multi

    def _leaf_handle(self, i: int) -> Handle:
        if self._multi is None:
            return self._h
        return self._multi.rank_handle(self._multi.owner_of_leaf(i, len(self.X_parts)))

    def close(self):
        if self._multi is not None:
            self._multi.close()
        if self._h is not None:
            self._h.close()


def fitmixtureGP_(η: MixtureGPType, y_parts: Sequence[np.ndarray], θ, σ2: float) -> MixtureGPType:
    """fitmixtureGP!(η, y_parts, θ, σ²) (mixtureGP.jl:70-118): per leaf, Gram + σ²I, weights c and the
    Cholesky factor L -- all leaves in one batched launch sequence on the GPU."""
    N_parts = len(η.X_parts)
    if len(y_parts) != N_parts:
        raise PMKError(_lib.PMK_ERR_ARG, "length(y_parts) != length(η.X_parts)")
    sizes = np.array([X.shape[0] for X in η.X_parts], dtype=np.int64)
    for n, y in zip(sizes, y_parts):
        if np.shape(y)[0] != n:
            raise PMKError(_lib.PMK_ERR_ARG, "DimensionMismatch: length(y) != length(X) in a leaf")
    D = η.X_parts[0].shape[1]
    leaf_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    X_packed = np.ascontiguousarray(np.concatenate(η.X_parts, axis=0))
    y_packed = np.ascontiguousarray(np.concatenate([np.asarray(y, dtype=np.float64) for y in y_parts]))
    L = lib()
    bad, info = C.c_int64(0), C.c_int(0)
    kp = θ.params
    η._fitted = False
    if η._multi is not None:
        rc = L.pmk_multi_fit(η._multi.raw, D, N_parts, ptr(leaf_off), ptr(X_packed), ptr(y_packed), θ.kernel_id, ptr(kp), kp.shape[0],
                             float(σ2), C.byref(bad), C.byref(info))
        if rc == _lib.PMK_ERR_NOT_POSDEF:
            raise PosDefException(info.value, bad.value, L.pmk_multi_last_error(η._multi.raw).decode())
        η._multi.check(rc)
    else:
        rc = L.pmk_fit(η._h.raw, D, N_parts, ptr(leaf_off), ptr(X_packed), ptr(y_packed), θ.kernel_id, ptr(kp), kp.shape[0],
                       float(σ2), C.byref(bad), C.byref(info))
        if rc == _lib.PMK_ERR_NOT_POSDEF:
            raise PosDefException(info.value, bad.value, L.pmk_last_error(η._h.raw).decode())
        η._h.check(rc)
    η._fitted = True
    η.θ = θ
    η.σ2_set = [float(σ2)] * N_parts
    return η


def savemixtureGP(η: MixtureGPType, path: str, root: Optional[BSPTree] = None, levels: Optional[int] = None) -> None:
    """Checkpoint of a fitted model (pmk_save_model): X_parts, c_set, L_set, kernel, σ² and the tree (`root`, or the one
    the last query used) in one flat file.  The reference has no serialisation; this is the §8f-4 row."""
    if not η._fitted:
        raise PMKError(_lib.PMK_ERR_STATE, "savemixtureGP before fitmixtureGP_")
    if η._multi is not None:
        raise PMKError(_lib.PMK_ERR_UNSUPPORTED, "savemixtureGP of a model sharded over several GPUs: a model file holds a whole model")
    if root is not None:
        _set_tree(η, root, root.levels if levels is None else levels)
    η._h.check(lib().pmk_save_model(η._h.raw, str(path).encode()))


def loadmixtureGP(path: str, device: int = 0):
    """Inverse of savemixtureGP: returns (η, root, levels).  η queries bit-identically to the model that was saved;
    root is the flattened tree stored in the file (None, levels 1 when the model was saved without one)."""
    from . import kernels as K
    η = MixtureGPType.__new__(MixtureGPType)
    η._multi = None
    η._h = Handle(device)
    L = lib()
    η._h.check(L.pmk_load_model(η._h.raw, str(path).encode()))
    D, nl, kid, lv = C.c_int(0), C.c_int64(0), C.c_int(0), C.c_int(0)
    kpar, s2 = C.c_double(0), C.c_double(0)
    η._h.check(L.pmk_model_info(η._h.raw, C.byref(D), C.byref(nl), C.byref(kid), C.byref(kpar), C.byref(s2), C.byref(lv)))
    η.X_parts = []
    for p in range(nl.value):
        n = C.c_int64(0)
        η._h.check(L.pmk_leaf_size(η._h.raw, p + 1, C.byref(n)))
        X = np.empty((n.value, D.value))
        η._h.check(L.pmk_get_X(η._h.raw, p + 1, ptr(X)))
        η.X_parts.append(X)
    by_id = {c.kernel_id: c for c in vars(K).values() if isinstance(c, type) and issubclass(c, K._Kernel) and c is not K._Kernel}
    η.θ = by_id[kid.value](kpar.value)
    η.σ2_set = [s2.value] * nl.value
    η.c_set = _LazyLeafList(η, "c")
    η.L_set = _LazyLeafList(η, "L")
    η.U_set = _LazyLeafList(η, "K")
    η._fitted = True
    η._tree_key = None
    root, levels = None, 1
    if lv.value > 1:
        n_hp = (1 << (lv.value - 1)) - 1
        hv, hc = np.empty((n_hp, D.value)), np.empty(n_hp)
        η._h.check(L.pmk_get_tree(η._h.raw, ptr(hv), ptr(hc)))
        root, levels = BSPTree(levels=lv.value, hps_v=hv, hps_c=hc), lv.value
        η._tree_key = _tree_fingerprint(root, levels)      # the handle already holds this tree
    η.hps = (root.hps_v, root.hps_c) if root is not None else (np.zeros((0, D.value)), np.zeros(0))
    return η, root, levels


def set_query_solver(η: MixtureGPType, solver: int):
    """_lib.SOLVER_AUTO (default: by the fit's conditioning estimate), _lib.SOLVER_INVERSE (s = inv(L) kq, inverse formed once
    per fit) or _lib.SOLVER_SUBSTITUTION (blocked forward substitution, closest to the reference's dtrsv)."""
    if η._multi is not None:
        η._multi.check(lib().pmk_multi_set_option(η._multi.raw, _lib.OPT_QUERY_SOLVER, solver))
    else:
        η._h.check(lib().pmk_set_option(η._h.raw, _lib.OPT_QUERY_SOLVER, solver))


def set_inverse_builder(η: MixtureGPType, builder: int):
    """How P = inv(L) is formed: 0 (default) recursive doubling on the packed tiles, 1 the substitution kernel on identity
    right-hand sides.  PMK_OPT_INVERSE_BUILDER of include/pmk.h."""
    η._h.check(lib().pmk_set_option(η._h.raw, _lib.OPT_INVERSE_BUILDER, builder))


def build_M(η: MixtureGPType):
    """Build the pair kernel's operand (P = inv(L), or M_IJ = L_IJ inv(L_JJ) for the substitution solver) now instead of in
    the first variance query (pmk_build_M)."""
    η._h.check(lib().pmk_build_M(η._h.raw))


def condition_estimate(η: MixtureGPType):
    """(lower bound of the worst leaf's cond(K + σ²I), query solver SOLVER_AUTO resolves to) of the fitted model."""
    cond, sv = C.c_double(0), C.c_int(0)
    hs = [η._h] if η._multi is None else [η._multi.rank_handle(r) for r in range(η._multi.size)]
    worst, solver = 0.0, 0
    for h in hs:
        h.check(lib().pmk_condition_estimate(h.raw, C.byref(cond), C.byref(sv)))
        worst, solver = max(worst, cond.value), max(solver, sv.value)
    return worst, solver


def _tree_fingerprint(root: BSPTree, levels: int):
    """Content key of the tree a handle holds: in-place edits of root.hps_v / hps_c and a recycled id() both change it."""
    import hashlib
    hv = np.ascontiguousarray(root.hps_v, dtype=np.float64)
    hc = np.ascontiguousarray(root.hps_c, dtype=np.float64)
    return (levels, hv.shape, hashlib.blake2b(hv.tobytes() + hc.tobytes(), digest_size=16).digest())


def _set_tree(η: MixtureGPType, root: Optional[BSPTree], levels: int):
    L = lib()
    D = η.X_parts[0].shape[1]

    def upload(lv, hv, hc):
        if η._multi is not None:
            η._multi.check(L.pmk_multi_set_tree(η._multi.raw, D, lv, ptr(hv), ptr(hc)))
        else:
            η._h.check(L.pmk_set_tree(η._h.raw, D, lv, ptr(hv), ptr(hc)))

    if root is None or levels == 1:
        key = ("none",)
        if η._tree_key != key:
            upload(1, None, None)
            η._tree_key = key
        return
    if levels != root.levels:
        raise PMKError(_lib.PMK_ERR_ARG, "levels does not match the tree")
    key = _tree_fingerprint(root, levels)       # content, not id(root): edits in place and recycled ids must re-upload
    if η._tree_key != key:
        hv = np.ascontiguousarray(root.hps_v, dtype=np.float64)
        hc = np.ascontiguousarray(root.hps_c, dtype=np.float64)
        upload(levels, hv, hc)
        η._tree_key = key


def _fetch_debug(η: MixtureGPType, Nq: int, debug_vars: MixtureGPDebugType):
    L = lib()
    npairs = C.c_int64(0)
    η._h.check(L.pmk_last_query_pairs(η._h.raw, C.byref(npairs)))
    P = npairs.value
    home = np.empty(Nq, dtype=np.int32)
    off = np.empty(Nq + 1, dtype=np.int64)
    leaf = np.empty(P, dtype=np.int32)
    hp = np.empty(P, dtype=np.int32)
    t, w, u, v = (np.empty(P) for _ in range(4))
    η._h.check(L.pmk_last_query_debug(η._h.raw, ptr(home), ptr(off), ptr(leaf), ptr(hp), ptr(t), ptr(w), ptr(u), ptr(v)))
    debug_vars.p_region_ind_set = home
    debug_vars.pair_off = off
    debug_vars.w_tilde_set = [w[off[j]:off[j + 1]] for j in range(Nq)]
    debug_vars.u_set = [u[off[j]:off[j + 1]] for j in range(Nq)]
    debug_vars.v_set = [v[off[j]:off[j + 1]] for j in range(Nq)]
    debug_vars.region_inds_set = [leaf[off[j]:off[j + 1] - 1] for j in range(Nq)]
    debug_vars.kept_hp_set = [hp[off[j]:off[j + 1] - 1] for j in range(Nq)]
    debug_vars.t_kept_set = [t[off[j]:off[j + 1] - 1] for j in range(Nq)]
    debug_vars._flat = dict(home=home, pair_off=off, pair_leaf=leaf, pair_hp=hp, pair_t=t, pair_w=w, pair_u=u, pair_v=v)
    # the reference's dense per-hyperplane arrays (mixtureGP.jl:256-258), when they are small enough to be wanted
    n_hp = int(np.asarray(η.hps[1]).shape[0]) if η.hps is not None else 0
    D = η.X_parts[0].shape[1]
    if n_hp > 0 and Nq * n_hp <= DENSE_DEBUG_LIMIT:
        keep = np.empty((Nq, n_hp), dtype=np.uint8)
        ts = np.empty((Nq, n_hp))
        zs = np.empty((Nq, n_hp, D))
        η._h.check(L.pmk_last_query_debug_dense(η._h.raw, 0, Nq, ptr(keep), ptr(ts), ptr(zs)))
        debug_vars.hps_keep_flags_set = list(keep.astype(bool))
        debug_vars.ts_set = list(ts)
        debug_vars.zs_set = list(zs)


def querymixtureGP_(Yq: np.ndarray, Vq: np.ndarray, Xq, η: MixtureGPType, root: Optional[BSPTree], levels: int,
                    radius: float, δ: float, θ, σ2, weight_θ, debug_vars: Optional[MixtureGPDebugType] = None, *,
                    debug_flag: bool = False) -> None:
    """querymixtureGP!(Yq, Vq, Xq, η, root, levels, radius, δ, θ, σ², weight_θ, debug_vars; debug_flag)
    (mixtureGP.jl:159-294).  Yq, Vq must be float64 arrays of length(Xq) (Julia resizes them in place;
    numpy arrays cannot be resized, so the caller allocates).  Returns None."""
    Xq = _as_points(Xq)
    Nq = Xq.shape[0]
    if Yq.shape != (Nq,) or Vq.shape != (Nq,) or Yq.dtype != np.float64 or Vq.dtype != np.float64:
        raise PMKError(_lib.PMK_ERR_ARG, "Yq / Vq must be float64 arrays of length(Xq)")
    if not η._fitted:
        raise PMKError(_lib.PMK_ERR_STATE, "query before fitmixtureGP_")
    if θ is not η.θ and (θ.kernel_id != η.θ.kernel_id or not np.array_equal(θ.params, η.θ.params)):
        raise PMKError(_lib.PMK_ERR_ARG, "θ differs from the kernel the model was fitted with")
    _set_tree(η, root, levels)
    wp = weight_θ.params
    if η._multi is not None:
        if debug_flag:
            raise PMKError(_lib.PMK_ERR_UNSUPPORTED, "debug_flag on a model sharded over several GPUs: query a single-GPU model for the debug outputs")
        η._multi.check(lib().pmk_multi_query(η._multi.raw, Nq, ptr(Xq), float(radius), float(δ), weight_θ.kernel_id, ptr(wp), wp.shape[0], 0,
                                             ptr(Yq), ptr(Vq)))
        return None
    η._h.check(lib().pmk_query(η._h.raw, Nq, ptr(Xq), float(radius), float(δ), weight_θ.kernel_id, ptr(wp), wp.shape[0], 0,
                               ptr(Yq), ptr(Vq)))
    if debug_flag and debug_vars is not None:
        _fetch_debug(η, Nq, debug_vars)
    return None


def querymixtureGP(Xq, η: MixtureGPType, root, levels, radius, δ, θ, σ2, weight_θ, *, debug_flag: bool = False):
    """querymixtureGP(xq or Xq, ...) -> (Yq, Vq, debug_vars)  (mixtureGP.jl:120-157)."""
    Xq = np.asarray(Xq, dtype=np.float64)
    D = η.X_parts[0].shape[1]
    if Xq.ndim == 1 and Xq.shape[0] == D and D > 1:
        Xq = Xq[None, :]                       # single query point
    Xq = _as_points(Xq)
    Yq = np.empty(Xq.shape[0])
    Vq = np.empty(Xq.shape[0])
    dv = MixtureGPDebugType()
    querymixtureGP_(Yq, Vq, Xq, η, root, levels, radius, δ, θ, σ2, weight_θ, dv, debug_flag=debug_flag)
    return Yq, Vq, dv
