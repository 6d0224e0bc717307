"""Host mirror of the reference's single-GP layer (src/RKHS/RKHS.jl, kernel.jl): same names;
bodies go through the C ABI to the same CUDA kernels as the mixture path (one leaf, no tree).

  evalkernel(x, z, θ)                        kernel.jl:277-295 (+ the per-kernel methods)
  constructkernelmatrix(X, θ) / (X, Z, θ)    RKHS.jl:4-34 / :95-110
  RKHSProblemType(c, X, θ, σ²)               declarations.jl:226-231
  fitRKHS_(η, y)                             fitRKHS!   RKHS.jl:182-217
  query_(Yq, Xq, η)                          query!     RKHS.jl:220-247   (mean only)
  evalquery(x, c, X, θ)                      querying.jl:2-5   Σ c_n k(x, X_n)
  setupGPquery(c, X, θ, σ²) -> fq            querying.jl:43-58; fq(xq) -> (mean, variance) = evalqueryGP! :60-79
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib
from ._lib import Handle, PMKError, PosDefException, lib, ptr

_shared_handle: Optional[Handle] = None


def _handle(device: int = 0) -> Handle:
    global _shared_handle
    if _shared_handle is None:
        _shared_handle = Handle(device)
    return _shared_handle


def _as_points(X) -> np.ndarray:
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    if X.ndim == 1:
        X = X[:, None]
    return X


def constructkernelmatrix(X, θ, Z=None, *, σ2: float = 0.0, fast_exp: bool = False) -> np.ndarray:
    """constructkernelmatrix(X, θ) -> n x n Gram (RKHS.jl:4-34); constructkernelmatrix(X, Z, θ) -> K_XZ
    (RKHS.jl:95-110), here spelled constructkernelmatrix(X, θ, Z).  fast_exp (PMK_OPT_GRAM_FAST_EXP): the squared exponential
    by the table-driven exp of the fit / query kernels, <= 2 ulp from the reference's sqrt / re-square / exp order."""
    X = _as_points(X)
    h = _handle()
    h.check(lib().pmk_set_option(h.raw, _lib.OPT_GRAM_FAST_EXP, 1 if fast_exp else 0))
    kp = θ.params
    if Z is None:
        K = np.empty((X.shape[0], X.shape[0]), order="F")
        h.check(lib().pmk_gram(h.raw, X.shape[1], X.shape[0], ptr(X), θ.kernel_id, ptr(kp), kp.shape[0], float(σ2), ptr(K)))
        return K
    Z = _as_points(Z)
    if Z.shape[1] != X.shape[1]:
        raise PMKError(_lib.PMK_ERR_ARG, "DimensionMismatch")
    K = np.empty((X.shape[0], Z.shape[0]), order="F")
    h.check(lib().pmk_cross_gram(h.raw, X.shape[1], X.shape[0], ptr(X), Z.shape[0], ptr(Z), θ.kernel_id, ptr(kp), kp.shape[0], ptr(K)))
    return K


def evalkernel(x, z, θ) -> float:
    """evalkernel(x, z, θ) for two points (kernel.jl:277-295): a 1 x 1 cross-Gram on the GPU."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))[None, :]
    z = np.atleast_1d(np.asarray(z, dtype=np.float64))[None, :]
    return float(constructkernelmatrix(x, θ, z)[0, 0])


@dataclass
class RKHSProblemType:
    """RKHSProblemType{KT,T,XT}(c, X, θ, σ²) (declarations.jl:226-231)."""
    c: np.ndarray
    X: np.ndarray
    θ: object
    σ2: float
    _h: Optional[Handle] = field(default=None, repr=False)
    _fitted: bool = field(default=False, repr=False)

    def __post_init__(self):
        self.X = _as_points(self.X)


def fitRKHS_(η: RKHSProblemType, y) -> None:
    """fitRKHS!(η, y) (RKHS.jl:182-217): η.c[:] = (K + σ²I) \\ y."""
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    if η.X.shape[0] == 0 or y.shape[0] == 0:
        raise PMKError(_lib.PMK_ERR_ARG, "AssertionError: !isempty(η.X) && !isempty(y)  (RKHS.jl:199-200)")
    if η.X.shape[0] != y.shape[0]:
        raise PMKError(_lib.PMK_ERR_ARG, "AssertionError: M == length(y)  (RKHS.jl:203)")
    if η._h is None:
        η._h = Handle(0)
    L = lib()
    n, D = η.X.shape
    leaf_off = np.array([0, n], dtype=np.int64)
    kp = η.θ.params
    bad, info = C.c_int64(0), C.c_int(0)
    rc = L.pmk_fit(η._h.raw, D, 1, ptr(leaf_off), ptr(η.X), ptr(y), η.θ.kernel_id, ptr(kp), kp.shape[0], float(η.σ2),
                   C.byref(bad), C.byref(info))
    if rc == _lib.PMK_ERR_NOT_POSDEF:
        # the reference's LU solve does not throw here; surface the failure instead of returning garbage
        raise PosDefException(info.value, bad.value, L.pmk_last_error(η._h.raw).decode())
    η._h.check(rc)
    η._h.check(L.pmk_set_tree(η._h.raw, D, 1, None, None))
    out = np.empty(n)
    η._h.check(L.pmk_get_alpha(η._h.raw, 1, ptr(out)))
    η.c[:] = out
    η._fitted = True
    return None


def query_(Yq: np.ndarray, Xq, η: RKHSProblemType) -> None:
    """query!(Yq, Xq, η) (RKHS.jl:220-247): Yq[iq] = dot(kq, η.c), mean only."""
    Xq = _as_points(Xq)
    if Xq.shape[0] == 0:
        raise PMKError(_lib.PMK_ERR_ARG, "AssertionError: !isempty(Xq)  (RKHS.jl:225)")
    if Yq.shape != (Xq.shape[0],):
        raise PMKError(_lib.PMK_ERR_ARG, "AssertionError: size(Yq) == size(Xq)  (RKHS.jl:227)")
    if not η._fitted:
        raise PMKError(_lib.PMK_ERR_STATE, "query before fitRKHS_")
    wp = np.array([1.0])
    η._h.check(lib().pmk_query(η._h.raw, Xq.shape[0], ptr(Xq), 0.0, 0.0, 1, ptr(wp), 1, 1, ptr(Yq), None))
    return None


def setupGPquery(c, X, θ, σ2: float):
    """setupGPquery(c, X, θ, σ²) (src/RKHS/querying.jl:43-58).  Returns fq with fq(xq) -> (mean, variance) for one point
    (evalqueryGP!, :60-79: mean = Σ c_n k(xq,X_n), variance = k(xq,xq) - kᵀ(K+σ²I)⁻¹k, NOT clamped) or, for an (m, D)
    array of points, two arrays.  The reference solves A\\k by LU per query; here ‖L⁻¹k‖² through the fused pair kernel."""
    X = _as_points(X)
    c = np.ascontiguousarray(np.asarray(c, dtype=np.float64))
    if c.shape[0] != X.shape[0]:
        raise PMKError(_lib.PMK_ERR_ARG, "DimensionMismatch: length(c) != length(X)")
    h = Handle(0)
    L = lib()
    n, D = X.shape
    leaf_off = np.array([0, n], dtype=np.int64)
    kp = θ.params
    bad, info = C.c_int64(0), C.c_int(0)
    y0 = np.zeros(n)          # bound to a name: ptr() hands out a bare address
    rc = L.pmk_fit(h.raw, D, 1, ptr(leaf_off), ptr(X), ptr(y0), θ.kernel_id, ptr(kp), kp.shape[0], float(σ2),
                   C.byref(bad), C.byref(info))
    if rc == _lib.PMK_ERR_NOT_POSDEF:
        raise PosDefException(info.value, bad.value, L.pmk_last_error(h.raw).decode())
    h.check(rc)
    h.check(L.pmk_set_alpha(h.raw, 1, ptr(c)))
    h.check(L.pmk_set_tree(h.raw, D, 1, None, None))
    wp = np.array([1.0])

    def fq(xq):
        xq = np.asarray(xq, dtype=np.float64)
        single = xq.ndim == 1 and (D > 1 or xq.shape[0] == 1)
        Xq = _as_points(xq[None, :] if single else xq)
        Yq, Vq = np.empty(Xq.shape[0]), np.empty(Xq.shape[0])
        h.check(L.pmk_query(h.raw, Xq.shape[0], ptr(Xq), 0.0, 0.0, 1, ptr(wp), 1, 2, ptr(Yq), ptr(Vq)))
        return (float(Yq[0]), float(Vq[0])) if single else (Yq, Vq)

    fq.handle = h
    return fq


def evalquery(x, c, X, θ):
    """evalquery(x, c, X, θ) = Σ c_n k(x, X_n) (src/RKHS/querying.jl:2-5), for one point or an (m, D) array."""
    x = np.asarray(x, dtype=np.float64)
    Xp = _as_points(X)
    single = x.ndim == 1 and (Xp.shape[1] > 1 or x.shape[0] == 1)
    K = constructkernelmatrix(x[None, :] if single else x, θ, Xp)
    out = K @ np.asarray(c, dtype=np.float64)
    return float(out[0]) if single else out
