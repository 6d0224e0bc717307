"""ctypes loader of oracle/_build/libpmk_oracle.so (C restatement; test infrastructure / CPU baseline).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_build", "libpmk_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pmk_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads() -> int:
    return int(lib().pmk_oracle_max_threads())


def fit(X_set, y_set, kind: int, param: float, sigma2: float, nthreads: int = 0, want_L: bool = True):
    """Returns alpha (packed), L (packed dense blocks), leaf_off.  Raises on a non-PD leaf."""
    sizes = np.array([len(x) for x in X_set], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    X = np.ascontiguousarray(np.concatenate(X_set, axis=0), dtype=np.float64)
    y = np.ascontiguousarray(np.concatenate(y_set), dtype=np.float64)
    alpha = np.empty(off[-1])
    L = np.empty(int((sizes ** 2).sum())) if want_L else None
    bad, info = C.c_int64(0), C.c_int(0)
    rc = lib().pmk_oracle_fit(C.c_int(X.shape[1]), C.c_int64(len(sizes)), _p(off), _p(X), _p(y), C.c_int(kind), C.c_double(param),
                              C.c_double(sigma2), _p(alpha), _p(L), C.byref(bad), C.byref(info), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"not positive definite: leaf {bad.value} info {info.value}")
    return alpha, L, off, X


def query(hv, hc, levels, off, X, alpha, L, kind, param, Xq, radius, delta, wkind, wparam, nthreads: int = 0,
          structure_only: bool = False):
    """Returns Yq, Vq, home, npairs, absent (absent[j] = 1 if query j touches a leaf of zero length in `off`)."""
    Xq = np.ascontiguousarray(Xq, dtype=np.float64)
    hv = np.ascontiguousarray(hv, dtype=np.float64)
    hc = np.ascontiguousarray(hc, dtype=np.float64)
    off = np.ascontiguousarray(off, dtype=np.int64)
    Nq, D = Xq.shape
    Yq, Vq = np.empty(Nq), np.empty(Nq)
    home = np.empty(Nq, dtype=np.int32)
    npairs = np.empty(Nq, dtype=np.int32)
    absent = np.empty(Nq, dtype=np.int32)
    lib().pmk_oracle_query(C.c_int(D), C.c_int(levels), _p(hv), _p(hc), C.c_int64(len(off) - 1), _p(off), _p(X), _p(alpha), _p(L),
                           C.c_int(kind), C.c_double(param), C.c_int64(Nq), _p(Xq), C.c_double(radius), C.c_double(delta),
                           C.c_int(wkind), C.c_double(wparam), _p(Yq), _p(Vq), _p(home), _p(npairs), _p(absent),
                           C.c_int(1 if structure_only else 0), C.c_int(nthreads))
    return Yq, Vq, home, npairs, absent


def gram(X, kind, param, sigma2=0.0):
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, D = X.shape
    K = np.empty((n, n), order="F")
    lib().pmk_oracle_gram(C.c_int(D), C.c_int64(n), _p(X), C.c_int(kind), C.c_double(param), C.c_double(sigma2), _p(K))
    return K
