"""ctypes binding of libpmk_b200.so (C ABI in include/pmk.h).  There is no fallback: if the
library is missing or no B200 is visible, the calls raise."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMK_LIB") or os.path.join(_HERE, "libpmk_b200.so")   # PMK_LIB: experiment builds (build.py PMK_VARIANT)

PMK_OK, PMK_ERR_CUDA, PMK_ERR_ARG, PMK_ERR_NOT_POSDEF, PMK_ERR_STATE, PMK_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
OPT_FULL_HYPERPLANE_SCAN, OPT_QUERY_SOLVER, OPT_INVERSE_BUILDER, OPT_ALPHA_REFINE, OPT_CHOL_VARIANT, OPT_GRAM_FAST_EXP = 1, 2, 3, 4, 5, 6
SOLVER_AUTO, SOLVER_INVERSE, SOLVER_SUBSTITUTION = -1, 0, 1
T_FIT_PACK, T_FIT_CHOL, T_FIT_SOLVE, T_Q_TREE, T_Q_PAIRS, T_Q_COMBINE, T_GRAM, T_COUNT = 0, 1, 2, 3, 4, 5, 6, 17
T_Q_MAKE_M = 13
T_Q_INVERT = 14
T_Q_PAIRS_CLASS0 = 8
T_FIT_GRAM = 7
T_FIT_REFINE = 15
T_Q_ROUTE_SORT = 16
MT_FIT, MT_QUERY, MT_Q_PLAN, MT_Q_ROUTE, MT_Q_PAIRS, MT_Q_RETURN, MT_COUNT = 0, 1, 2, 3, 4, 5, 8

# every symbol include/pmk.h declares
SYMBOLS = [
    "pmk_create", "pmk_destroy", "pmk_last_error", "pmk_version", "pmk_inverse_plan", "pmk_gram", "pmk_cross_gram", "pmk_fit", "pmk_fit_dev",
    "pmk_leaf_size", "pmk_get_alpha", "pmk_set_alpha", "pmk_get_L", "pmk_get_Linv", "pmk_get_K", "pmk_set_tree", "pmk_find_partition", "pmk_organize_training_sets", "pmk_organize_fetch", "pmk_query",
    "pmk_query_dev", "pmk_last_query_pairs", "pmk_last_query_debug", "pmk_last_query_debug_dense", "pmk_last_query_leaf_pairs", "pmk_build_M", "pmk_query_plan_dev",
    "pmk_query_pairs_dev", "pmk_query_combine_dev", "pmk_set_leaf_base", "pmk_query_plan_segments", "pmk_query_plan_pack_dev", "pmk_query_pairs_routed_dev",
    "pmk_query_plan_unpack_dev", "pmk_query_set_flags", "pmk_set_option", "pmk_condition_estimate", "pmk_measure_fp64_peak",
    "pmk_multi_create", "pmk_multi_destroy", "pmk_multi_last_error", "pmk_multi_size", "pmk_multi_leaf_range", "pmk_multi_query_range", "pmk_multi_owned_range", "pmk_multi_balanced_ranges", "pmk_multi_handle",
    "pmk_multi_set_option", "pmk_multi_fit", "pmk_multi_set_tree", "pmk_multi_query", "pmk_multi_stage_training", "pmk_multi_fit_staged",
    "pmk_multi_stage_queries", "pmk_multi_query_staged", "pmk_multi_fetch_results", "pmk_multi_leaf_pairs", "pmk_multi_get_timings", "pmk_multi_launch_count", "pmk_partition_begin", "pmk_partition_level_z", "pmk_partition_level_split", "pmk_partition_fetch",
    "pmk_partition_sum_plan", "pmk_save_model", "pmk_load_model", "pmk_model_info", "pmk_get_X", "pmk_get_tree", "pmk_get_timings", "pmk_debug_counters", "pmk_launch_count", "pmk_stream", "pmk_synchronize",
]

_lib = None


class PMKError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpmk_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class PosDefException(PMKError):
    """Mirror of LinearAlgebra.PosDefException thrown by cholesky(U) (reference mixtureGP.jl:109)."""

    def __init__(self, info: int, leaf: int, msg: str):
        PMKError.__init__(self, PMK_ERR_NOT_POSDEF, msg)
        self.info = info
        self.leaf = leaf


def lib() -> C.CDLL:
    """Load the CUDA library; raises if it was not built (python -m patchmixturekriging_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m patchmixturekriging_b200.build` "
                           "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    dp = C.c_void_p  # raw pointers (host numpy or device addresses)
    L.pmk_create.argtypes = [C.POINTER(vp), i32]
    L.pmk_destroy.argtypes = [vp]
    L.pmk_destroy.restype = None
    L.pmk_last_error.argtypes = [vp]
    L.pmk_last_error.restype = C.c_char_p
    L.pmk_version.argtypes = []
    L.pmk_inverse_plan.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.pmk_gram.argtypes = [vp, i32, i64, dp, i32, dp, i32, dbl, dp]
    L.pmk_cross_gram.argtypes = [vp, i32, i64, dp, i64, dp, i32, dp, i32, dp]
    L.pmk_fit.argtypes = [vp, i32, i64, dp, dp, dp, i32, dp, i32, dbl, C.POINTER(i64), C.POINTER(i32)]
    L.pmk_fit_dev.argtypes = L.pmk_fit.argtypes
    L.pmk_leaf_size.argtypes = [vp, i64, C.POINTER(i64)]
    L.pmk_get_alpha.argtypes = [vp, i64, dp]
    L.pmk_set_alpha.argtypes = [vp, i64, dp]
    L.pmk_get_L.argtypes = [vp, i64, dp]
    L.pmk_get_K.argtypes = [vp, i64, dp]
    L.pmk_get_Linv.argtypes = [vp, i64, dp]
    L.pmk_set_tree.argtypes = [vp, i32, i32, dp, dp]
    L.pmk_find_partition.argtypes = [vp, i64, dp, dp]
    L.pmk_organize_training_sets.argtypes = [vp, i64, dp, dbl, dp, C.POINTER(i64)]
    L.pmk_organize_fetch.argtypes = [vp, dp, dp, dp]
    L.pmk_query.argtypes = [vp, i64, dp, dbl, dbl, i32, dp, i32, i32, dp, dp]
    L.pmk_query_dev.argtypes = L.pmk_query.argtypes
    L.pmk_last_query_pairs.argtypes = [vp, C.POINTER(i64)]
    L.pmk_last_query_debug.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp, dp]
    L.pmk_last_query_debug_dense.argtypes = [vp, i64, i64, dp, dp, dp]
    L.pmk_last_query_leaf_pairs.argtypes = [vp, dp]
    L.pmk_build_M.argtypes = [vp]
    L.pmk_set_leaf_base.argtypes = [vp, i64, i64]
    L.pmk_query_plan_segments.argtypes = [vp, i32, dp, dp]
    L.pmk_query_plan_pack_dev.argtypes = [vp, dp, dp]
    L.pmk_query_pairs_routed_dev.argtypes = [vp, i64, dp, dp, i32, dp, dp]
    L.pmk_query_plan_unpack_dev.argtypes = [vp, dp, dp, dp, dp]
    L.pmk_query_set_flags.argtypes = [vp, i32]
    L.pmk_condition_estimate.argtypes = [vp, C.POINTER(dbl), C.POINTER(i32)]
    L.pmk_measure_fp64_peak.argtypes = [vp, C.POINTER(dbl)]
    L.pmk_multi_create.argtypes = [C.POINTER(vp), i32, dp]
    L.pmk_multi_destroy.argtypes = [vp]
    L.pmk_multi_destroy.restype = None
    L.pmk_multi_last_error.argtypes = [vp]
    L.pmk_multi_last_error.restype = C.c_char_p
    L.pmk_multi_size.argtypes = [vp]
    L.pmk_multi_leaf_range.argtypes = [i32, i64, i32, C.POINTER(i64), C.POINTER(i64)]
    L.pmk_multi_owned_range.argtypes = [C.c_void_p, i32, C.POINTER(i64), C.POINTER(i64)]
    L.pmk_multi_balanced_ranges.argtypes = [i32, i64, vp, vp]
    L.pmk_multi_query_range.argtypes = [i32, i64, i32, C.POINTER(i64), C.POINTER(i64)]
    L.pmk_multi_handle.argtypes = [vp, i32, C.POINTER(vp)]
    L.pmk_multi_set_option.argtypes = [vp, i32, i64]
    L.pmk_multi_fit.argtypes = [vp, i32, i64, dp, dp, dp, i32, dp, i32, dbl, C.POINTER(i64), C.POINTER(i32)]
    L.pmk_multi_set_tree.argtypes = [vp, i32, i32, dp, dp]
    L.pmk_multi_query.argtypes = [vp, i64, dp, dbl, dbl, i32, dp, i32, i32, dp, dp]
    L.pmk_multi_stage_training.argtypes = [vp, i32, i64, dp, dp, dp]
    L.pmk_multi_fit_staged.argtypes = [vp, i32, dp, i32, dbl, C.POINTER(i64), C.POINTER(i32)]
    L.pmk_multi_stage_queries.argtypes = [vp, i64, dp]
    L.pmk_multi_query_staged.argtypes = [vp, dbl, dbl, i32, dp, i32, i32]
    L.pmk_multi_fetch_results.argtypes = [vp, dp, dp]
    L.pmk_multi_leaf_pairs.argtypes = [vp, dp]
    L.pmk_multi_get_timings.argtypes = [vp, dp, dp]
    L.pmk_multi_launch_count.argtypes = [vp]
    L.pmk_multi_launch_count.restype = i64
    L.pmk_query_plan_dev.argtypes = [vp, i64, dp, dbl, dbl, i32, dp, i32, C.POINTER(i64)]
    L.pmk_query_pairs_dev.argtypes = [vp, i32, dp, dp]
    L.pmk_query_combine_dev.argtypes = [vp, dp, dp, dp, dp]
    L.pmk_set_option.argtypes = [vp, i32, i64]
    L.pmk_partition_begin.argtypes = [vp, i32, i64, dp, i32]
    L.pmk_partition_level_z.argtypes = [vp, i32, dp]
    L.pmk_partition_level_split.argtypes = [vp, i32, dp, dp]
    L.pmk_partition_fetch.argtypes = [vp, dp, dp]
    L.pmk_partition_sum_plan.argtypes = [i64, i64, dp, dp, dp, C.POINTER(i64)]
    L.pmk_save_model.argtypes = [vp, C.c_char_p]
    L.pmk_load_model.argtypes = [vp, C.c_char_p]
    L.pmk_model_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i32)]
    L.pmk_get_X.argtypes = [vp, i64, dp]
    L.pmk_get_tree.argtypes = [vp, dp, dp]
    L.pmk_get_timings.argtypes = [vp, dp]
    L.pmk_debug_counters.argtypes = [vp, dp, i32]
    L.pmk_launch_count.argtypes = [vp]
    L.pmk_launch_count.restype = i64
    L.pmk_stream.argtypes = [vp]
    L.pmk_stream.restype = vp
    L.pmk_synchronize.argtypes = [vp]
    _lib = L
    return L


def ptr(a) -> int | None:
    """Raw address of a contiguous numpy array (or an int device address, passed through).  The address is only valid while
    the caller keeps a reference to `a`: bind temporaries to a name before passing them (ptr(np.zeros(n)) would hand the
    library freed memory)."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    assert a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
    return a.ctypes.data


class Handle:
    """One fitted model on one GPU (pmk_handle)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        L = lib()
        rc = L.pmk_create(C.byref(self._h), int(device))
        if rc != PMK_OK:
            msg = L.pmk_last_error(None).decode()
            self._h = C.c_void_p()
            raise PMKError(rc, msg)
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().pmk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != PMK_OK:
            raise PMKError(rc, lib().pmk_last_error(self._h).decode())

    @property
    def raw(self):
        return self._h

    def timings(self) -> np.ndarray:
        ms = np.zeros(T_COUNT)
        self.check(lib().pmk_get_timings(self._h, ptr(ms)))
        return ms

    def launch_count(self) -> int:
        return int(lib().pmk_launch_count(self._h))

    def synchronize(self):
        self.check(lib().pmk_synchronize(self._h))


class _BorrowedHandle(Handle):
    """A rank's handle of a MultiHandle: owned by the pmk_multi object, never destroyed from here."""

    def __init__(self, raw, device):
        self._h = raw
        self.device = device

    def close(self):
        self._h = C.c_void_p()


class MultiHandle:
    """One model over several GPUs of a box by sub-tree ownership (pmk_multi)."""

    def __init__(self, devices):
        self._m = C.c_void_p()
        L = lib()
        devices = list(range(devices)) if isinstance(devices, (int, np.integer)) else [int(d) for d in devices]
        ids = np.asarray(devices, dtype=np.int32)
        rc = L.pmk_multi_create(C.byref(self._m), len(devices), ptr(ids))
        if rc != PMK_OK:
            msg = L.pmk_multi_last_error(None).decode()
            self._m = C.c_void_p()
            raise PMKError(rc, msg)
        self.devices = devices

    @property
    def raw(self):
        return self._m

    @property
    def size(self) -> int:
        return len(self.devices)

    def close(self):
        if getattr(self, "_m", None) is not None and self._m.value:
            lib().pmk_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != PMK_OK:
            raise PMKError(rc, lib().pmk_multi_last_error(self._m).decode())

    def rank_handle(self, rank: int) -> Handle:
        h = C.c_void_p()
        self.check(lib().pmk_multi_handle(self._m, rank, C.byref(h)))
        return _BorrowedHandle(h, self.devices[rank])

    def owned_range(self, rank: int):
        """(first, count): the 0-based leaves `rank` owns (pmk_multi_owned_range; valid once the training data is staged)."""
        a, c = C.c_int64(0), C.c_int64(0)
        self.check(lib().pmk_multi_owned_range(self._m, rank, C.byref(a), C.byref(c)))
        return a.value, c.value

    def owner_of_leaf(self, leaf0: int, n_leaves: int | None = None) -> int:
        """rank owning the 0-based leaf (the library's cost-balanced leaf -> rank map)."""
        for r in range(self.size):
            a, c = self.owned_range(r)
            if a <= leaf0 < a + c:
                return r
        raise IndexError(leaf0)

    def timings(self):
        ms = np.zeros(MT_COUNT)
        per = np.zeros((self.size, T_COUNT))
        self.check(lib().pmk_multi_get_timings(self._m, ptr(ms), ptr(per)))
        return ms, per

    def launch_count(self) -> int:
        return int(lib().pmk_multi_launch_count(self._m))
