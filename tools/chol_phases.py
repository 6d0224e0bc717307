"""Per-phase cycle breakdown of k_chol on a workload (instrumentation via pmk_debug_counters)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
import patchmixturekriging_b200 as P
from patchmixturekriging_b200 import _lib
w = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
root, sizes, leaf_off, Xp, yp = bench.partition(w)
X_set = [Xp[leaf_off[p]:leaf_off[p + 1]] for p in range(len(sizes))]
y_set = [yp[leaf_off[p]:leaf_off[p + 1]] for p in range(len(sizes))]
eta = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
th = P.GaussianKernel1DType(w["eps_sq"])
P.fitmixtureGP_(eta, y_set, th, w["sigma2"])
out = np.zeros(8, dtype=np.uint64)
eta.handle.check(_lib.lib().pmk_debug_counters(eta.handle.raw, _lib.ptr(out), 1))
P.fitmixtureGP_(eta, y_set, th, w["sigma2"])
eta.handle.check(_lib.lib().pmk_debug_counters(eta.handle.raw, _lib.ptr(out), 1))
names = ["total", "diag_block(warp0)", "offdiag_work(warp0)", "factor(warp0)", "panel_solve", "barrier_wait", "ctas"]
n = max(float(out[6]), 1.0)   # counters are zero unless built with -DPMK_PROFILE_CYCLES
print({k: round(float(v) / n) for k, v in zip(names, out[:6])}, "ctas", int(n), "chol ms", eta.handle.timings()[_lib.T_FIT_CHOL], "gram tiles ms", eta.handle.timings()[_lib.T_FIT_GRAM])
