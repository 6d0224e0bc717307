"""Two fits of a workload through the public host API (profiling target for the fit kernels): python tools/fit_only.py c3 [n_fits]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import patchmixturekriging_b200 as P
from patchmixturekriging_b200 import _lib

w = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
root, sizes, leaf_off, Xp, yp = bench.partition(w, device=True)
X_set = [Xp[leaf_off[p]:leaf_off[p + 1]] for p in range(len(sizes))]
y_set = [yp[leaf_off[p]:leaf_off[p + 1]] for p in range(len(sizes))]
eta = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
th = P.GaussianKernel1DType(w["eps_sq"])
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    P.fitmixtureGP_(eta, y_set, th, w["sigma2"])
    t = eta.handle.timings()
    print(f"gram {t[_lib.T_FIT_GRAM]:.3f} chol {t[_lib.T_FIT_CHOL]:.3f} solve {t[_lib.T_FIT_SOLVE]:.3f} ms")
