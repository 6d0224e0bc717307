"""One fit + one mixture query of a workload through the public host API (profiling target for the query-side kernels):
python tools/query_once.py c3 [nq]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import patchmixturekriging_b200 as P
from patchmixturekriging_b200 import _lib

w = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3", int(sys.argv[2]) if len(sys.argv) > 2 else None)
root, sizes, leaf_off, Xp, yp = bench.partition(w, device=True)
X_set = [Xp[leaf_off[p]:leaf_off[p + 1]] for p in range(len(sizes))]
y_set = [yp[leaf_off[p]:leaf_off[p + 1]] for p in range(len(sizes))]
eta = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
th = P.GaussianKernel1DType(w["eps_sq"]); wth = P.Spline34KernelType(1.0 / w["radius"])
P.fitmixtureGP_(eta, y_set, th, w["sigma2"])
Xq = bench.gen_queries(w, 0, w["nq"])
Yq = np.empty(len(Xq)); Vq = np.empty(len(Xq))
P.querymixtureGP_(Yq, Vq, Xq, eta, root, w["levels"], w["radius"], w["delta"], th, w["sigma2"], wth)
t = eta.handle.timings()
print(f"tree {t[_lib.T_Q_TREE]:.3f} invert {t[_lib.T_Q_INVERT]:.3f} pairs {t[_lib.T_Q_PAIRS]:.3f} combine {t[_lib.T_Q_COMBINE]:.3f} ms; checksum {np.nansum(Yq):.12e} {np.nansum(Vq):.12e}")
