"""k_chol timing across build variants (libpmk_b200_<name>.so, see build.py PMK_VARIANT) on one workload, partition built once.
usage: python tools/chol_variants.py c3 product nw12 nw12:2 ...   (name[:PMK_CHOL_CTAS_PER_SM])"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import patchmixturekriging_b200 as P
from patchmixturekriging_b200 import _lib

w = bench.workload(sys.argv[1])
root, sizes, leaf_off, Xp, yp = bench.partition(w, device=True)
X_set = [Xp[leaf_off[p]:leaf_off[p + 1]] for p in range(len(sizes))]
y_set = [yp[leaf_off[p]:leaf_off[p + 1]] for p in range(len(sizes))]
th = P.GaussianKernel1DType(w["eps_sq"])
here = os.path.dirname(_lib.LIB_PATH)
ref = None
for spec in sys.argv[2:]:
    name, _, ctas = spec.partition(":")
    _lib._lib = None
    _lib.LIB_PATH = os.path.join(here, "libpmk_b200.so" if name == "product" else f"libpmk_b200_{name}.so")
    if ctas:
        os.environ["PMK_CHOL_CTAS_PER_SM"] = ctas
    else:
        os.environ.pop("PMK_CHOL_CTAS_PER_SM", None)
    try:
        eta = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
        ts = []
        for _ in range(4):
            P.fitmixtureGP_(eta, y_set, th, w["sigma2"])
            ts.append(eta.handle.timings().copy())
        t = np.min(np.array(ts)[1:], axis=0)
        L0 = eta.L_set[len(sizes) // 2]
        if ref is None:
            ref = L0
        print(f"{spec:12s} chol {t[_lib.T_FIT_CHOL]:7.3f} ms  gram {t[_lib.T_FIT_GRAM]:6.3f}  solve {t[_lib.T_FIT_SOLVE]:6.3f}  "
              f"L identical to first variant: {np.array_equal(L0, ref)}", flush=True)
        eta.close()
    except Exception as e:
        print(f"{spec:12s} FAILED: {e}", flush=True)
