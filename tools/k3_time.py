"""Pair-kernel time per size class on a workload (used with experiment builds: PMK_LIB=... python tools/k3_time.py c3_mini)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
import patchmixturekriging_b200 as P
from patchmixturekriging_b200 import _lib
w = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3_mini")
root, sizes, leaf_off, Xp, yp = bench.partition(w)
X_set = [Xp[leaf_off[p]:leaf_off[p + 1]] for p in range(len(sizes))]
y_set = [yp[leaf_off[p]:leaf_off[p + 1]] for p in range(len(sizes))]
eta = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
th = P.GaussianKernel1DType(w["eps_sq"]); wth = P.Spline34KernelType(1.0 / w["radius"])
P.fitmixtureGP_(eta, y_set, th, w["sigma2"])
Xq = bench.gen_queries(w, 0, w["nq"])
Yq = np.empty(len(Xq)); Vq = np.empty(len(Xq))
best = None
for rep in range(4):
    P.querymixtureGP_(Yq, Vq, Xq, eta, root, w["levels"], w["radius"], w["delta"], th, w["sigma2"], wth)
    t = eta.handle.timings()
    cur = [float(t[_lib.T_Q_PAIRS])] + [float(t[_lib.T_Q_PAIRS_CLASS0 + c]) for c in range(5)]
    best = cur if best is None else [min(a, b) for a, b in zip(best, cur)]
# operand build (once per fit): time a fresh fit + first query
P.fitmixtureGP_(eta, y_set, th, w["sigma2"])
P.querymixtureGP_(Yq, Vq, Xq, eta, root, w["levels"], w["radius"], w["delta"], th, w["sigma2"], wth)
t = eta.handle.timings()
print("operands: make_M ms %.3f, P = inv(L) ms %.3f" % (float(t[_lib.T_Q_MAKE_M]), float(t[_lib.T_Q_INVERT])))
print(os.path.basename(os.environ.get("PMK_LIB", "product")), "pairs ms %.3f" % best[0], "by class", ["%.3f" % x for x in best[1:]],
      "checksum", float(np.nansum(Yq)), float(np.nansum(Vq)))
