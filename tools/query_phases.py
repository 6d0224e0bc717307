"""Per-phase cycle breakdown of k_query_pairs on a workload (instrumentation via pmk_debug_counters)."""
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
out = np.zeros(8, dtype=np.uint64)
for rep in range(2):
    P.querymixtureGP_(Yq, Vq, Xq, eta, root, w["levels"], w["radius"], w["delta"], th, w["sigma2"], wth)
    eta.handle.check(_lib.lib().pmk_debug_counters(eta.handle.raw, _lib.ptr(out), 3))
# row-panel kernel (default solver): cycles per tile, averaged over the 16 compute warps
names = ["total", "tile start (staging waits, ring prefill)", "phase E", "wait K complete", "phase M", "tiles", "operand waits in M", "epilogue + wait K free"]
n = float(out[5])
print({k: round(float(v) / n) for k, v in zip(names, out) if k != "tiles"}, "tiles", int(n), "pairs ms", eta.handle.timings()[_lib.T_Q_PAIRS])
