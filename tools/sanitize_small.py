"""Small end-to-end pass for compute-sanitizer (memcheck): device partition build, mirrored Gram, fit, mixture query, checkpoint."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
import patchmixturekriging_b200 as P

case = cases.mixgp_driver(N=3000, levels=4)
X, y = case["X"], case["y"]
root, _, _ = P.setuppartition_device(X, case["levels"])
X_set, X_set_inds, _, _ = P.organizetrainingsets_device(root, case["levels"], X, case["eps"])
th, wth = P.GaussianKernel1DType(case["kernel"][1]), P.Spline34KernelType(case["wkernel"][1])
K = P.constructkernelmatrix(X[:301], th)
assert np.array_equal(K, K.T)
eta = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
P.fitmixtureGP_(eta, [y[i - 1] for i in X_set_inds], th, case["sigma2"])
Xq = case["Xq"][::40]
Y0, V0, _ = P.querymixtureGP(Xq, eta, root, case["levels"], case["radius"], case["delta"], th, case["sigma2"], wth)
with tempfile.TemporaryDirectory() as d:
    P.savemixtureGP(eta, os.path.join(d, "m.pmk"), root, case["levels"])
    eta2, root2, lv = P.loadmixtureGP(os.path.join(d, "m.pmk"))
Y1, V1, _ = P.querymixtureGP(Xq, eta2, root2, lv, case["radius"], case["delta"], th, case["sigma2"], wth)
assert np.array_equal(Y0, Y1) and np.array_equal(V0, V1)
print("sanitize_small ok:", len(X_set), "leaves,", len(Xq), "queries, launches", eta.handle.launch_count())
