"""Small fit through the public host API (debugging aid): python tools/fit_small.py [n_points] [levels]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import patchmixturekriging_b200 as P
from patchmixturekriging_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(1)
X = rng.random((N, 2))
y = np.sin(6 * X[:, 0]) * np.cos(4 * X[:, 1])
root, _, _ = P.setuppartition(X, levels)
X_set, X_set_inds, _, _ = P.organizetrainingsets(root, levels, X, 0.02)
th = P.GaussianKernel1DType(20.0)
eta = P.MixtureGPType(X_set, P.fetchhyperplanes(root))
P.fitmixtureGP_(eta, [y[i - 1] for i in X_set_inds], th, 1e-3)
worst = 0.0
for p in range(len(X_set)):
    L = eta.L_set[p]
    n = L.shape[0]
    Xp = X_set[p]
    d2 = ((Xp[:, None, :] - Xp[None, :, :]) ** 2).sum(-1)
    K = np.exp(-20.0 * d2) + 1e-3 * np.eye(n)
    Lr = np.linalg.cholesky(K)
    worst = max(worst, np.abs(L - Lr).max() / np.abs(Lr).max())
    a = np.linalg.solve(K, y[X_set_inds[p] - 1])
    worst = max(worst, np.abs(eta.c_set[p] - a).max() / np.abs(a).max())
print("leaves", len(X_set), "sizes", [x.shape[0] for x in X_set][:8], "worst rel err", worst)
