"""Portable synthetic inputs: splitmix64 counter-based generator, the same stream
in numpy here and in the C baseline under oracle/, so workloads are reproducible on any box
(Julia's Random.seed!(25) stream, examples/mixGP.jl:27, cannot be reproduced outside Julia)."""
from __future__ import annotations

import numpy as np

_G = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """n uint64 outputs of splitmix64 started at `seed`, skipping `offset` outputs."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + _G * (np.arange(1, n + 1, dtype=np.uint64) + np.uint64(offset))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """doubles in [0,1): top 53 bits."""
    return (splitmix64(seed, n, offset) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform_points(seed: int, n: int, lo, hi) -> np.ndarray:
    """(n, D) points uniform on the box [lo, hi]; dimension d uses stream offset d*n.
    convertcompactdomain(rand(), 0, 1, a, b) = (x-0)*(b-a)/(1-0)+a  (examples/helpers/utils.jl:29-32)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    D = lo.shape[0]
    X = np.empty((n, D))
    for d in range(D):
        u = uniform01(seed, n, d * n)
        X[:, d] = (u - 0.0) * (hi[d] - lo[d]) / (1.0 - 0.0) + lo[d]
    return X


def f_mixgp(X: np.ndarray) -> np.ndarray:
    """examples/mixGP.jl:44-48: sinc((x'Ax/3.2)^2) * (norm(x)/4)^3, A = 0.1*[1 .4; .4 1] (first two dims)."""
    A = 0.1 * np.array([[1.0, 0.4], [0.4, 1.0]])
    q = np.einsum("ni,ij,nj->n", X[:, :2], A, X[:, :2])
    return np.sinc((q / 3.2) ** 2) * (np.linalg.norm(X, axis=1) / 4.0) ** 3


def f_ibb1d(x: np.ndarray) -> np.ndarray:
    """examples/IBB1D.jl:31: sinc(4x) * x^3 (Julia sinc(x) = sin(pi x)/(pi x) = numpy sinc)."""
    return np.sinc(4.0 * x) * x ** 3
