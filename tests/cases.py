"""Seeded workloads shared by the CPU and GPU tests (sizes the oracle finishes in seconds)."""
from __future__ import annotations

import numpy as np

from patchmixturekriging_b200 import synth


def grid2d(n1, n2, lo=(-5.0, -10.0), hi=(5.0, 10.0)):
    """x1 fastest, like vec(X_nD) in examples/mixGP.jl:55-63,158."""
    g1 = np.linspace(lo[0], hi[0], n1)
    g2 = np.linspace(lo[1], hi[1], n2)
    return np.array([[a, b] for b in g2 for a in g1])


def mixgp_file():
    """examples/mixGP.jl as written: N=850, levels=3, Spline34(1/15), eps=1.5, radius=0.3, delta=1e-5, sigma2=1e-5."""
    X = synth.uniform_points(25, 850, [-5.0, -10.0], [5.0, 10.0])
    return dict(name="mixgp_file", X=X, y=synth.f_mixgp(X), levels=3, eps=1.5, radius=0.3, delta=1e-5, sigma2=1e-5,
                kernel=("SPLINE34", 1.0 / 15.0), wkernel=("SPLINE34", 1.0 / 0.3), Xq=grid2d(100, 200))


def mixgp_driver(N=20000, levels=7):
    """BASELINE configs[1]: 2-D BSP mixture GP, SqExp, N=20k, 64 leaves."""
    X = synth.uniform_points(25, N, [-5.0, -10.0], [5.0, 10.0])
    return dict(name="mixgp_driver", X=X, y=synth.f_mixgp(X), levels=levels, eps=0.5, radius=0.3, delta=1e-5, sigma2=1e-3,
                kernel=("SQEXP", 8.0), wkernel=("SPLINE34", 1.0 / 0.3), Xq=grid2d(100, 200))


def c3_mini(N=16384, levels=7, eps=0.31, sigma2=1e-3, nq=30000):
    """Scaled-down C3 (2-D, ~512-point overlapped leaves, radius = eps, SqExp length-scale ~3.5 spacings)."""
    X = synth.uniform_points(7, N, [-5.0, -10.0], [5.0, 10.0])
    spacing = np.sqrt(200.0 / N)
    eps_sq = 1.0 / (3.5 * spacing) ** 2
    Xq = synth.uniform_points(11, nq, [-5.0, -10.0], [5.0, 10.0])
    return dict(name="c3_mini", X=X, y=synth.f_mixgp(X), levels=levels, eps=eps, radius=eps, delta=1e-5, sigma2=sigma2,
                kernel=("SQEXP", eps_sq), wkernel=("SPLINE34", 1.0 / eps), Xq=Xq)


def c3_mini_ill():
    """c3_mini at the stress noise level sigma2 = 1e-5 ("and 1e-5 as a stress variant"): cond(K + sigma2 I) ~ 3e6."""
    c = c3_mini(sigma2=1e-5, nq=12000)
    c["name"] = "c3_mini_ill"
    return c


def c4_mini(N=12000, levels=5, eps=0.35, nq=8000):
    """Scaled-down C4 (3-D, ~1000-point leaves)."""
    lo, hi = [-5.0, -10.0, -5.0], [5.0, 10.0, 5.0]
    X = synth.uniform_points(3, N, lo, hi)
    spacing = (2000.0 / N) ** (1.0 / 3.0)
    eps_sq = 1.0 / (2.0 * spacing) ** 2
    Xq = synth.uniform_points(5, nq, lo, hi)
    return dict(name="c4_mini", X=X, y=synth.f_mixgp(X), levels=levels, eps=eps, radius=eps, delta=1e-5, sigma2=1e-3,
                kernel=("SQEXP", eps_sq), wkernel=("SPLINE34", 1.0 / eps), Xq=Xq)


def c5_mini(G=96, levels=4, eps=0.04):
    """Scaled-down C5 (dev/image_upscale.jl style): training points = a regular G x G pixel grid scaled to [0,1]^2, values =
    a smooth synthetic image, Spline34 kernel with a support radius of 5 pixels, sigma2 = 1e-5, ~1400-1800-point leaves
    (the two largest size classes), query = the 2x upsampled grid."""
    g = np.linspace(0.0, 1.0, G)
    X = np.array([[a, b] for b in g for a in g])
    y = np.sin(7.0 * X[:, 0]) * np.cos(5.0 * X[:, 1]) + 0.5 * np.exp(-8.0 * ((X[:, 0] - 0.6) ** 2 + (X[:, 1] - 0.3) ** 2))
    a = 0.2 * (G - 1)
    return dict(name="c5_mini", X=X, y=y, levels=levels, eps=eps, radius=eps, delta=1e-5, sigma2=1e-5,
                kernel=("SPLINE34", a), wkernel=("SPLINE34", 1.0 / eps), Xq=grid2d(2 * G - 1, 2 * G - 1, (0.0, 0.0), (1.0, 1.0)))


def ibb1d(N=15, Nq=100, kind="BB10"):
    """examples/IBB1D.jl: X = LinRange(1e-5, 1-1e-5, N), y = sinc(4x) x^3, sigma2 = 1e-5, query LinRange(0,1,Nq)."""
    x = np.linspace(1e-5, 1.0 - 1e-5, N)
    return dict(name=f"ibb1d_{kind}_{N}", X=x[:, None], y=synth.f_ibb1d(x), sigma2=1e-5, kernel=(kind, 1.0),
                Xq=np.linspace(0.0, 1.0, Nq)[:, None])
