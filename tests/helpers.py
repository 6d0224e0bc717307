from __future__ import annotations

import numpy as np

import patchmixturekriging_b200 as P
from oracle import pmk_oracle as O

_K = {"SQEXP": (O.SQEXP, P.GaussianKernel1DType), "SPLINE34": (O.SPLINE34, P.Spline34KernelType),
      "BB10": (O.BB10, P.BrownianBridge10), "BB20": (O.BB20, P.BrownianBridge20),
      "BB1EPS": (O.BB1EPS, P.BrownianBridge1ϵ), "BB2EPS": (O.BB2EPS, P.BrownianBridge2ϵ),
      "SPLINE12": (O.SPLINE12, P.Spline12KernelType), "SPLINE32": (O.SPLINE32, P.Spline32KernelType),
      "RQ": (O.RQ, P.RationalQuadraticKernelType)}


def kernels(spec):
    """(oracle kernel, product kernel descriptor) for a (name, param) spec."""
    name, p = spec
    ok, pk = _K[name]
    return O.Kernel(ok, p), pk(p)


def oracle_model(case):
    """Partition + ε-overlap sets + fit with the oracle.  Returns dict."""
    X, y = case["X"], case["y"]
    th, _ = kernels(case["kernel"])
    root, X_parts, X_parts_inds = O.setuppartition(X, case["levels"])
    hv, hc = O.fetchhyperplanes(root)
    X_set_inds = O.organizetrainingsets_vec(hv, hc, case["levels"], X, case["eps"])
    X_set = [X[i - 1] for i in X_set_inds]
    eta = O.MixtureGP(X_set, hv, hc)
    O.fitmixtureGP(eta, [y[i - 1] for i in X_set_inds], th, case["sigma2"], keep_U=False)
    return dict(root=root, hv=hv, hc=hc, X_parts_inds=X_parts_inds, X_set_inds=X_set_inds, eta=eta, th=th)


def err_stats(a, b):
    """norm-wise and point-wise relative errors of a against reference b."""
    a, b = np.asarray(a), np.asarray(b)
    d = np.abs(a - b)
    scale = np.sqrt(np.mean(b * b)) if b.size else 1.0
    with np.errstate(divide="ignore", invalid="ignore"):
        pw = np.where(b != 0, d / np.abs(b), 0.0)
    return dict(max_abs=float(d.max()) if d.size else 0.0, normwise=float(d.max() / scale) if d.size else 0.0,
                pointwise=float(pw.max()) if d.size else 0.0,
                frac_pw_gt_1e9=float(np.mean(pw > 1e-9)) if d.size else 0.0)


# ---- measured parity errors, written next to the assertions (profiles/parity_r02.json is a copy of a GPU run's file) ----------
import json as _json
import os as _os

_PARITY_PATH = _os.path.join(_os.environ.get("GRAFT_REPO_ROOT", _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))),
                             "gpurun_out", "parity_r02.json")


def record_parity(key: str, **vals):
    """Append measured errors (err_stats dicts or plain numbers) under `key`; -q swallows prints, this file does not."""
    try:
        _os.makedirs(_os.path.dirname(_PARITY_PATH), exist_ok=True)
        data = {}
        if _os.path.exists(_PARITY_PATH):
            with open(_PARITY_PATH) as f:
                data = _json.load(f)
        data.setdefault(key, {}).update(vals)
        with open(_PARITY_PATH, "w") as f:
            _json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass


def bound_stats(a, b):
    """err_stats plus the quantity the tests bound: max |a - b| / max(|b|, rms(b)) ("floored relative error")."""
    a, b = np.asarray(a), np.asarray(b)
    st = err_stats(a, b)
    scale = np.sqrt(np.mean(b * b)) if b.size else 1.0
    st["floored_rel"] = float((np.abs(a - b) / np.maximum(np.abs(b), scale)).max()) if b.size else 0.0
    return st
