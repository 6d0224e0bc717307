"""Generates tests/golden/*.npz from the CPU oracle (oracle/pmk_oracle.py).

The reference is Julia and cannot run in this image, and its own test-suite holds no vectors (test/runtests.jl:4-6),
so these are ORACLE-generated fixtures ("parity unpinned", see oracle/pmk_oracle.py): they freeze the oracle's
answers on the reference's own example workloads so that (a) a change to the oracle is noticed and (b) the CUDA
path can be checked on a box without re-running the slow scalar oracle.  Inputs come from the portable splitmix64
stream (patchmixturekriging_b200/synth.py), so the fixtures are reproducible anywhere:

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import helpers  # noqa: E402
from oracle import pmk_oracle as O  # noqa: E402


def mixgp(name, case, nq_stride):
    m = helpers.oracle_model(case)
    oth, _ = helpers.kernels(case["kernel"])
    owth, _ = helpers.kernels(case["wkernel"])
    Xq = case["Xq"][::nq_stride]
    Yq, Vq, dbg = O.querymixtureGP(Xq, m["eta"], m["root"], case["levels"], case["radius"], case["delta"], oth, case["sigma2"],
                                   owth, debug=True)
    sizes = np.array([len(i) for i in m["X_set_inds"]], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), hps_v=m["hv"], hps_c=m["hc"], set_sizes=sizes,
                        set_inds_head=np.concatenate([i[:8] for i in m["X_set_inds"]]),
                        set_inds_sum=np.array([int(i.sum()) for i in m["X_set_inds"]], dtype=np.int64),
                        Xq=Xq, Yq=Yq, Vq=Vq, home=np.array(dbg["p_region_ind"], dtype=np.int32),
                        npairs=np.array([len(r) + 1 for r in dbg["region_inds"]], dtype=np.int32),
                        region_inds=np.concatenate([np.array(r, dtype=np.int32) for r in dbg["region_inds"]] + [np.zeros(0, np.int32)]),
                        alpha0_head=m["eta"].c_set[0][:16], L0_diag=np.diag(m["eta"].L_set[0]))


def ibb1d(kind):
    case = cases.ibb1d(15, 100, kind)
    th, _ = helpers.kernels(case["kernel"])
    c = O.fitRKHS(case["X"], case["y"], th, case["sigma2"])
    yq = O.query_rkhs(case["Xq"], case["X"], c, th)
    K = O.constructkernelmatrix(case["X"], th)
    np.savez_compressed(os.path.join(HERE, f"ibb1d_{kind}.npz"), c=c, yq=yq, K=K)


if __name__ == "__main__":
    mixgp("mixgp_file", cases.mixgp_file(), 50)                     # examples/mixGP.jl as written (400 of the 20 000 grid queries)
    mixgp("mixgp_sqexp", cases.mixgp_driver(N=3000, levels=4), 50)  # the SqExp variant at a size the scalar oracle handles
    ibb1d("BB10")                                                   # examples/IBB1D.jl
    ibb1d("BB20")
    print("written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
