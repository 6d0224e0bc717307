"""How close can two correct FP64 algorithms be on the ill-conditioned configurations?  (CPU only.)
Per leaf of mixgp_file / c5_mini / c3_mini at sigma2 = 1e-5 and for IBB1D N = 1000: the posterior mean and variance by
  truth      : Cholesky + substitution in x87 extended precision (numpy longdouble, 64-bit mantissa)
  oracle     : the reference's path -- LU for alpha, dpotrf + dtrsv for the variance
  dtrsm      : the same factor, substitution batched over right-hand sides
  inv        : s = inv(L) kq with the explicit inverse (what the default GPU pair kernel does), + 1 / 2 refinement steps
  chol[_ref] : alpha by Cholesky (+ one refinement step with an FP64 / an extended-precision residual)
Errors are max |a - b| / rms(b) over the leaf's home queries.  Output: profiles/parity_floor_r02.json."""
import sys, time, json
import numpy as np, scipy.linalg as sla
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases, helpers
from oracle import pmk_oracle as O
LD = np.longdouble

def chol_ld(U):
    n = U.shape[0]; A = U.astype(LD).copy()
    for j in range(n):
        if j: A[j:, j] -= A[j:, :j] @ A[j, :j]
        A[j, j] = np.sqrt(A[j, j]); A[j+1:, j] /= A[j, j]
    return np.tril(A)
def fwd_ld(L, B):   # L lower LD, B (n,m)
    n = L.shape[0]; S = B.astype(LD).copy()
    for i in range(n):
        if i: S[i] -= L[i, :i] @ S[:i]
        S[i] /= L[i, i]
    return S
def bwd_ld(L, B):
    n = L.shape[0]; S = B.astype(LD).copy()
    for i in range(n-1, -1, -1):
        if i < n-1: S[i] -= L[i+1:, i] @ S[i+1:]
        S[i] /= L[i, i]
    return S

def leaf_study(X, y, th, s2, Xq, name):
    K = O.constructkernelmatrix(X, th); U = K + s2*np.eye(len(X))
    kq = O.kernel_cross(Xq, X, th).T.copy()   # n x m
    kss = 1.0 if True else None
    # truth
    Lx = chol_ld(U); Sx = fwd_ld(Lx, kq); vx = (LD(1.0) - (Sx*Sx).sum(0)).astype(np.float64)
    ax = bwd_ld(Lx, fwd_ld(Lx, y[:, None]))[:, 0]; ux = (kq.astype(LD).T @ ax).astype(np.float64)
    # oracle
    Lo = O.cholesky_L(U); 
    So = np.stack([sla.solve_triangular(Lo, kq[:, j], lower=True) for j in range(kq.shape[1])], 1)   # dtrsv each
    vo = 1.0 - (So*So).sum(0)
    ao = O.backslash(U, y); uo = kq.T @ ao
    # dtrsm
    Sm = sla.solve_triangular(Lo, kq, lower=True); vm = 1.0 - (Sm*Sm).sum(0)
    # explicit inverse
    P = sla.solve_triangular(Lo, np.eye(len(X)), lower=True); Sp = P @ kq; vp = 1.0 - (Sp*Sp).sum(0)
    # one refinement step fp64
    Sr = Sp + P @ (kq - Lo @ Sp); vr = 1.0 - (Sr*Sr).sum(0)
    Sr2 = Sr + P @ (kq - Lo @ Sr); vr2 = 1.0 - (Sr2*Sr2).sum(0)
    # cholesky alpha and refinement
    ac = sla.cho_solve((Lo, True), y); uc = kq.T @ ac
    r = y - U @ ac; acr = ac + sla.cho_solve((Lo, True), r); ucr = kq.T @ acr
    rl = (y.astype(LD) - U.astype(LD) @ ac.astype(LD)).astype(np.float64); acx = ac + sla.cho_solve((Lo, True), rl); ucx = kq.T @ acx
    sv = np.sqrt(np.mean(vx**2)); su = np.sqrt(np.mean(ux**2))
    e = lambda a, b, s: float(np.abs(a-b).max()/s)
    out = dict(name=name, n=len(X), cond=float(np.linalg.cond(U)), var_rms=float(sv), mean_rms=float(su),
      var=dict(oracle_vs_truth=e(vo,vx,sv), dtrsm_vs_truth=e(vm,vx,sv), inv_vs_truth=e(vp,vx,sv), inv_ref1_vs_truth=e(vr,vx,sv), inv_ref2_vs_truth=e(vr2,vx,sv),
               dtrsm_vs_oracle=e(vm,vo,sv), inv_vs_oracle=e(vp,vo,sv), inv_ref1_vs_oracle=e(vr,vo,sv), inv_ref2_vs_oracle=e(vr2,vo,sv)),
      mean=dict(lu_vs_truth=e(uo,ux,su), chol_vs_truth=e(uc,ux,su), chol_ref_vs_truth=e(ucr,ux,su), chol_refx_vs_truth=e(ucx,ux,su),
                chol_vs_lu=e(uc,uo,su), chol_ref_vs_lu=e(ucr,uo,su), chol_refx_vs_lu=e(ucx,uo,su)))
    return out

res = []
# ibb1d N=1000
for kind in ("BB10", "BB20"):
    c = cases.ibb1d(1000, 100, kind); th, _ = helpers.kernels(c["kernel"])
    # k** != 1 for BB; variance study irrelevant (mean only) but fine
    t = time.time(); r = leaf_study(c["X"], c["y"], th, c["sigma2"], c["Xq"], c["name"]); r.pop("var"); res.append(r); print(json.dumps(r), time.time()-t, flush=True)
for nm in ("mixgp_file", "c5_mini", "c3_mini"):
    c = getattr(cases, nm)();
    if nm == "c3_mini": c["sigma2"] = 1e-5
    th, _ = helpers.kernels(c["kernel"])
    root, Xp, Xpi = O.setuppartition(c["X"], c["levels"]); hv, hc = O.fetchhyperplanes(root)
    inds = O.organizetrainingsets_vec(hv, hc, c["levels"], c["X"], c["eps"])
    home = O._descend_vec(c["Xq"], hv, hc, c["levels"])
    for leaf in (0, len(inds)-1):
        i = inds[leaf]-1; X = c["X"][i]; y = c["y"][i]
        # queries near this leaf: take those whose home is leaf+1
        if home is not None: Xq = c["Xq"][home == leaf+1][:300]
        else: Xq = X[:300] + 1e-3
        t = time.time(); r = leaf_study(X, y, th, c["sigma2"], Xq, f"{nm}(s2={c['sigma2']}) leaf {leaf+1}"); res.append(r); print(json.dumps(r), time.time()-t, flush=True)
json.dump(res, open(sys.argv[1] if len(sys.argv) > 1 else 'profiles/parity_floor_r02.json', 'w'), indent=1)
