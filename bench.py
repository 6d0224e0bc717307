#!/usr/bin/env python
"""bench.py -- local-GP fit + mixture query on B200 (BASELINE.json: "GP fit leaves/s and query pts/s").

One STEP = one full pass of the hot path over the synthetic workload: fit every BSP leaf
(fused Gram + Cholesky + alpha), then answer every query (home leaf, neighbours, fused
cross-covariance / mean / variance, convex mixture).  `value` = query points per second (the
10M-point mixture query dominates the step); the fit throughput is reported beside it as
`fit_leaves_per_s`.  Inputs are resident in HBM for `value`; `e2e` repeats the step through the
public host API (patchmixturekriging_b200.fitmixtureGP_ / querymixtureGP_) with pinned host buffers,
host<->device copies inside the timed region.

    python bench.py --gpus N --steps K --warmup W [--workload c3|c3_mini|c2] [--impl reference]

N > 1 is launched by torchrun, one rank per GPU: leaves are dealt to ranks by a contiguous
leaf->rank map (fit), the factors are exchanged with NCCL broadcasts over NVLink, queries are sliced
across ranks with no data-path collective and gathered once at the end (strong scaling of the named
workload).  `--impl reference` times the CPU restatement of the reference's algorithm (oracle/pmk_oracle.c,
all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_PEAK_FALLBACK_TFLOPS = 37.1   # tools/fp64_peak.cu on this pool's B200 (profiles/fp64_peaks_r01.json)
def eval_flops(D):                 # one SqExp evaluation: differences/squares/sum + sqrt + exp (DESIGN.md "flop counts")
    return 3 * D + 2 + 20


# --------------------------------------------------------------------------------------------
def workload(name: str, nq_override: int | None = None):
    from patchmixturekriging_b200 import synth
    lo, hi = [-5.0, -10.0], [5.0, 10.0]
    if name == "c3":          # BASELINE configs[2]: 2-D, N = 1M, 4096 leaves of ~512 points with overlap, 10M queries
        N, levels, eps, nq = 1_000_000, 13, 0.043, 10_000_000
    elif name == "c3_mini":   # same shape, 1/16 size (for quick runs), same point density
        N, levels, eps, nq = 62_500, 9, 0.043, 625_000
        lo, hi = [-1.25, -2.5], [1.25, 2.5]
    elif name == "c2":        # BASELINE configs[1]: examples/mixGP.jl, SqExp, N = 20k, 64 leaves
        N, levels, eps, nq = 20_000, 7, 0.5, 20_000
    elif name == "c4":        # BASELINE configs[3]: 3-D, N = 4M, 8192 leaves of ~1024 points
        N, levels, eps, nq = 4_000_000, 14, 0.075, 10_000_000
        lo, hi = [-5.0, -10.0, -5.0], [5.0, 10.0, 5.0]
    elif name == "c4_mini":   # 1/16 of c4, same density
        N, levels, eps, nq = 250_000, 10, 0.075, 625_000
        lo, hi = [-2.5, -5.0, -1.25], [2.5, 5.0, 1.25]
    else:
        raise SystemExit(f"unknown workload {name}")
    D = len(lo)
    nq = nq_override or nq
    X = synth.uniform_points(25, N, lo, hi)
    y = synth.f_mixgp(X)
    vol = float(np.prod(np.asarray(hi) - np.asarray(lo)))
    spacing = (vol / N) ** (1.0 / D)
    if name == "c2":
        eps_sq = 8.0
    elif D == 2:
        eps_sq = round(1.0 / (3.5 * spacing) ** 2)      # c3: 408 ~ SURVEY's "eps_sq ~ 400" (length-scale ~3.5 spacings)
    else:
        eps_sq = round(1.0 / (2.0 * spacing) ** 2)      # 3-D: length-scale ~2 spacings
    radius = 0.3 if name == "c2" else eps
    return dict(name=name, X=X, y=y, levels=levels, eps=eps, radius=radius, delta=1e-5, sigma2=1e-3, eps_sq=float(eps_sq),
                nq=nq, lo=lo, hi=hi, D=D)


def gen_queries(w, first: int, count: int) -> np.ndarray:
    """queries [first, first+count) of the workload's stream (uniform on the domain)."""
    from patchmixturekriging_b200 import synth
    lo, hi = np.asarray(w["lo"]), np.asarray(w["hi"])
    D = len(lo)
    Xq = np.empty((count, D))
    for d in range(D):
        u = synth.uniform01(1234567, count, d * w["nq"] + first)
        Xq[:, d] = u * (hi[d] - lo[d]) + lo[d]
    return Xq


def partition(w, device: bool = False):
    import patchmixturekriging_b200 as P
    root, X_parts, X_parts_inds = P.setuppartition(w["X"], w["levels"])
    org = P.organizetrainingsets_device if device else P.organizetrainingsets
    X_set, X_set_inds, _, _ = org(root, w["levels"], w["X"], w["eps"])
    sizes = np.array([len(i) for i in X_set_inds], dtype=np.int64)
    leaf_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    idx = np.concatenate(X_set_inds) - 1
    return root, sizes, leaf_off, np.ascontiguousarray(w["X"][idx]), np.ascontiguousarray(w["y"][idx])


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads() -> int:
    """Host threads the CPU arm may use: the cores this process may run on, whatever OMP_NUM_THREADS says."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def fp64_peak():
    p = os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")
    try:
        d = json.load(open(p))
        return float(d["dmma_tflops"]), "measured: tools/fp64_peak.cu DMMA.8x8x4 loop on this pool's B200 (profiles/fp64_peaks_r01.json); MEASURED_PEAKS.json has no FP64 entry"
    except Exception:
        return FP64_PEAK_FALLBACK_TFLOPS, "fallback constant (profiles/fp64_peaks_r01.json missing)"


def hbm_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json (driver-written copy bandwidth)"
    except Exception:
        return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json missing)"


# --------------------------------------------------------------------------------------------
def cpu_baseline(w, root, sizes, leaf_off, Xp, yp, steps=1, warmup=0, sample_leaves=512, sample_queries=200000, threads=0):
    """The reference's algorithm (C restatement, oracle/pmk_oracle.c) on the host cores, bounded sample:
    the first `sample_leaves` leaves of the REAL tree are fitted; sample queries are drawn from the workload's own
    query stream among those whose home and neighbour leaves are all in that set (the all-hyperplanes scan of
    findneighbourpartitions runs over the full tree, as in the reference)."""
    from oracle import c_oracle
    SQEXP, SPLINE34 = 0, 1
    # all host cores, explicitly: torchrun exports OMP_NUM_THREADS=1 into every rank, which omp_get_max_threads() obeys
    threads = threads or host_threads()
    nl = min(sample_leaves, len(sizes))
    off_s = leaf_off.copy()
    off_s[nl + 1:] = off_s[nl]                      # leaves >= nl absent
    Xs, ys = Xp[:leaf_off[nl]], yp[:leaf_off[nl]]
    X_set = [Xs[leaf_off[p]:leaf_off[p + 1]] for p in range(nl)]
    y_set = [ys[leaf_off[p]:leaf_off[p + 1]] for p in range(nl)]
    hv, hc = np.ascontiguousarray(root.hps_v), np.ascontiguousarray(root.hps_c)
    # candidate queries: those of the stream whose home leaf is in the sample (host findpartition), then the
    # structure-only pass of the oracle keeps the ones that touch no absent leaf
    import patchmixturekriging_b200 as P
    cand = []
    first, chunk = 0, 400_000
    while sum(len(c) for c in cand) < 3 * sample_queries and first < w["nq"]:
        cnt = min(chunk, w["nq"] - first)
        Xq = gen_queries(w, first, cnt)
        home = P.findpartition(Xq, root)
        cand.append(Xq[home <= nl])
        first += cnt
    cand = np.concatenate(cand) if cand else np.zeros((0, w["D"]))
    _, _, _, _, absent = c_oracle.query(hv, hc, w["levels"], off_s, Xs, np.zeros(len(Xs)), np.zeros(1), SQEXP, w["eps_sq"], cand,
                                        w["radius"], w["delta"], SPLINE34, 1.0 / w["radius"], threads, structure_only=True)
    Xq_s = np.ascontiguousarray(cand[absent == 0][:sample_queries])
    fit_t, q_t = [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        alpha, L, off2, Xpk = c_oracle.fit(X_set, y_set, SQEXP, w["eps_sq"], w["sigma2"], threads)
        t1 = time.perf_counter()
        Yq, Vq, _, npairs, ab = c_oracle.query(hv, hc, w["levels"], off_s, Xpk, alpha, L, SQEXP, w["eps_sq"], Xq_s, w["radius"],
                                               w["delta"], SPLINE34, 1.0 / w["radius"], threads)
        t2 = time.perf_counter()
        if it >= warmup:
            fit_t.append(t1 - t0); q_t.append(t2 - t1)
    assert np.isfinite(Yq).all()
    return dict(fit_leaves_per_s=nl / np.mean(fit_t), query_pts_per_s=len(Xq_s) / np.mean(q_t), cores=threads,
                sample=f"first {nl} leaves of the real {len(sizes)}-leaf tree fitted ({np.mean(fit_t):.2f} s); {len(Xq_s)} queries of the "
                       f"workload's stream whose leaves are all in that set ({np.mean(q_t):.2f} s); full-tree hyperplane scan per query",
                fit_s=float(np.mean(fit_t)), query_s=float(np.mean(q_t)), sample_xq=Xq_s, sample_y=Yq, sample_v=Vq,
                pairs_per_query=float(npairs.mean()))


# --------------------------------------------------------------------------------------------
class _CudaSpan:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--nq", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference" and rank != 0:
        return                      # the CPU arm runs on rank 0 alone; the other ranks exit 0 without work
    w = workload(args.workload, args.nq)
    cfg = {"workload": f"{w['name']}: {w['D']}-D mixture-GP, N={len(w['X'])}, {1 << (w['levels'] - 1)} BSP leaves, eps={w['eps']}, "
                       f"radius={w['radius']}, delta={w['delta']}, SqExp eps_sq={w['eps_sq']}, sigma2={w['sigma2']}, Nq={w['nq']}",
           "l2": "inputs larger than L2 (packed factors + queries + pair arrays are GBs per step vs 126 MB L2)"}

    # ---------------- reference arm: CPU restatement on the host cores -----------------------
    if args.impl == "reference":
        root, sizes, leaf_off, Xp, yp = partition(w)
        # a bounded run whatever --steps says: one step of the sample is seconds of CPU work, three give a stable mean
        ref_steps, ref_warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
        cb = cpu_baseline(w, root, sizes, leaf_off, Xp, yp, steps=ref_steps, warmup=ref_warmup)
        line = {"impl": "reference", "metric": "query_pts_per_s (mixture-GP query; fit throughput in fit_leaves_per_s)",
                "value": cb["query_pts_per_s"], "unit": "pts/s", "fit_leaves_per_s": cb["fit_leaves_per_s"], "n_gpus": args.gpus,
                "steps": ref_steps, "warmup": ref_warmup, "steps_requested": args.steps,
                "ms_per_step": 1e3 * (cb["fit_s"] + cb["query_s"]),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "cpu_baseline": {"value": cb["query_pts_per_s"], "unit": "pts/s", "fit_leaves_per_s": cb["fit_leaves_per_s"],
                                 "cores": cb["cores"], "kind": "port", "sample": cb["sample"]},
                "e2e": {"value": cb["query_pts_per_s"], "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------- B200 arm ---------------------------------------------------------------
    import torch
    import torch.distributed as dist
    import patchmixturekriging_b200 as P
    from patchmixturekriging_b200 import _lib, mixturegp, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    if world > 1:
        # NCCL_DEBUG=VERSION (set in some images) makes NCCL print its version to STDOUT, next to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))

    root, sizes, leaf_off, Xp, yp = partition(w, device=True)
    n_leaves = len(sizes)
    Nq = w["nq"]
    q0, q1 = sharding.query_slice(rank, world, Nq)
    Xq_h = gen_queries(w, q0, q1 - q0)
    nq_loc = q1 - q0
    l0, lc = sharding.leaf_range(rank, world, n_leaves)
    l1 = l0 + lc

    θ = P.GaussianKernel1DType(w["eps_sq"])
    wθ = P.Spline34KernelType(1.0 / w["radius"])
    X_set = [Xp[leaf_off[p]:leaf_off[p + 1]] for p in range(n_leaves)]
    y_set = [yp[leaf_off[p]:leaf_off[p + 1]] for p in range(n_leaves)]
    η = P.MixtureGPType(X_set, P.fetchhyperplanes(root), device=dev, fit_range=(l0, l1 - l0) if world > 1 else None)
    h = η.handle
    L = _lib.lib()
    hv = np.ascontiguousarray(root.hps_v); hc = np.ascontiguousarray(root.hps_c)
    h.check(L.pmk_set_tree(h.raw, w["D"], w["levels"], _lib.ptr(hv), _lib.ptr(hc)))
    if world > 1:
        h.check(L.pmk_set_fit_range(h.raw, l0, l1 - l0))

    dX = torch.from_numpy(Xp).cuda(); dy = torch.from_numpy(yp).cuda(); dXq = torch.from_numpy(Xq_h).cuda()
    dYq = torch.empty(nq_loc, dtype=torch.float64, device="cuda"); dVq = torch.empty_like(dYq)
    gathered = [None, None]
    kp = θ.params; wp = wθ.params
    stream = torch.cuda.ExternalStream(int(L.pmk_stream(h.raw)), device=dev)
    bad, info = C.c_int64(0), C.c_int(0)

    def span_of(which, first, count):
        ptr, nb = mixturegp.model_buffer(η, which, first, count)
        return torch.as_tensor(_CudaSpan(ptr, nb), device=f"cuda:{dev}") if nb else torch.empty(0, dtype=torch.float64, device="cuda")

    def exchange_factors():
        """leaf -> rank map: every rank broadcasts the spans it factorised (NCCL over NVLink)."""
        mixturegp.build_M(η)          # pair-kernel operands of the leaves this rank factorised; only P = inv(L) travels
        h.synchronize()
        sharding.exchange_spans(span_of, n_leaves, (_lib.BUF_P, _lib.BUF_ALPHA))
        torch.cuda.current_stream().synchronize()
        mixturegp.mark_fitted(η, p_exchanged=True)

    def step_device():
        """fit + query with inputs resident in HBM; returns (fit_ms, exchange_ms, query_ms) from CUDA events on the
        handle's stream (the NCCL work is bracketed by stream synchronisation, so the events see it)."""
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(stream)
        h.check(L.pmk_fit_dev(h.raw, w["D"], n_leaves, _lib.ptr(leaf_off), dX.data_ptr(), dy.data_ptr(), θ.kernel_id, _lib.ptr(kp), 1,
                              w["sigma2"], C.byref(bad), C.byref(info)))
        e[1].record(stream)
        if world > 1:
            h.synchronize()
            exchange_factors()
        e[2].record(stream)
        h.check(L.pmk_query_dev(h.raw, nq_loc, dXq.data_ptr(), w["radius"], w["delta"], wθ.kernel_id, _lib.ptr(wp), 1, 0,
                                dYq.data_ptr(), dVq.data_ptr()))
        if world > 1:
            h.synchronize()
            gathered[0] = sharding.gather_slices(dYq, Nq)
            gathered[1] = sharding.gather_slices(dVq, Nq)
            torch.cuda.current_stream().synchronize()
        e[3].record(stream)
        e[3].synchronize()
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = h.launch_count()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    fit_ms, exch_ms, query_ms, kt = [], [], [], []
    for _ in range(args.steps):
        f, x, q = step_device()
        fit_ms.append(f + x); exch_ms.append(x); query_ms.append(q)
        kt.append(h.timings().copy())
    ev1.record(stream)
    barrier()
    ev1.synchronize()
    total_ms = ev0.elapsed_time(ev1)
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clocks = sampler.stop() if rank == 0 else None
    launches = h.launch_count() - launches0

    t = torch.tensor([total_ms, float(np.mean(fit_ms)), float(np.mean(query_ms)), float(np.mean(exch_ms))], dtype=torch.float64,
                     device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, fit_ms_m, query_ms_m, exch_ms_m = (float(x) for x in t.cpu())
    kt = np.mean(np.array(kt), axis=0)

    # ---------------- e2e through the public host API, pinned host buffers --------------------
    e2e = None
    if not args.no_e2e:
        hXq = torch.from_numpy(Xq_h).pin_memory().numpy()
        hYq = torch.empty(nq_loc, dtype=torch.float64).pin_memory().numpy()
        hVq = torch.empty(nq_loc, dtype=torch.float64).pin_memory().numpy()
        hXp = torch.from_numpy(Xp).pin_memory().numpy()
        hyp = torch.from_numpy(yp).pin_memory().numpy()
        η.X_parts = [hXp[leaf_off[p]:leaf_off[p + 1]] for p in range(n_leaves)]
        y_set_p = [hyp[leaf_off[p]:leaf_off[p + 1]] for p in range(n_leaves)]

        def step_host():
            t0 = time.perf_counter()
            # fitmixtureGP_ packs the leaf list; the packed arrays are what crosses the ABI
            h.check(L.pmk_fit(h.raw, w["D"], n_leaves, _lib.ptr(leaf_off), _lib.ptr(hXp), _lib.ptr(hyp), θ.kernel_id, _lib.ptr(kp), 1,
                              w["sigma2"], C.byref(bad), C.byref(info)))
            if world > 1:
                exchange_factors()
            η._fitted, η.θ = True, θ
            t1 = time.perf_counter()
            P.querymixtureGP_(hYq, hVq, hXq, η, root, w["levels"], w["radius"], w["delta"], θ, w["sigma2"], wθ)
            if world > 1:
                gathered[0] = sharding.gather_slices(torch.from_numpy(hYq).cuda(), Nq)
                gathered[1] = sharding.gather_slices(torch.from_numpy(hVq).cuda(), Nq)
                torch.cuda.synchronize()
            t2 = time.perf_counter()
            return 1e3 * (t1 - t0), 1e3 * (t2 - t1)

        step_host()
        barrier()
        ef, eq = [], []
        for _ in range(max(1, min(args.steps, 3))):
            a, b = step_host()
            ef.append(a); eq.append(b)
        te = torch.tensor([float(np.mean(ef)), float(np.mean(eq))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ef_m, eq_m = (float(x) for x in te.cpu())
        e2e = {"value": Nq / (eq_m * 1e-3), "unit": "pts/s", "fit_leaves_per_s": n_leaves / (ef_m * 1e-3),
               "h2d_bytes_per_step": int(Xq_h.nbytes + Xp.nbytes + yp.nbytes), "d2h_bytes_per_step": int(2 * 8 * nq_loc),
               "ms_fit": ef_m, "ms_query": eq_m,
               "api": "pmk_fit(host X,y) + patchmixturekriging_b200.querymixtureGP_(Yq, Vq, Xq, ...) with pinned host arrays"}

    # ---------------- roofline of the dominant kernel (K3, k_query_pairs) ---------------------
    if rank == 0:
        npairs = C.c_int64(0)
        h.check(L.pmk_last_query_pairs(h.raw, C.byref(npairs)))
        pl = np.empty(npairs.value, dtype=np.int32)
        h.check(L.pmk_last_query_debug(h.raw, None, None, _lib.ptr(pl), None, None, None, None, None))
        per_leaf = np.bincount(pl - 1, minlength=n_leaves).astype(np.float64)
        nn = sizes.astype(np.float64)
        flops_pairs = float((per_leaf * (nn * nn + nn * (eval_flops(w["D"]) + 4))).sum())      # TRSM n^2 + n*(eval + mean 2 + ||s||^2 2)
        flops_fit = float((nn ** 3 / 3 + 2 * nn * nn + nn * (nn + 1) / 2 * eval_flops(w["D"])).sum()) * (l1 - l0) / n_leaves if world > 1 else \
            float((nn ** 3 / 3 + 2 * nn * nn + nn * (nn + 1) / 2 * eval_flops(w["D"])).sum())
        peak, peak_src = fp64_peak()
        ach = flops_pairs / (kt[_lib.T_Q_PAIRS] * 1e-3) / 1e12
        try:      # DRAM bytes of the dominant kernel from the committed ncu --set full capture (per launch, leaf class <=512)
            tr_ = json.load(open(os.path.join(ROOT, "profiles", "k3_traffic_r01.json")))
            traffic = float(tr_["dram_bytes_read"] + tr_["dram_bytes_write"])
            traffic_of = {"kernel": tr_["kernel"], "source": tr_["source"]}
        except Exception:
            traffic, traffic_of = None, None
        roofline = {"kernel": "k_query_rowp (fused cross-covariance + mean + variance, s = inv(L) kq as a row-panel product on DMMA, per (query, leaf) pair)", "bound": "tensor",
                    "pipe": "FP64 DMMA.8x8x4 (mma.sync.m8n8k4.f64); tcgen05 has no f64 kind", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak, "peak_source": peak_src, "traffic": traffic, "traffic_of": traffic_of,
                    "algorithmic_flops_per_launch": flops_pairs, "ms_per_launch": float(kt[_lib.T_Q_PAIRS]),
                    "pairs_per_launch": int(npairs.value)}
        fit_ach = flops_fit / ((kt[_lib.T_FIT_CHOL] + kt[_lib.T_FIT_GRAM]) * 1e-3) / 1e12
        phases = {"fit_pack_ms": float(kt[_lib.T_FIT_PACK]), "fit_gram_ms": float(kt[_lib.T_FIT_GRAM]),
                  "fit_chol_ms": float(kt[_lib.T_FIT_CHOL]),
                  "fit_solve_ms": float(kt[_lib.T_FIT_SOLVE]), "query_tree_ms": float(kt[_lib.T_Q_TREE]),
                  "query_make_M_ms": float(kt[_lib.T_Q_MAKE_M]), "query_invert_ms": float(kt[_lib.T_Q_INVERT]), "query_pairs_ms": float(kt[_lib.T_Q_PAIRS]),
                  "query_combine_ms": float(kt[_lib.T_Q_COMBINE]),
                  "fit_gram_chol_tflops": fit_ach, "fit_gram_chol_frac_of_fp64_peak": fit_ach / peak}
        # HBM side (north star: "achieved HBM GB/s for Gram build"): the fit's Gram tiles write the lower tiles of every leaf
        # once (8 B x packed factor doubles); the standalone Gram kernel (constructkernelmatrix) writes 8 n^2 B.
        hbm_peak, hbm_src = hbm_peak_gbs()
        try:
            _, l_bytes = mixturegp.model_buffer(η, _lib.BUF_L, l0, l1 - l0)
            gt = l_bytes / (kt[_lib.T_FIT_GRAM] * 1e-3) / 1e9
            ng = 8192
            Xg = np.ascontiguousarray(w["X"][:ng])
            hg = _lib.Handle(dev)
            Kg = np.empty((ng, ng), order="F")
            tg = []
            for _ in range(3):
                hg.check(L.pmk_gram(hg.raw, w["D"], ng, _lib.ptr(Xg), θ.kernel_id, _lib.ptr(kp), kp.shape[0], 0.0, _lib.ptr(Kg)))
                tg.append(float(hg.timings()[_lib.T_GRAM]))
            hg.close()
            g_ms = float(np.mean(tg[1:]))
            phases["hbm"] = {"peak_gbs": hbm_peak, "peak_source": hbm_src,
                             "k_gram_tiles": {"bytes_written": int(l_bytes), "ms": float(kt[_lib.T_FIT_GRAM]), "gbs": gt, "frac": gt / hbm_peak},
                             "k_gram": {"what": f"constructkernelmatrix, n={ng} (8 n^2 = {8 * ng * ng >> 20} MiB written, > L2)", "ms": g_ms,
                                        "gbs": 8.0 * ng * ng / (g_ms * 1e-3) / 1e9, "frac": 8.0 * ng * ng / (g_ms * 1e-3) / 1e9 / hbm_peak}}
        except Exception as exc:      # an instrumentation extra: never lose the bench line to it
            phases["hbm"] = {"error": repr(exc)}
        # per leaf-size class of the pair kernel: ms and achieved TFLOP/s
        npad_ = (sizes + 31) // 32 * 32
        cls_ = np.where(npad_ <= 512, 0, np.where(npad_ <= 768, 1, np.where(npad_ <= 1024, 2, np.where(npad_ <= 1536, 3, 4))))
        phases["leaf_points_min_mean_max"] = [int(sizes.min()), float(sizes.mean()), int(sizes.max())]
        fl_leaf = per_leaf * (nn * nn + nn * (eval_flops(w["D"]) + 4))
        phases["pairs_by_class"] = [
            {"class": c, "leaves": int((cls_ == c).sum()), "pairs": int(per_leaf[cls_ == c].sum()),
             "ms": float(kt[_lib.T_Q_PAIRS_CLASS0 + c]),
             "tflops": float(fl_leaf[cls_ == c].sum() / max(kt[_lib.T_Q_PAIRS_CLASS0 + c], 1e-9) / 1e9)}
            for c in range(5) if (cls_ == c).any()]

        line = {"metric": "query_pts_per_s (mixture-GP query; fit throughput in fit_leaves_per_s)", "value": Nq / (query_ms_m * 1e-3),
                "unit": "pts/s", "fit_leaves_per_s": n_leaves / (fit_ms_m * 1e-3), "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "ms_fit": fit_ms_m, "ms_query": query_ms_m,
                "ms_factor_exchange": exch_ms_m, "fit_compute_leaves_per_s": n_leaves / max((fit_ms_m - exch_ms_m) * 1e-3, 1e-9),
                "wall_ms_per_step": wall_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks, "gpu_launches": int(launches),
                "roofline": roofline, "phases": phases}
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(w, root, sizes, leaf_off, Xp, yp)
            # parity spot check of the sample against the GPU through the public API
            Yg, Vg, _ = P.querymixtureGP(cb["sample_xq"], η, root, w["levels"], w["radius"], w["delta"], θ, w["sigma2"], wθ)
            sc_y, sc_v = np.sqrt(np.mean(cb["sample_y"] ** 2)), np.sqrt(np.mean(cb["sample_v"] ** 2))
            line["cpu_baseline"] = {"value": cb["query_pts_per_s"], "unit": "pts/s", "fit_leaves_per_s": cb["fit_leaves_per_s"],
                                    "cores": cb["cores"], "kind": "port", "sample": cb["sample"],
                                    "parity_vs_gpu": {"mean_max_err_over_rms": float(np.abs(Yg - cb["sample_y"]).max() / sc_y),
                                                      "var_max_err_over_rms": float(np.abs(Vg - cb["sample_v"]).max() / sc_v)}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
