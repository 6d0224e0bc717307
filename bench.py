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
    import patchmixturekriging_b200 as P
    from patchmixturekriging_b200 import synth
    if name in ("c5", "c5_mini"):
        return workload_image(name, nq_override)
    lo, hi = [-5.0, -10.0], [5.0, 10.0]
    if name == "c3":          # BASELINE configs[2]: 2-D, N = 1M, 4096 leaves of ~512 points with overlap, 10M queries
        N, levels, eps, nq = 1_000_000, 13, 0.043, 10_000_000
    elif name == "c3_8th":    # one GPU's share of c3 at 8 GPUs: 512 leaves, same point density
        N, levels, eps, nq = 125_000, 10, 0.043, 1_250_000
        lo, hi = [-1.76776695, -3.5355339], [1.76776695, 3.5355339]
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
    eps_sq = float(eps_sq)
    return dict(name=name, X=X, y=y, levels=levels, eps=eps, radius=radius, delta=1e-5, sigma2=1e-3, eps_sq=eps_sq,
                kernel_desc=f"SqExp eps_sq={eps_sq}", kernel_oracle=(0, eps_sq), theta=lambda: P.GaussianKernel1DType(eps_sq),
                nq=nq, lo=lo, hi=hi, D=D, queries="uniform")


def workload_image(name: str, nq_override: int | None = None):
    """BASELINE configs[4]: dev/image_upscale.jl-style image kriging (reference dev/image_upscale.jl:147-158,179-180,217-219):
    training inputs = a regular G x G pixel grid scaled to [0,1]^2, values = a smooth synthetic image, Spline34 kernel with a
    support radius of 5 pixels, sigma2 = 1e-5; queries = the 2x upsampled grid (4096^2 for c5).  levels = 13 with eps = 0.0015
    gives 4096 leaves of 1235 .. 2040 points (mean 1562): every leaf inside PMK_MAX_LEAF_POINTS = 2048, the two largest size
    classes of the pair kernel."""
    import patchmixturekriging_b200 as P
    if name == "c5":
        G, levels, eps = 2048, 13, 0.0015
    else:                     # 1/16 of it, same pixels per leaf
        G, levels, eps = 512, 9, 0.0015 * 4
    g = np.linspace(0.0, 1.0, G)
    X = np.empty((G * G, 2))
    X[:, 0] = np.tile(g, G)                 # x1 fastest, like vec(X_nD) in examples/mixGP.jl:55-63
    X[:, 1] = np.repeat(g, G)
    y = np.sin(7.0 * X[:, 0]) * np.cos(5.0 * X[:, 1]) + 0.5 * np.exp(-8.0 * ((X[:, 0] - 0.6) ** 2 + (X[:, 1] - 0.3) ** 2))
    a = 0.2 * (G - 1)
    Gq = 2 * G
    nq = nq_override or Gq * Gq
    return dict(name=name, X=X, y=y, levels=levels, eps=eps, radius=eps, delta=1e-5, sigma2=1e-5, eps_sq=None,
                kernel_desc=f"Spline34 a={a}", kernel_oracle=(1, a), theta=lambda: P.Spline34KernelType(a),
                nq=nq, lo=[0.0, 0.0], hi=[1.0, 1.0], D=2, queries=f"{Gq}x{Gq} grid", Gq=Gq)


def gen_queries(w, first: int, count: int) -> np.ndarray:
    """queries [first, first+count) of the workload's stream (uniform on the domain; the upsampled grid, x1 fastest, for the
    image workloads)."""
    from patchmixturekriging_b200 import synth
    lo, hi = np.asarray(w["lo"]), np.asarray(w["hi"])
    D = len(lo)
    if w["queries"] != "uniform":
        Gq = w["Gq"]
        g = np.linspace(lo[0], hi[0], Gq)
        idx = np.arange(first, first + count, dtype=np.int64)
        return np.ascontiguousarray(np.column_stack([g[idx % Gq], g[(idx // Gq) % Gq]]))
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
    SPLINE34 = 1
    KIND, KPAR = w["kernel_oracle"]
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
    _, _, _, _, absent = c_oracle.query(hv, hc, w["levels"], off_s, Xs, np.zeros(len(Xs)), np.zeros(1), KIND, KPAR, cand,
                                        w["radius"], w["delta"], SPLINE34, 1.0 / w["radius"], threads, structure_only=True)
    Xq_s = np.ascontiguousarray(cand[absent == 0][:sample_queries])
    fit_t, q_t = [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        alpha, L, off2, Xpk = c_oracle.fit(X_set, y_set, KIND, KPAR, w["sigma2"], threads)
        t1 = time.perf_counter()
        Yq, Vq, _, npairs, ab = c_oracle.query(hv, hc, w["levels"], off_s, Xpk, alpha, L, KIND, KPAR, Xq_s, w["radius"],
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
    ap.add_argument("--no-setup", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference" and rank != 0:
        return                      # the CPU arm runs on rank 0 alone; the other ranks exit 0 without work
    w = workload(args.workload, args.nq)
    cfg = {"workload": f"{w['name']}: {w['D']}-D mixture-GP, N={len(w['X'])}, {1 << (w['levels'] - 1)} BSP leaves, eps={w['eps']}, "
                       f"radius={w['radius']}, delta={w['delta']}, {w['kernel_desc']}, sigma2={w['sigma2']}, Nq={w['nq']} ({w['queries']})",
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
    # One model over the N GPUs of the box by sub-tree ownership, through the library's own multi-GPU entry points
    # (pmk_multi_*: one host thread + stream set per GPU, CUDA peer copies over NVLink) -- the path a caller of
    # fitmixtureGP! / querymixtureGP! reaches.  Under torchrun every rank joins the NCCL process group (launch barrier, final
    # reduction of the timings); rank 0's process drives all N GPUs through the library, the other ranks wait on the host.
    import torch
    import patchmixturekriging_b200 as P
    from patchmixturekriging_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    dist = None
    host_group = None
    if world > 1:
        import torch.distributed as dist
        # NCCL_DEBUG=VERSION (set in some images) makes NCCL print its version to STDOUT, next to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        host_group = dist.new_group(backend="gloo")     # long waits happen on the host: no NCCL kernel spins on a GPU rank 0 is using
        my_dev = torch.device(f"cuda:{local_rank}")
        one = torch.ones(1, device=my_dev)
        dist.all_reduce(one)                            # every rank's GPU is up and NCCL connects the N of them
        torch.cuda.synchronize()
        assert int(one.item()) == world
    n_gpus = world if world > 1 else max(1, args.gpus)
    result = None
    if rank == 0:
        if torch.cuda.device_count() < n_gpus:
            raise SystemExit(f"bench.py: {n_gpus} GPUs asked, {torch.cuda.device_count()} visible to rank 0")
        result = run_b200(args, w, cfg, n_gpus)
    if world > 1:
        dist.barrier(group=host_group)                  # ranks > 0 wait here (gloo: on the host) while rank 0 measures
        t = torch.tensor([result["ms_per_step"] if result else 0.0], dtype=torch.float64, device=my_dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # the contract's max over ranks (rank 0 timed all N GPUs)
        if result:
            result["ms_per_step"] = float(t.item())
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.barrier(group=host_group)
        dist.destroy_process_group()


def measure_setup(w):
    """setuppartition + organizetrainingsets with the O(N) work on the GPU (what a caller pays once per data set, outside the
    timed fit + query step): milliseconds, second of two runs."""
    import patchmixturekriging_b200 as P
    ms = None
    for _ in range(2):
        t0 = time.perf_counter()
        root, _, _ = P.setuppartition_device(w["X"], w["levels"])
        t1 = time.perf_counter()
        P.organizetrainingsets_device(root, w["levels"], w["X"], w["eps"])
        t2 = time.perf_counter()
        ms = {"setuppartition_device_ms": 1e3 * (t1 - t0), "organizetrainingsets_device_ms": 1e3 * (t2 - t1)}
    return ms


def run_b200(args, w, cfg, n_gpus):
    import torch
    import patchmixturekriging_b200 as P
    from patchmixturekriging_b200 import _lib
    L = _lib.lib()
    D = w["D"]
    root, sizes, leaf_off, Xp, yp = partition(w, device=True)
    n_leaves = len(sizes)
    Nq = w["nq"]
    Xq_h = gen_queries(w, 0, Nq)
    θ = w["theta"]()
    wθ = P.Spline34KernelType(1.0 / w["radius"])
    kp, wp = θ.params, wθ.params
    hv, hc = np.ascontiguousarray(root.hps_v), np.ascontiguousarray(root.hps_c)
    # pinned host buffers: what the reference-side caller hands over (packed leaves, queries) and gets back
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    hXp, hyp, hXq = pin(Xp), pin(yp), pin(Xq_h)
    hYq, hVq = pin(np.empty(Nq)), pin(np.empty(Nq))

    m = _lib.MultiHandle(list(range(n_gpus)))
    m.check(L.pmk_multi_set_tree(m.raw, D, w["levels"], _lib.ptr(hv), _lib.ptr(hc)))
    bad, info = C.c_int64(0), C.c_int(0)

    def sync_all():
        for d in range(n_gpus):
            torch.cuda.synchronize(d)

    # ---- value: inputs resident in HBM (staged once), K timed steps of fit + query ------------------------------------
    m.check(L.pmk_multi_stage_training(m.raw, D, n_leaves, _lib.ptr(leaf_off), _lib.ptr(hXp), _lib.ptr(hyp)))
    m.check(L.pmk_multi_stage_queries(m.raw, Nq, _lib.ptr(hXq)))

    def step_device():
        m.check(L.pmk_multi_fit_staged(m.raw, θ.kernel_id, _lib.ptr(kp), kp.shape[0], w["sigma2"], C.byref(bad), C.byref(info)))
        m.check(L.pmk_multi_query_staged(m.raw, w["radius"], w["delta"], wθ.kernel_id, _lib.ptr(wp), wp.shape[0], 0))
        return m.timings()

    for _ in range(args.warmup):
        step_device()
    sync_all()
    launches0 = m.launch_count()
    sampler = ClockSampler(0)
    sampler.start()
    t_wall0 = time.perf_counter()
    fit_ms, query_ms, mt_all, kt_all = [], [], [], []
    for _ in range(args.steps):
        mt, per = step_device()          # device times: CUDA events on every rank's stream, max over the ranks
        fit_ms.append(mt[_lib.MT_FIT]); query_ms.append(mt[_lib.MT_QUERY]); mt_all.append(mt.copy()); kt_all.append(per.copy())
    sync_all()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clocks = sampler.stop()
    launches = m.launch_count() - launches0
    fit_ms_m, query_ms_m = float(np.mean(fit_ms)), float(np.mean(query_ms))
    mt = np.mean(np.array(mt_all), axis=0)
    kt = np.mean(np.array(kt_all), axis=0)            # (n_gpus, T_COUNT) per-kernel times of every rank
    ktmax = kt.max(axis=0)
    pairs_per_leaf = np.empty(n_leaves, dtype=np.int64)
    m.check(L.pmk_multi_leaf_pairs(m.raw, _lib.ptr(pairs_per_leaf)))

    # ---- e2e: the caller's path, host buffers in, all Nq results in the caller's host arrays ----------------------------
    e2e = None
    if not args.no_e2e:
        def step_host():
            t0 = time.perf_counter()
            m.check(L.pmk_multi_fit(m.raw, D, n_leaves, _lib.ptr(leaf_off), _lib.ptr(hXp), _lib.ptr(hyp), θ.kernel_id, _lib.ptr(kp),
                                    kp.shape[0], w["sigma2"], C.byref(bad), C.byref(info)))
            t1 = time.perf_counter()
            m.check(L.pmk_multi_query(m.raw, Nq, _lib.ptr(hXq), w["radius"], w["delta"], wθ.kernel_id, _lib.ptr(wp), wp.shape[0], 0,
                                      _lib.ptr(hYq), _lib.ptr(hVq)))
            t2 = time.perf_counter()
            return 1e3 * (t1 - t0), 1e3 * (t2 - t1)

        step_host()
        ef, eq = zip(*[step_host() for _ in range(max(1, min(args.steps, 3)))])
        ef_m, eq_m = float(np.mean(ef)), float(np.mean(eq))
        assert np.isfinite(hYq).all() and np.all(hVq >= 1e-12)
        e2e = {"value": Nq / (eq_m * 1e-3), "unit": "pts/s", "fit_leaves_per_s": n_leaves / (ef_m * 1e-3),
               "h2d_bytes_per_step": int(Xq_h.nbytes + Xp.nbytes + yp.nbytes), "d2h_bytes_per_step": int(2 * 8 * Nq),
               "ms_fit": ef_m, "ms_query": eq_m,
               "api": "pmk_multi_fit(host X, y) + pmk_multi_query(host Xq -> host Yq, Vq) on pinned host arrays; all Nq results end "
                      "in the calling process's host buffers"}

    # ---- rooflines ----------------------------------------------------------------------------------------------------
    nn = sizes.astype(np.float64)
    per_leaf = pairs_per_leaf.astype(np.float64)
    flops_pairs = float((per_leaf * (nn * nn + nn * (eval_flops(D) + 4))).sum())      # TRSM n^2 + n*(eval + mean 2 + ||s||^2 2)
    flops_chol = float((nn ** 3 / 3).sum())
    flops_fit = float((nn ** 3 / 3 + 2 * nn * nn + nn * (nn + 1) / 2 * eval_flops(D)).sum())
    h0 = m.rank_handle(0)
    peak_run = C.c_double(0.0)
    try:
        h0.check(L.pmk_measure_fp64_peak(h0.raw, C.byref(peak_run)))
        peak, peak_src = float(peak_run.value), ("measured in this run: register-resident mma.sync.m8n8k4.f64 (DMMA.8x8x4) loop on GPU 0, "
                                                 "best of 3 (pmk_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry")
    except Exception:
        peak, peak_src = fp64_peak()
    hbm_peak, hbm_src = hbm_peak_gbs()
    # per-GPU rate of the dominant kernel: all pairs' flops over the GPU-seconds the pair kernel ran on all ranks
    pair_gpu_s = float(kt[:, _lib.T_Q_PAIRS].sum()) * 1e-3
    ach = flops_pairs / pair_gpu_s / 1e12
    npairs_total = int(pairs_per_leaf.sum())

    def committed_traffic(fname):
        """DRAM bytes per launch from a committed `ncu --set full` capture (ncu cannot run inside the bench): only quoted for
        the configuration it was captured on (c3, one GPU); otherwise null."""
        if n_gpus != 1 or w["name"] != "c3":
            return None, None
        try:
            tr_ = json.load(open(os.path.join(ROOT, "profiles", fname)))
            return float(tr_["dram_bytes_read"] + tr_["dram_bytes_write"]), {"kernel": tr_["kernel"], "source": tr_["source"]}
        except Exception:
            return None, None

    traffic, traffic_of = committed_traffic("k3_traffic_r02.json")
    if traffic is None:
        traffic, traffic_of = committed_traffic("k3_traffic_r01.json")
    roofline = {"kernel": "k_query_rowp (fused cross-covariance + mean + variance, s = inv(L) kq as a row-panel product on DMMA, per (query, leaf) pair)",
                "bound": "tensor", "pipe": "FP64 DMMA.8x8x4 (mma.sync.m8n8k4.f64); tcgen05 has no f64 kind", "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "peak_source": peak_src, "traffic": traffic, "traffic_of": traffic_of,
                "algorithmic_flops_per_step": flops_pairs, "ms_per_launch_set_max_over_ranks": float(ktmax[_lib.T_Q_PAIRS]),
                "gpu_ms_all_ranks": pair_gpu_s * 1e3, "pairs_per_step": npairs_total,
                "note": "achieved = algorithmic flops of all (query, leaf) pairs / GPU-seconds of the pair kernel summed over the ranks = per-GPU rate"}
    chol_gpu_s = float(kt[:, _lib.T_FIT_CHOL].sum()) * 1e-3
    gram_gpu_s = float(kt[:, _lib.T_FIT_GRAM].sum()) * 1e-3
    chol_traffic, chol_traffic_of = committed_traffic("k_chol_traffic_r02.json")
    l_bytes = float((np.square((sizes + 31) // 32 * 32 // 8) + (sizes + 31) // 32 * 32 // 8).sum() / 2 * 512)      # packed lower tiles
    roofline_fit = {"kernel": "k_chol (batched blocked Cholesky, one leaf per CTA, DMMA trailing updates)", "bound": "tensor",
                    "achieved": flops_chol / chol_gpu_s / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": flops_chol / chol_gpu_s / 1e12 / peak,
                    "algorithmic_flops_per_step": flops_chol, "algorithmic_bytes_per_step": 2.0 * l_bytes,
                    "traffic": chol_traffic, "traffic_of": chol_traffic_of, "gpu_ms_all_ranks": chol_gpu_s * 1e3,
                    "with_gram_tiles": {"achieved": flops_fit / (chol_gpu_s + gram_gpu_s) / 1e12,
                                        "frac": flops_fit / (chol_gpu_s + gram_gpu_s) / 1e12 / peak}}
    phases = {"ranks": n_gpus,
              "max_over_ranks_ms": {"fit_pack": float(ktmax[_lib.T_FIT_PACK]), "fit_gram": float(ktmax[_lib.T_FIT_GRAM]),
                                    "fit_chol": float(ktmax[_lib.T_FIT_CHOL]), "fit_solve": float(ktmax[_lib.T_FIT_SOLVE]),
                                    "fit_refine": float(ktmax[_lib.T_FIT_REFINE]), "fit_operand_invert": float(ktmax[_lib.T_Q_INVERT]),
                                    "fit_operand_make_M": float(ktmax[_lib.T_Q_MAKE_M]), "query_tree": float(ktmax[_lib.T_Q_TREE]),
                                    "query_route_sort": float(ktmax[_lib.T_Q_ROUTE_SORT]), "query_pairs": float(ktmax[_lib.T_Q_PAIRS]),
                                    "query_combine": float(ktmax[_lib.T_Q_COMBINE])},
              "query_stages_ms": {"plan": float(mt[_lib.MT_Q_PLAN]), "route": float(mt[_lib.MT_Q_ROUTE]), "pairs": float(mt[_lib.MT_Q_PAIRS]),
                                  "return_combine": float(mt[_lib.MT_Q_RETURN])},
              "per_rank_query_pairs_ms": [float(x) for x in kt[:, _lib.T_Q_PAIRS]],
              "per_rank_fit_chol_ms": [float(x) for x in kt[:, _lib.T_FIT_CHOL]],
              "setup_ms": measure_setup(w) if not args.no_setup else None}
    # HBM side (north star: "achieved HBM GB/s for Gram build")
    try:
        gt = l_bytes / gram_gpu_s / 1e9
        ng = 8192
        Xg = np.ascontiguousarray(w["X"][:ng])
        Kg = np.empty((ng, ng), order="F")
        tg = []
        for _ in range(3):
            h0.check(L.pmk_gram(h0.raw, D, ng, _lib.ptr(Xg), θ.kernel_id, _lib.ptr(kp), kp.shape[0], 0.0, _lib.ptr(Kg)))
            tg.append(float(h0.timings()[_lib.T_GRAM]))
        g_ms = float(np.mean(tg[1:]))
        gf_ms = None
        if θ.kernel_id == 0:          # squared exponential: the same matrix with the table-driven exp (PMK_OPT_GRAM_FAST_EXP)
            h0.check(L.pmk_set_option(h0.raw, _lib.OPT_GRAM_FAST_EXP, 1))
            tg = []
            for _ in range(3):
                h0.check(L.pmk_gram(h0.raw, D, ng, _lib.ptr(Xg), θ.kernel_id, _lib.ptr(kp), kp.shape[0], 0.0, _lib.ptr(Kg)))
                tg.append(float(h0.timings()[_lib.T_GRAM]))
            h0.check(L.pmk_set_option(h0.raw, _lib.OPT_GRAM_FAST_EXP, 0))
            gf_ms = float(np.mean(tg[1:]))
        phases["hbm"] = {"peak_gbs": hbm_peak, "peak_source": hbm_src,
                         "k_gram_tiles": {"bytes_written": int(l_bytes), "gpu_ms_all_ranks": gram_gpu_s * 1e3, "gbs": gt, "frac": gt / hbm_peak},
                         "k_gram": {"what": f"constructkernelmatrix, n={ng} (8 n^2 = {8 * ng * ng >> 20} MiB written, > L2)", "ms": g_ms,
                                    "gbs": 8.0 * ng * ng / (g_ms * 1e-3) / 1e9, "frac": 8.0 * ng * ng / (g_ms * 1e-3) / 1e9 / hbm_peak,
                                    "fast_exp": None if gf_ms is None else {"ms": gf_ms, "gbs": 8.0 * ng * ng / (gf_ms * 1e-3) / 1e9,
                                                                           "frac": 8.0 * ng * ng / (gf_ms * 1e-3) / 1e9 / hbm_peak}}}
    except Exception as exc:      # an instrumentation extra: never lose the bench line to it
        phases["hbm"] = {"error": repr(exc)}
    npad_ = (sizes + 31) // 32 * 32
    cls_ = np.where(npad_ <= 512, 0, np.where(npad_ <= 768, 1, np.where(npad_ <= 1024, 2, np.where(npad_ <= 1536, 3, 4))))
    phases["leaf_points_min_mean_max"] = [int(sizes.min()), float(sizes.mean()), int(sizes.max())]
    fl_leaf = per_leaf * (nn * nn + nn * (eval_flops(D) + 4))
    phases["pairs_by_class"] = [
        {"class": c, "leaves": int((cls_ == c).sum()), "pairs": int(per_leaf[cls_ == c].sum()),
         "gpu_ms_all_ranks": float(kt[:, _lib.T_Q_PAIRS_CLASS0 + c].sum()),
         "tflops_per_gpu": float(fl_leaf[cls_ == c].sum() / max(kt[:, _lib.T_Q_PAIRS_CLASS0 + c].sum(), 1e-9) / 1e9)}
        for c in range(5) if (cls_ == c).any()]
    cond = C.c_double(0.0); sv = C.c_int(0)
    h0.check(L.pmk_condition_estimate(h0.raw, C.byref(cond), C.byref(sv)))
    phases["query_solver"] = {"in_use": int(sv.value), "cond_lower_bound_rank0": float(cond.value)}

    line = {"metric": "query_pts_per_s (mixture-GP query; fit throughput in fit_leaves_per_s)", "value": Nq / (query_ms_m * 1e-3),
            "unit": "pts/s", "fit_leaves_per_s": n_leaves / (fit_ms_m * 1e-3), "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall_ms / args.steps, "ms_fit": fit_ms_m, "ms_query": query_ms_m,
            "ms_fit_plus_query_device": fit_ms_m + query_ms_m,
            "timing": "ms_fit / ms_query: CUDA events on every rank's stream inside the library, max over the ranks, mean over the steps "
                      "(the query operand P = inv(L) is built inside the fit at every rank count); ms_per_step: host clock around the K "
                      "steps, bracketed by a synchronize of all GPUs",
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks, "gpu_launches": int(launches),
            "roofline": roofline, "roofline_fit": roofline_fit, "phases": phases}
    if e2e:
        line["e2e"] = e2e
    if n_gpus == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(w, root, sizes, leaf_off, Xp, yp)
        # parity spot check of the sample against the GPU through the multi-GPU entry point
        ns = len(cb["sample_xq"])
        Yg, Vg = np.empty(ns), np.empty(ns)
        Xs = np.ascontiguousarray(cb["sample_xq"])
        m.check(L.pmk_multi_query(m.raw, ns, _lib.ptr(Xs), w["radius"], w["delta"], wθ.kernel_id, _lib.ptr(wp), wp.shape[0], 0,
                                  _lib.ptr(Yg), _lib.ptr(Vg)))
        sc_y, sc_v = np.sqrt(np.mean(cb["sample_y"] ** 2)), np.sqrt(np.mean(cb["sample_v"] ** 2))
        line["cpu_baseline"] = {"value": cb["query_pts_per_s"], "unit": "pts/s", "fit_leaves_per_s": cb["fit_leaves_per_s"],
                                "cores": cb["cores"], "kind": "port", "sample": cb["sample"],
                                "parity_vs_gpu": {"mean_max_err_over_rms": float(np.abs(Yg - cb["sample_y"]).max() / sc_y),
                                                  "var_max_err_over_rms": float(np.abs(Vg - cb["sample_v"]).max() / sc_v)}}
    m.close()
    return line


if __name__ == "__main__":
    main()
