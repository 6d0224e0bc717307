// Query-side BSP work on the device: home leaf (findpartition), neighbour leaves
// (findneighbourpartitions), mixture weights, pair-list construction and the final convex combine.
//
// Replaces reference src/patchwork/partition.jl:248-262, src/RKHS/mixtureGP.jl:339-405 and the
// weight / normalise / combine lines of querymixtureGP! (mixtureGP.jl:224-272).
// Every floating-point value that feeds a comparison is computed with the oracle's operation order
// (explicit __dmul_rn/__dadd_rn, no FMA contraction), so leaf ids and neighbour lists are bit-exact.
#include "pmk_internal.cuh"

namespace pmk {

template <int D>
__device__ __forceinline__ int descend(const TreeDev& tr, const double* x) {
  const int Lv = tr.levels - 1;
  int node = 0, leaf = 0;
  for (int d = 0; d < Lv; ++d) {
    double v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = tr.hv[k * tr.n_hp + node];
    const double h = dot_seq<D>(v, x);
    const int right = !(h < tr.hc[node]);                 // partition.jl:254: dot(v,x) < c ? left : right
    leaf = leaf * 2 + right;
    node += right ? (1 << (Lv - 1 - d)) : 1;
  }
  return leaf + 1;
}

template <int D>
__global__ void k_home(TreeDev tr, int64_t Nq, const double* __restrict__ Xq, int32_t* __restrict__ home,
                       int32_t* __restrict__ leaf_qcount /* may be null */) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= Nq) return;
  double x[D];
#pragma unroll
  for (int d = 0; d < D; ++d) x[d] = Xq[j * D + d];
  const int h = descend<D>(tr, x);
  home[j] = h;
  if (leaf_qcount) atomicAdd(&leaf_qcount[h - 1], 1);
}

// ---------------------------------------------------------------------------------------------
// Exact pruning of findneighbourpartitions' scan over ALL hyperplanes (mixtureGP.jl:354).
// A hyperplane i can only pass the reference's test norm(z - p) < radius for a query p of leaf l if
// its plane comes within ~radius of the bounding box of l's queries.  Per leaf we therefore build the
// ascending list of such hyperplanes (interval arithmetic with a generous slack), and every query then
// runs the reference's test -- same operations, same order -- on its home leaf's list only.  The kept
// set, its order and every t are identical to the full scan (checked against the brute-force kernel
// in tests/test_gpu_parity.py::test_pruned_neighbour_search_equals_full_scan).
template <int D>
__global__ void k_leaf_bbox(int64_t Nq, const double* __restrict__ Xq, const int32_t* __restrict__ qperm,
                            const int64_t* __restrict__ leaf_qstart, double* __restrict__ bbox /* [leaf][2][D] */) {
  const int leaf = blockIdx.x;
  const int64_t a = leaf_qstart[leaf], b = leaf_qstart[leaf + 1];
  double lo[D], hi[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { lo[d] = INFINITY; hi[d] = -INFINITY; }
  for (int64_t k = a + threadIdx.x; k < b; k += blockDim.x) {
    const int64_t j = qperm[k];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double x = Xq[j * D + d];
      lo[d] = fmin(lo[d], x);
      hi[d] = fmax(hi[d], x);
    }
  }
  __shared__ double red[2 * D][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int d = 0; d < D; ++d) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], m));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], m));
    }
    if (lane == 0) { red[d][warp] = lo[d]; red[D + d][warp] = hi[d]; }
  }
  __syncthreads();
  if (threadIdx.x < D) {
    double l = INFINITY, h = -INFINITY;
    for (int w = 0; w < nw; ++w) { l = fmin(l, red[threadIdx.x][w]); h = fmax(h, red[D + threadIdx.x][w]); }
    bbox[(leaf * 2 + 0) * D + threadIdx.x] = l;
    bbox[(leaf * 2 + 1) * D + threadIdx.x] = h;
  }
}

template <int D, bool FILL>
__global__ void __launch_bounds__(256)
k_leaf_candidates(TreeDev tr, const double* __restrict__ bbox, double radius, int32_t* __restrict__ cand_count,
                  const int64_t* __restrict__ cand_start, int32_t* __restrict__ cand) {
  const int leaf = blockIdx.x;
  double lo[D], hi[D];
  double amax = 0.0;
  bool empty = false;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    lo[d] = bbox[(leaf * 2 + 0) * D + d];
    hi[d] = bbox[(leaf * 2 + 1) * D + d];
    if (!(lo[d] <= hi[d])) empty = true;
    amax += fmax(fabs(lo[d]), fabs(hi[d]));
  }
  if (empty) {
    if (!FILL && threadIdx.x == 0) cand_count[leaf] = 0;
    return;
  }
  __shared__ int wsum[8];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t out0 = FILL ? cand_start[leaf] : 0;
  for (int base = 0; base < tr.n_hp; base += 256) {
    const int i = base + threadIdx.x;
    bool f = false;
    if (i < tr.n_hp) {
      double slo = 0.0, shi = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double u = tr.hv[d * tr.n_hp + i];
        const double a = u * lo[d], b = u * hi[d];
        slo += fmin(a, b);
        shi += fmax(a, b);
      }
      const double c = tr.hc[i];
      const double dist = fmax(0.0, fmax(c - shi, slo - c));       // min |c - u.p| over the box
      f = dist <= radius * (1.0 + 1e-6) + 1e-9 * (1.0 + fabs(c) + amax);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int wpre = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (FILL && f) cand[out0 + off + wpre] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += wsum[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (!FILL && threadIdx.x == 0) cand_count[leaf] = s_base;
}

// Pass FILL=false counts the slots of every query (kept neighbours + 1); FILL=true writes them.
// Slot order = reference order: kept hyperplanes in increasing index, home leaf last.
// The reference's per-hyperplane test (mixtureGP.jl:357-399) for hyperplane i; returns the neighbour leaf or 0.
template <int D>
__device__ __forceinline__ int neighbour_test(const TreeDev& tr, const double* p, int home, int i, double radius,
                                              double delta, double* t_out) {
  double u[D];
#pragma unroll
  for (int d = 0; d < D; ++d) u[d] = tr.hv[d * tr.n_hp + i];
  const double c = tr.hc[i];
  const double t = __dadd_rn(-dot_seq<D>(u, p), c);              // mixtureGP.jl:361  t = -dot(u,p) + c
  // z = p + t.*u ; norm(z - p)                                     mixtureGP.jl:362,367
  double s = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double z = __dadd_rn(p[d], __dmul_rn(t, u[d]));
    const double dd = __dsub_rn(z, p[d]);
    s = (d == 0) ? __dmul_rn(dd, dd) : __dadd_rn(s, __dmul_rn(dd, dd));
  }
  if (!(__dsqrt_rn(s) < radius)) return 0;
  double z1[D], z2[D];
  const double tp = __dadd_rn(t, delta), tm = __dsub_rn(t, delta);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    z1[d] = __dadd_rn(p[d], __dmul_rn(tp, u[d]));                // mixtureGP.jl:370-371
    z2[d] = __dadd_rn(p[d], __dmul_rn(tm, u[d]));
  }
  const int r1 = descend<D>(tr, z1);
  const int r2 = descend<D>(tr, z2);
  if ((r2 == home) == (r1 == home)) return 0;                    // mixtureGP.jl:387 xor
  *t_out = t;
  return (r1 == home) ? r2 : r1;                                 // mixtureGP.jl:392-395
}

// Pass FILL=false counts the slots of every query (kept neighbours + 1); FILL=true writes them.
// Slot order = reference order: kept hyperplanes in increasing index, home leaf last.
// PRUNED: thread k handles query qperm[k] (queries sorted by home leaf, so a warp shares one candidate
// list) and scans its leaf's candidate hyperplanes; otherwise thread j scans all hyperplanes.
template <int D, bool FILL, bool PRUNED>
__global__ void k_neighbours(TreeDev tr, QueryPlan q, double radius, double delta, int wkind, double wparam,
                             int32_t* __restrict__ leaf_count /* pairs per global leaf */, const int32_t* __restrict__ qperm,
                             const int64_t* __restrict__ cand_start, const int32_t* __restrict__ cand) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= q.Nq) return;
  const int64_t j = PRUNED ? (int64_t)qperm[k] : k;
  double p[D];
#pragma unroll
  for (int d = 0; d < D; ++d) p[d] = q.Xq[j * D + d];
  const int home = q.home[j];
  int64_t slot = FILL ? q.pair_off[j] : 0;
  int kept = 0;
  const int64_t c0 = PRUNED ? cand_start[home - 1] : 0;
  const int64_t c1 = PRUNED ? cand_start[home] : tr.n_hp;
  for (int64_t cc = c0; cc < c1; ++cc) {
    const int i = PRUNED ? cand[cc] : (int)cc;
    double t;
    const int nb = neighbour_test<D>(tr, p, home, i, radius, delta, &t);
    if (nb != 0) {
      if (FILL) {
        q.pair_leaf[slot] = nb;
        q.pair_q[slot] = (int32_t)j;
        q.pair_hp[slot] = i + 1;
        q.pair_t[slot] = t;
        q.pair_w[slot] = k_tau(wkind, wparam, fabs(t));            // mixtureGP.jl:231
        atomicAdd(&leaf_count[nb - 1], 1);
        ++slot;
      }
      ++kept;
    }
  }
  if (FILL) {
    q.pair_leaf[slot] = home;                                       // mixtureGP.jl:237-239, w[end] = 1
    q.pair_q[slot] = (int32_t)j;
    q.pair_hp[slot] = 0;
    q.pair_t[slot] = 0.0;
    q.pair_w[slot] = 1.0;
    atomicAdd(&leaf_count[home - 1], 1);
  } else {
    q.npairs[j] = kept + 1;
  }
}

// Pruned search, one CTA per home leaf (scan pass).  The leaf's candidate hyperplanes (u, c) and the
// hyperplanes of its ANCESTORS are staged in shared memory once.  Per query:
//   A. every candidate gets the cheap conservative rejection |t| > radius(1+1e-9)+... (norm(z - p)
//      equals |t| up to a few ulp of |p|, so the reference's test would fail as well); survivors are
//      only recorded, so the warp does not diverge into the expensive path per candidate;
//   B. each survivor runs the reference's exact test.  "findpartition(z) == home" is evaluated as
//      "z takes home's branch at each of home's ancestors" -- the same comparisons findpartition makes
//      on the way to home, but as independent shared-memory dot products instead of two dependent
//      pointer-chasing descents.
// The kept candidates (slot numbers in the leaf's list, at most 7, else an overflow mark) go to a
// 16-byte record per query; the fill pass replays only those.
static constexpr int kKeptMax = 7;
static constexpr int kCandCap = 1024;     // planes staged in shared memory per leaf ((D + 2) x 8 KB; C4's 3-D leaves have ~800, and the ones beyond the cap cost three global loads per query)

template <int D>
__global__ void __launch_bounds__(256)
k_neighbours_scan(TreeDev tr, QueryPlan q, double radius, double delta, const int32_t* __restrict__ qperm,
                  const int64_t* __restrict__ leaf_qstart, const int64_t* __restrict__ cand_start,
                  const int32_t* __restrict__ cand, uint16_t* __restrict__ kept_rec /* [Nq][8] */) {
  __shared__ double s_u[D][kCandCap];
  __shared__ double s_c[kCandCap];
  __shared__ double a_u[D][32];
  __shared__ double a_c[32];
  __shared__ double a_n[32];                 // |u| of home's ancestors
  __shared__ double s_n[kCandCap];           // |u| of the candidates
  __shared__ unsigned long long s_umax;      // largest |u| among the leaf's candidates (bit pattern of a positive double)
  const int leaf = blockIdx.x;
  const int64_t qa = leaf_qstart[leaf], qb = leaf_qstart[leaf + 1];
  if (qa == qb) return;
  const int64_t ca = cand_start[leaf];
  const int nc = (int)(cand_start[leaf + 1] - ca);
  const int ns = nc < kCandCap ? nc : kCandCap;
  const int Lv = tr.levels - 1;
  if (threadIdx.x == 0) s_umax = 0ull;
  __syncthreads();
  {
    double umax = 0.0;
    for (int k = threadIdx.x; k < nc; k += blockDim.x) {
      const int i = cand[ca + k];
      double nn = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double ud = tr.hv[d * tr.n_hp + i];
        if (k < ns) s_u[d][k] = ud;
        nn += ud * ud;
      }
      const double un = sqrt(nn);
      if (k < ns) {
        s_c[k] = tr.hc[i];
        s_n[k] = un;
      }
      umax = fmax(umax, un);
    }
    atomicMax(&s_umax, (unsigned long long)__double_as_longlong(umax));
  }
  if (threadIdx.x == 0) {        // home's ancestors: node_0 = 0, node_{d+1} = node_d + (right ? 2^(Lv-1-d) : 1)
    int node = 0;
    for (int d = 0; d < Lv; ++d) {
      a_c[d] = tr.hc[node];
      double nn = 0.0;
      for (int dd = 0; dd < D; ++dd) {
        a_u[dd][d] = tr.hv[dd * tr.n_hp + node];
        nn += a_u[dd][d] * a_u[dd][d];
      }
      a_n[d] = sqrt(nn);
      const int right = (leaf >> (Lv - 1 - d)) & 1;
      node += right ? (1 << (Lv - 1 - d)) : 1;
    }
  }
  __syncthreads();
  // findpartition(x1) == home  XOR  findpartition(x2) == home, both descents in one walk over home's ancestors
  // (same comparisons as partition.jl:254; a point is in home iff it takes home's branch at every ancestor)
  auto exactly_one_in_home = [&](const double* x1, const double* x2) -> bool {
    bool in1 = true, in2 = true;
    // deepest ancestor first: the probes are within radius + delta of a query of home, so if they are outside home it is one of the
    // cell's own (deep) boundary planes that says so, and the walk ends after a step or two instead of ten (the conjunction does
    // not depend on the order; ncu: this loop was 57 % of the kernel's instructions, nine lanes active on average)
    for (int d = Lv - 1; d >= 0; --d) {
      double v[D];
#pragma unroll
      for (int dd = 0; dd < D; ++dd) v[dd] = a_u[dd][d];
      const double c = a_c[d];
      const int hb = (leaf >> (Lv - 1 - d)) & 1;
      in1 = in1 && ((int)!(dot_seq<D>(v, x1) < c) == hb);
      in2 = in2 && ((int)!(dot_seq<D>(v, x2) < c) == hb);
      if (!in1 && !in2) return false;
    }
    return in1 != in2;
  };
  // Every warp walks its queries 32 at a time with ALL lanes in the loop (lanes past the end are masked): the cheap |t| test
  // over the candidate list runs in lock step, the survivors are buffered per lane, and the expensive exact test (square
  // root, two 12-level ancestor checks) is then replayed survivor index by survivor index with the warp reconverged at
  // every step.  (Processing the survivors where they were found ran the exact test one lane at a time: ncu showed 94 % of
  // the executed instructions with fewer than 4 active threads and 54 survivors per query on C4.)
  for (int64_t kq0 = qa + (threadIdx.x & ~31); kq0 < qb; kq0 += blockDim.x) {
    const int64_t kq = kq0 + (threadIdx.x & 31);
    const bool valid = kq < qb;
    const int64_t j = valid ? qperm[kq] : qperm[qb - 1];
    double p[D];
    double pabs = 1.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      p[d] = q.Xq[j * D + d];
      pabs += fabs(p[d]);
    }
    const double thr = radius * (1.0 + 1e-9) + 1e-12 * pabs;
    // distance from the query to the nearest boundary of its home cell (home's ancestors' planes).  A candidate plane that
    // is closer than that, by more than delta, projects the query to a point whose two probes both lie strictly inside the
    // home cell: findpartition gives home for both, the reference's xor test fails, the hyperplane is not kept -- no need
    // to run the two descents.  (Slack 1e-9 * (1 + |p|_1): six orders above the rounding of the dot products, four below
    // delta; whatever falls inside the slack takes the exact test.)
    // ... and which of home's ancestors are NEAR the query: a probe of a buffered candidate is p + (t +- delta) u with |t| <= thr,
    // i.e. at most (thr + |delta|) |u| away from p, so an ancestor plane farther from p than that (plus slack) has BOTH probes of
    // EVERY buffered candidate on home's side -- only the near ones (typically one to three of twelve) can decide the xor test.
    double mrel = 1e300;
    const double near_r = (thr + fabs(delta)) * __longlong_as_double((long long)s_umax) * (1.0 + 1e-9) + 1e-9 * pabs;
    unsigned long long near = 0ull;
    for (int d = 0; d < Lv; ++d) {
      double v[D];
#pragma unroll
      for (int dd = 0; dd < D; ++dd) v[dd] = a_u[dd][d];
      const double dist = fabs(dot_seq<D>(v, p) - a_c[d]) / a_n[d];
      mrel = fmin(mrel, dist);
      if (!(dist > near_r)) near |= 1ull << d;
    }
    mrel -= 1e-9 * pabs;
    unsigned short surv[64], surv2[16];
    int nsurv = 0, nsurv2 = 0, kept = 0;
    unsigned short kl[kKeptMax];
    auto load_plane = [&](int k, double (&u)[D], double& c, double& un) {
      if (k < ns) {
        c = s_c[k];
        un = s_n[k];
#pragma unroll
        for (int d = 0; d < D; ++d) u[d] = s_u[d][k];
      } else {
        const int i = cand[ca + k];
        c = tr.hc[i];
        un = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          u[d] = tr.hv[d * tr.n_hp + i];
          un += u[d] * u[d];
        }
        un = sqrt(un);
      }
    };
    // Stage 1 (cheap, exact-safe): can candidate k's xor test come out true at all?  Directional form of the "both probes inside
    // home" argument, ancestor by ancestor (near ones only, deepest first): with f_a(x) = u_a . x - c_a,
    // f_a(p + (t +- delta) u) = f_a(p) + t g +- delta g, g = u_a . u.  If the midpoint value is on home's side by more than
    // |delta g| + slack for every near ancestor, both probes are in home; if it is on the other side by more than that for one
    // ancestor, neither is: the reference's xor (mixtureGP.jl:387) is false either way and the exact test is not needed.  Whatever
    // falls inside a band -- a plane of home's own boundary at the projected point, above all -- goes on to stage 2.
    auto may_keep = [&](int k) -> bool {
      double u[D], c, un;
      load_plane(k, u, c, un);
      const double t = __dadd_rn(-dot_seq<D>(u, p), c);            // mixtureGP.jl:361
      bool all_in = true;
      for (unsigned long long m = near; m != 0ull;) {
        const int d = 63 - __clzll((long long)m);
        m &= ~(1ull << d);
        double v[D];
#pragma unroll
        for (int dd = 0; dd < D; ++dd) v[dd] = a_u[dd][d];
        const double gd = dot_seq<D>(v, u);
        const double fz = fma(t, gd, dot_seq<D>(v, p) - a_c[d]);
        const double band = fabs(delta * gd) + 1e-9 * pabs * (a_n[d] + 1.0);
        const double sm = ((leaf >> (Lv - 1 - d)) & 1) ? fz : -fz;     // > 0: home's side of ancestor d
        if (sm < -band) return false;
        all_in = all_in && (sm > band);
      }
      return !all_in;
    };
    // Stage 2: the reference's exact test.
    auto exact = [&](int k) {
      double u[D], c, un;
      load_plane(k, u, c, un);
      const double t = __dadd_rn(-dot_seq<D>(u, p), c);            // mixtureGP.jl:361
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double z = __dadd_rn(p[d], __dmul_rn(t, u[d]));      // mixtureGP.jl:362
        const double dd = __dsub_rn(z, p[d]);
        s = (d == 0) ? __dmul_rn(dd, dd) : __dadd_rn(s, __dmul_rn(dd, dd));
      }
      if (!(__dsqrt_rn(s) < radius)) return;                       // mixtureGP.jl:367
      double z1[D], z2[D];
      const double tp = __dadd_rn(t, delta), tm = __dsub_rn(t, delta);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        z1[d] = __dadd_rn(p[d], __dmul_rn(tp, u[d]));              // mixtureGP.jl:370-371
        z2[d] = __dadd_rn(p[d], __dmul_rn(tm, u[d]));
      }
      if (exactly_one_in_home(z1, z2)) {                           // mixtureGP.jl:387 xor
        if (kept < kKeptMax) kl[kept] = (unsigned short)k;
        ++kept;
      }
    };
    // Both stages are replayed entry index by entry index with the warp reconverged at every step, each from its own per-lane
    // list (ascending k throughout, so the kept list stays in the reference's order).  Stage 1 thins 25-30 radius survivors per
    // query to the one or two planes that bound home near the query; stage 2 then runs with most of its lanes busy (with one list
    // the exact walk ran at two active lanes of 32: ncu, profiles/ncu_r02_neighbours_scan_c3.csv).
    auto flush2 = [&]() {
      const int nmax = __reduce_max_sync(0xffffffffu, nsurv2);
      for (int s_ = 0; s_ < nmax; ++s_) {
        __syncwarp();
        if (s_ < nsurv2) exact(surv2[s_]);
      }
      nsurv2 = 0;
    };
    auto flush = [&]() {
      const int nmax = __reduce_max_sync(0xffffffffu, nsurv);
      for (int s_ = 0; s_ < nmax; ++s_) {
        __syncwarp();
        if (__any_sync(0xffffffffu, nsurv2 == 16)) flush2();
        if (s_ < nsurv && may_keep(surv[s_])) surv2[nsurv2++] = surv[s_];
      }
      nsurv = 0;
    };
    // The candidate scan proper: |t| against the radius (conservative form), and the isotropic "closer than the nearest boundary of
    // the home cell" rejection (both probes inside home) for those within it.  A list can grow by one per candidate: with the check
    // every 16 candidates nobody passes 48 + 16 entries.
    for (int k = 0; k < ns; ++k) {
      double u[D];
#pragma unroll
      for (int d = 0; d < D; ++d) u[d] = s_u[d][k];
      const double tq = __dadd_rn(-dot_seq<D>(u, p), s_c[k]);
      if ((k & 15) == 0 && __any_sync(0xffffffffu, nsurv > 48)) flush();
      if (valid && !(fabs(tq) > thr)) {
        if (!(s_n[k] * (fabs(tq) + fabs(delta)) * (1.0 + 1e-9) < mrel)) surv[nsurv++] = (unsigned short)k;
      }
    }
    for (int k = ns; k < nc; ++k) {          // candidates beyond the shared-memory list (very large radius only)
      double u[D], c, un;
      load_plane(k, u, c, un);
      const double tq = __dadd_rn(-dot_seq<D>(u, p), c);
      if (__any_sync(0xffffffffu, nsurv == 64)) flush();
      if (valid && !(fabs(tq) > thr) && !(un * (fabs(tq) + fabs(delta)) * (1.0 + 1e-9) < mrel)) surv[nsurv++] = (unsigned short)k;
    }
    flush();
    flush2();
    if (valid) {
      q.npairs[j] = kept + 1;
      uint16_t* rec = kept_rec + j * 8;
#pragma unroll
      for (int m = 0; m < kKeptMax; ++m) rec[m] = m < kept ? kl[m] : (uint16_t)0;
      rec[7] = kept <= kKeptMax ? (uint16_t)kept : (uint16_t)0xFFFF;
    }
  }
}

// Fill pass: replay the kept candidates of every query (the reference's exact test again, now with the real
// descents that also yield the neighbour leaf id) and write the pair slots in reference order.
template <int D>
__global__ void k_neighbours_fill(TreeDev tr, QueryPlan q, double radius, double delta, int wkind, double wparam,
                                  int32_t* __restrict__ leaf_count, const int64_t* __restrict__ cand_start,
                                  const int32_t* __restrict__ cand, const uint16_t* __restrict__ kept_rec) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= q.Nq) return;
  double p[D];
#pragma unroll
  for (int d = 0; d < D; ++d) p[d] = q.Xq[j * D + d];
  const int home = q.home[j];
  int64_t slot = q.pair_off[j];
  const uint16_t* rec = kept_rec + j * 8;
  const int64_t ca = cand_start[home - 1];
  const int nk = rec[7];
  const int n_iter = nk == 0xFFFF ? (int)(cand_start[home] - ca) : nk;
  for (int m = 0; m < n_iter; ++m) {
    const int k = nk == 0xFFFF ? m : (int)rec[m];
    const int i = cand[ca + k];
    double t;
    const int nb = neighbour_test<D>(tr, p, home, i, radius, delta, &t);
    if (nb != 0) {
      q.pair_leaf[slot] = nb;
      q.pair_q[slot] = (int32_t)j;
      q.pair_hp[slot] = i + 1;
      q.pair_t[slot] = t;
      q.pair_w[slot] = k_tau(wkind, wparam, fabs(t));              // mixtureGP.jl:231
      atomicAdd(&leaf_count[nb - 1], 1);
      ++slot;
    }
  }
  q.pair_leaf[slot] = home;                                         // mixtureGP.jl:237-239, w[end] = 1
  q.pair_q[slot] = (int32_t)j;
  q.pair_hp[slot] = 0;
  q.pair_t[slot] = 0.0;
  q.pair_w[slot] = 1.0;
  atomicAdd(&leaf_count[home - 1], 1);
}

// ε-overlap training sets on the device: findεpartitions! (partition.jl:269-298) per training point -- left iff
// v.x < c + ε, right iff v.x > c - ε, both possible -- as an explicit-stack DFS (left first, so a point's leaves
// come out in the reference's left-to-right order).  FILL=false counts, FILL=true writes (leaf, point) pairs at the
// point's offset; a stable sort by leaf then yields every leaf's point list in ascending global index, which is
// exactly X_set_inds of organizetrainingsets (partition.jl:301-357).
template <int D, bool FILL>
__global__ void k_eps_partitions(TreeDev tr, int64_t N, const double* __restrict__ X, double eps,
                                 int32_t* __restrict__ counts, const int64_t* __restrict__ off,
                                 int32_t* __restrict__ pair_leaf, int32_t* __restrict__ pair_pt,
                                 int32_t* __restrict__ leaf_count) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  double x[D];
#pragma unroll
  for (int d = 0; d < D; ++d) x[d] = X[n * D + d];
  const int Lv = tr.levels - 1;
  int st_node[26], st_prefix[26];
  signed char st_depth[26];
  int sp = 0;
  st_node[0] = 0; st_prefix[0] = 0; st_depth[0] = 0; sp = 1;
  int cnt = 0;
  int64_t slot = FILL ? off[n] : 0;
  while (sp > 0) {
    --sp;
    const int node = st_node[sp], prefix = st_prefix[sp], depth = st_depth[sp];
    if (depth == Lv) {
      if (FILL) {
        pair_leaf[slot] = prefix + 1;
        pair_pt[slot] = (int32_t)n + 1;
        atomicAdd(&leaf_count[prefix], 1);
        ++slot;
      }
      ++cnt;
      continue;
    }
    double v[D];
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = tr.hv[d * tr.n_hp + node];
    const double h = dot_seq<D>(v, x);
    const double c = tr.hc[node];
    const bool go_l = h < __dadd_rn(c, eps);       // partition.jl:287
    const bool go_r = h > __dsub_rn(c, eps);       // partition.jl:292
    if (go_r) {                                     // pushed first, popped last: left subtree is visited first
      st_node[sp] = node + (1 << (Lv - 1 - depth));
      st_prefix[sp] = prefix * 2 + 1;
      st_depth[sp] = (signed char)(depth + 1);
      ++sp;
    }
    if (go_l) {
      st_node[sp] = node + 1;
      st_prefix[sp] = prefix * 2;
      st_depth[sp] = (signed char)(depth + 1);
      ++sp;
    }
  }
  if (!FILL) counts[n] = cnt;
}

// Yq = dot(w,u), Vq = dot(w, v.*w) with w = w_tilde / sum(w_tilde)   (mixtureGP.jl:263-272)
__global__ void k_combine(int64_t Nq, const int64_t* __restrict__ pair_off, const double* __restrict__ pw,
                          const double* __restrict__ pu, const double* __restrict__ pv, double* __restrict__ Yq,
                          double* __restrict__ Vq, int mean_only) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= Nq) return;
  const int64_t a = pair_off[j], b = pair_off[j + 1];
  double sw = 0.0;
  for (int64_t s = a; s < b; ++s) sw = __dadd_rn(sw, pw[s]);
  double y = 0.0, v = 0.0;
  for (int64_t s = a; s < b; ++s) {
    const double w = __ddiv_rn(pw[s], sw);
    y = __dadd_rn(y, __dmul_rn(w, pu[s]));
    if (!mean_only) v = __dadd_rn(v, __dmul_rn(w, __dmul_rn(pv[s], w)));
  }
  Yq[j] = y;
  if (!mean_only) Vq[j] = v;
}

// single-block exclusive scan of leaf pair counts -> leaf pair starts (n <= a few 10^4)
__global__ void k_scan_small(const int32_t* __restrict__ in, int64_t* __restrict__ out, int n) {
  __shared__ int64_t carry;
  __shared__ int64_t buf[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    int64_t v = i < n ? in[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      int64_t add = threadIdx.x >= off ? buf[threadIdx.x - off] : 0;
      __syncthreads();
      buf[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < n) out[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}


// ---------------------------------------------------------------------------------------------
// Sub-tree ownership: the pairs of a plan, in leaf-sorted order, as what travels to the owners of the leaves --
// every pair's query point (D doubles) and its global leaf id.  Leaf-sorted order makes each owner's share one contiguous
// segment (owners hold contiguous leaf ranges).
template <int D>
__global__ void k_pack_sorted_pairs(QueryPlan q, const int32_t* __restrict__ sorted_pair, double* __restrict__ Xs,
                                    int32_t* __restrict__ leaf_s) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= q.n_pairs) return;
  const int32_t id = sorted_pair[j];
  const int64_t qi = q.pair_q[id];
#pragma unroll
  for (int d = 0; d < D; ++d) Xs[j * D + d] = q.Xq[qi * D + d];
  leaf_s[j] = q.pair_leaf[id];
}

// the owners' answers, received in leaf-sorted order, back to pair-id order
__global__ void k_unpack_sorted_pairs(int64_t n, const int32_t* __restrict__ sorted_pair, const double* __restrict__ us,
                                      const double* __restrict__ vs, double* __restrict__ pu, double* __restrict__ pv) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int32_t id = sorted_pair[j];
  pu[id] = us[j];
  if (vs && pv) pv[id] = vs[j];
}

// start of every global leaf's run in an ascending key list (keys = 1-based leaf ids): start[g] = first index with key > g,
// g = 0 .. TL (binary search, one thread per leaf); *foreign is raised when a key lies outside the owned leaf range.
__global__ void k_run_starts(const int32_t* __restrict__ keys, int64_t R, int64_t TL, int64_t leaf_base, int64_t n_own,
                             int64_t* __restrict__ start, int* __restrict__ foreign) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g > TL) return;
  int64_t lo = 0, hi = R;               // first index whose key >= g + 1
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)keys[mid] < g + 1) lo = mid + 1;
    else hi = mid;
  }
  start[g] = lo;
  if (g == 0 && R > 0) {
    const int64_t first = (int64_t)keys[0] - 1, last = (int64_t)keys[R - 1] - 1;
    if (first < leaf_base || last >= leaf_base + n_own) *foreign = 1;
  }
}

// debug_flag outputs of findneighbourpartitions for EVERY hyperplane (mixtureGP.jl:347-352,364-365,389 -> debug_vars.ts_set,
// zs_set, hps_keep_flags_set, :256-258): t_i, z_i = p + t_i u_i and the keep flag, dense Nq x n_hp.  One thread per (query,
// hyperplane); same arithmetic as the search itself (neighbour_test).  Small Nq only: the caller bounds Nq * n_hp.
template <int D>
__global__ void k_dense_debug(TreeDev tr, QueryPlan q, double radius, double delta, uint8_t* __restrict__ keep,
                              double* __restrict__ ts, double* __restrict__ zs) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= q.Nq * tr.n_hp) return;
  const int64_t j = idx / tr.n_hp;
  const int i = (int)(idx % tr.n_hp);
  double p[D], u[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    p[d] = q.Xq[j * D + d];
    u[d] = tr.hv[d * tr.n_hp + i];
  }
  const double t = __dadd_rn(-dot_seq<D>(u, p), tr.hc[i]);       // mixtureGP.jl:361
  if (ts) ts[idx] = t;
  if (zs) {
#pragma unroll
    for (int d = 0; d < D; ++d) zs[idx * D + d] = __dadd_rn(p[d], __dmul_rn(t, u[d]));   // mixtureGP.jl:362
  }
  if (keep) {
    double t2;
    keep[idx] = neighbour_test<D>(tr, p, q.home[j], i, radius, delta, &t2) != 0 ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------------
void launch_home(int D, const TreeDev& tr, int64_t Nq, const double* dXq, int32_t* d_home, int32_t* d_leaf_qcount,
                 cudaStream_t s) {
  const int T = 256;
  const unsigned B = (unsigned)((Nq + T - 1) / T);
  if (B == 0) return;
  switch (D) {
    case 1: k_home<1><<<B, T, 0, s>>>(tr, Nq, dXq, d_home, d_leaf_qcount); break;
    case 2: k_home<2><<<B, T, 0, s>>>(tr, Nq, dXq, d_home, d_leaf_qcount); break;
    case 3: k_home<3><<<B, T, 0, s>>>(tr, Nq, dXq, d_home, d_leaf_qcount); break;
    default: break;
  }
}

void launch_leaf_bbox(int D, int n_leaves, int64_t Nq, const double* dXq, const int32_t* qperm, const int64_t* leaf_qstart,
                      double* bbox, cudaStream_t s) {
  switch (D) {
    case 1: k_leaf_bbox<1><<<n_leaves, 128, 0, s>>>(Nq, dXq, qperm, leaf_qstart, bbox); break;
    case 2: k_leaf_bbox<2><<<n_leaves, 128, 0, s>>>(Nq, dXq, qperm, leaf_qstart, bbox); break;
    case 3: k_leaf_bbox<3><<<n_leaves, 128, 0, s>>>(Nq, dXq, qperm, leaf_qstart, bbox); break;
    default: break;
  }
}

template <bool FILL>
static void launch_cand_t(int D, int n_leaves, const TreeDev& tr, const double* bbox, double radius, int32_t* cand_count,
                          const int64_t* cand_start, int32_t* cand, cudaStream_t s) {
  switch (D) {
    case 1: k_leaf_candidates<1, FILL><<<n_leaves, 256, 0, s>>>(tr, bbox, radius, cand_count, cand_start, cand); break;
    case 2: k_leaf_candidates<2, FILL><<<n_leaves, 256, 0, s>>>(tr, bbox, radius, cand_count, cand_start, cand); break;
    case 3: k_leaf_candidates<3, FILL><<<n_leaves, 256, 0, s>>>(tr, bbox, radius, cand_count, cand_start, cand); break;
    default: break;
  }
}

void launch_leaf_candidates(int D, bool fill, int n_leaves, const TreeDev& tr, const double* bbox, double radius,
                            int32_t* cand_count, const int64_t* cand_start, int32_t* cand, cudaStream_t s) {
  if (fill) launch_cand_t<true>(D, n_leaves, tr, bbox, radius, cand_count, cand_start, cand, s);
  else launch_cand_t<false>(D, n_leaves, tr, bbox, radius, cand_count, cand_start, cand, s);
}

template <bool FILL, bool PRUNED>
static void launch_nb_t(int D, const TreeDev& tr, const QueryPlan& q, double radius, double delta, int wkind,
                        double wparam, int32_t* d_leaf_count, const int32_t* qperm, const int64_t* cand_start,
                        const int32_t* cand, cudaStream_t s) {
  const int T = 128;
  const unsigned B = (unsigned)((q.Nq + T - 1) / T);
  if (B == 0) return;
  switch (D) {
    case 1: k_neighbours<1, FILL, PRUNED><<<B, T, 0, s>>>(tr, q, radius, delta, wkind, wparam, d_leaf_count, qperm, cand_start, cand); break;
    case 2: k_neighbours<2, FILL, PRUNED><<<B, T, 0, s>>>(tr, q, radius, delta, wkind, wparam, d_leaf_count, qperm, cand_start, cand); break;
    case 3: k_neighbours<3, FILL, PRUNED><<<B, T, 0, s>>>(tr, q, radius, delta, wkind, wparam, d_leaf_count, qperm, cand_start, cand); break;
    default: break;
  }
}

void launch_neighbours(int D, bool fill, bool pruned, int n_leaves, const TreeDev& tr, const QueryPlan& q, double radius,
                       double delta, int wkind, double wparam, int32_t* d_leaf_count, const int32_t* qperm,
                       const int64_t* leaf_qstart, const int64_t* cand_start, const int32_t* cand, uint16_t* kept_rec,
                       cudaStream_t s) {
  if (!pruned) {
    if (fill) launch_nb_t<true, false>(D, tr, q, radius, delta, wkind, wparam, d_leaf_count, qperm, cand_start, cand, s);
    else launch_nb_t<false, false>(D, tr, q, radius, delta, wkind, wparam, d_leaf_count, qperm, cand_start, cand, s);
    return;
  }
  if (!fill) {
    switch (D) {
      case 1: k_neighbours_scan<1><<<n_leaves, 256, 0, s>>>(tr, q, radius, delta, qperm, leaf_qstart, cand_start, cand, kept_rec); break;
      case 2: k_neighbours_scan<2><<<n_leaves, 256, 0, s>>>(tr, q, radius, delta, qperm, leaf_qstart, cand_start, cand, kept_rec); break;
      case 3: k_neighbours_scan<3><<<n_leaves, 256, 0, s>>>(tr, q, radius, delta, qperm, leaf_qstart, cand_start, cand, kept_rec); break;
      default: break;
    }
  } else {
    const int T = 128;
    const unsigned B = (unsigned)((q.Nq + T - 1) / T);
    switch (D) {
      case 1: k_neighbours_fill<1><<<B, T, 0, s>>>(tr, q, radius, delta, wkind, wparam, d_leaf_count, cand_start, cand, kept_rec); break;
      case 2: k_neighbours_fill<2><<<B, T, 0, s>>>(tr, q, radius, delta, wkind, wparam, d_leaf_count, cand_start, cand, kept_rec); break;
      case 3: k_neighbours_fill<3><<<B, T, 0, s>>>(tr, q, radius, delta, wkind, wparam, d_leaf_count, cand_start, cand, kept_rec); break;
      default: break;
    }
  }
}

void launch_eps_partitions(int D, bool fill, const TreeDev& tr, int64_t N, const double* dX, double eps, int32_t* counts,
                           const int64_t* off, int32_t* pair_leaf, int32_t* pair_pt, int32_t* leaf_count, cudaStream_t s) {
  const int T = 128;
  const unsigned B = (unsigned)((N + T - 1) / T);
  if (B == 0) return;
#define PMK_EPS(DD)                                                                                              \
  if (fill) k_eps_partitions<DD, true><<<B, T, 0, s>>>(tr, N, dX, eps, counts, off, pair_leaf, pair_pt, leaf_count); \
  else k_eps_partitions<DD, false><<<B, T, 0, s>>>(tr, N, dX, eps, counts, off, pair_leaf, pair_pt, leaf_count);
  switch (D) {
    case 1: PMK_EPS(1) break;
    case 2: PMK_EPS(2) break;
    case 3: PMK_EPS(3) break;
    default: break;
  }
#undef PMK_EPS
}

void launch_combine(int64_t Nq, const int64_t* pair_off, const double* pw, const double* pu, const double* pv,
                    double* dYq, double* dVq, int mean_only, cudaStream_t s) {
  const int T = 256;
  const unsigned B = (unsigned)((Nq + T - 1) / T);
  if (B == 0) return;
  k_combine<<<B, T, 0, s>>>(Nq, pair_off, pw, pu, pv, dYq, dVq, mean_only);
}

void launch_scan_small(const int32_t* in, int64_t* out, int n, cudaStream_t s) {
  k_scan_small<<<1, 1024, 0, s>>>(in, out, n);
}

void launch_pack_sorted_pairs(int D, const QueryPlan& q, const int32_t* sorted_pair, double* X_sorted, int32_t* leaf_sorted,
                              cudaStream_t s) {
  const int T = 256;
  const unsigned B = (unsigned)((q.n_pairs + T - 1) / T);
  if (B == 0) return;
  switch (D) {
    case 1: k_pack_sorted_pairs<1><<<B, T, 0, s>>>(q, sorted_pair, X_sorted, leaf_sorted); break;
    case 2: k_pack_sorted_pairs<2><<<B, T, 0, s>>>(q, sorted_pair, X_sorted, leaf_sorted); break;
    case 3: k_pack_sorted_pairs<3><<<B, T, 0, s>>>(q, sorted_pair, X_sorted, leaf_sorted); break;
    default: break;
  }
}

void launch_unpack_sorted_pairs(int64_t n, const int32_t* sorted_pair, const double* us, const double* vs, double* pu, double* pv,
                                cudaStream_t s) {
  const int T = 256;
  const unsigned B = (unsigned)((n + T - 1) / T);
  if (B == 0) return;
  k_unpack_sorted_pairs<<<B, T, 0, s>>>(n, sorted_pair, us, vs, pu, pv);
}

void launch_run_starts(const int32_t* keys, int64_t R, int64_t TL, int64_t leaf_base, int64_t n_own, int64_t* start, int* foreign,
                       cudaStream_t s) {
  const int T = 256;
  k_run_starts<<<(unsigned)((TL + 1 + T - 1) / T), T, 0, s>>>(keys, R, TL, leaf_base, n_own, start, foreign);
}

void launch_dense_debug(int D, const TreeDev& tr, const QueryPlan& q, double radius, double delta, uint8_t* keep, double* ts,
                        double* zs, cudaStream_t s) {
  const int T = 128;
  const int64_t total = q.Nq * (int64_t)tr.n_hp;
  const unsigned B = (unsigned)((total + T - 1) / T);
  if (B == 0) return;
  switch (D) {
    case 1: k_dense_debug<1><<<B, T, 0, s>>>(tr, q, radius, delta, keep, ts, zs); break;
    case 2: k_dense_debug<2><<<B, T, 0, s>>>(tr, q, radius, delta, keep, ts, zs); break;
    case 3: k_dense_debug<3><<<B, T, 0, s>>>(tr, q, radius, delta, keep, ts, zs); break;
    default: break;
  }
}

}  // namespace pmk
