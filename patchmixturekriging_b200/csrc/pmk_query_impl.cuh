// K3: fused per-(query, leaf) kernel -- cross-covariance k(x*, X_p), predictive mean
// dot(k, alpha), and latent variance k(x*,x*) - ||L^-1 k||^2 -- in one pass, so k(x*, X_p) never
// touches HBM.  Replaces queryinner! (reference src/RKHS/mixtureGP.jl:296-316) for all pairs
// of a leaf at once (the reference does one dtrsv per pair).
//
// One CTA = (leaf p, tile of MQ = 8*NQT pairs binned to p).  The n_pad x MQ cross-covariance tile
// lives in REGISTERS as DMMA accumulators for the whole kernel (row tiles dealt cyclically to the
// warps so the shrinking triangular work stays balanced); a right-looking blocked TRSM walks the
// 32-row blocks J:
//     W_J = C_J                                (rows of block J are final; W_J = L_JJ S_J, published via smem)
//     C_I -= M_IJ * W_J   for all I > J        (DMMA; M_IJ = L_IJ inv(L_JJ) precomputed by k_make_M, streamed
//                                               from L2/HBM as packed fragment tiles through a cp.async ring)
//     S_J = inv(L_JJ) * W_J                    (DMMA side job, only for ||s||^2 -- not on the critical path)
// One block barrier per J.  Only W_J (32 x MQ, double-buffered) ever sits in shared memory; ||s||^2 and
// dot(k, alpha) are reduced on the fly.
#pragma once
#include "pmk_internal.cuh"

namespace pmk {

static constexpr unsigned kFullQ = 0xffffffffu;

// per-phase cycle counters of k_query_pairs (thread 0 of every CTA): total, init (cross-covariance), publish+barrier,
// diagonal solve+barrier, update, #CTAs -- read through pmk_debug_counters
static __device__ unsigned long long g_query_cycles[8];
static void read_query_cycles_tu(unsigned long long* out, bool reset) {
  cudaMemcpyFromSymbol(out, g_query_cycles, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_query_cycles, z, sizeof z);
  }
}


template <int D, int NW, int NT, int NQT, int DEPTH, int GI>
__global__ void __launch_bounds__(NW * 32, 1)
k_query_pairs(LeafTable lt, PairWork w, QueryPlan q, KParams kp, int mean_only, double* __restrict__ pair_u,
              double* __restrict__ pair_v) {
  constexpr int MQ = 8 * NQT;
  constexpr int LDQ = MQ + 4;                 // == 4 or 12 (mod 16) for MQ in {8,16,24,32}: conflict-free fragment loads
  constexpr int OT = (4 * NQT + NW - 1) / NW; // diagonal-solve output tiles per warp
  __shared__ __align__(16) double Wbuf[2][32 * LDQ];   // W_J = L_JJ S_J of two consecutive blocks
  __shared__ double s_xq[D * MQ];
  __shared__ int64_t s_pair[MQ];
  __shared__ double ured[NW * MQ];
  __shared__ double vred[4 * NQT * 8];

  const int64_t tile = blockIdx.x;
  if (tile >= w.tile_off[w.n_class_leaves]) return;
  int lo = 0, hi = w.n_class_leaves;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (w.tile_off[mid] <= tile) lo = mid; else hi = mid;
  }
  const int p = w.class_leaves[lo];
  const int64_t gleaf = w.leaf_base + p;
  const int64_t pstart = w.leaf_pair_start[gleaf] + (tile - w.tile_off[lo]) * MQ;
  const int64_t pend = w.leaf_pair_start[gleaf + 1];
  const int cnt = (int)((pend - pstart) < (int64_t)MQ ? (pend - pstart) : (int64_t)MQ);

  const int n = lt.n[p], npad = lt.npad[p];
  const int nblk = npad >> 5, ntl = npad >> 3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, l = lane & 3;
  const double* __restrict__ xs = lt.xs + lt.xoff[p];
  const double* __restrict__ al = lt.alpha + lt.xoff[p];
  const int64_t xstride = lt.xstride;
  const double2* __restrict__ Lp = reinterpret_cast<const double2*>(lt.M + lt.loff[p]);   // M_IJ = L_IJ inv(L_JJ)
  const double2* __restrict__ Ip = reinterpret_cast<const double2*>(lt.Linv + lt.ioff[p]);

  PMK_CYC(long long q_total = clock64(), q_init = 0, q_pub = 0, q_diag = 0, q_upd = 0;)
  if (tid < MQ) {
    const int qi = tid < cnt ? tid : cnt - 1;
    const int64_t gp = w.sorted_pair[pstart + qi];
    s_pair[tid] = tid < cnt ? gp : (int64_t)-1;
    const int64_t j = q.pair_q[gp];
#pragma unroll
    for (int d = 0; d < D; ++d) s_xq[d * MQ + tid] = q.Xq[j * D + d];
  }
  __syncthreads();

  // ---- cross-covariance tile straight into the accumulators (acc = -k), mean partials ---------
  // (query n-tile outermost so that only two query points are live at a time)
  double acc[NT][NQT][2];
#pragma unroll
  for (int nt = 0; nt < NQT; ++nt) {
    double xq0[D], xq1[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      xq0[d] = s_xq[d * MQ + nt * 8 + 2 * l];
      xq1[d] = s_xq[d * MQ + nt * 8 + 2 * l + 1];
    }
    double up0 = 0.0, up1 = 0.0;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int t = warp + NW * i;
      const int row = 8 * t + g;
      double k0 = 0.0, k1 = 0.0;
      if ((t < ntl) && (row < n)) {
        double xr[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xr[d] = xs[d * xstride + row];
        const double a_row = al[row];
        k0 = eval_kernel<D>(kp, xq0, xr);        // evalkernel(xq, X[i])   mixtureGP.jl:304
        k1 = eval_kernel<D>(kp, xq1, xr);
        up0 = fma(k0, a_row, up0);               // dot(kq, c)             mixtureGP.jl:308
        up1 = fma(k1, a_row, up1);
      }
      acc[i][nt][0] = -k0;
      acc[i][nt][1] = -k1;
    }
    up0 += __shfl_xor_sync(kFullQ, up0, 4);
    up0 += __shfl_xor_sync(kFullQ, up0, 8);
    up0 += __shfl_xor_sync(kFullQ, up0, 16);
    up1 += __shfl_xor_sync(kFullQ, up1, 4);
    up1 += __shfl_xor_sync(kFullQ, up1, 8);
    up1 += __shfl_xor_sync(kFullQ, up1, 16);
    if (g == 0) {
      ured[warp * MQ + nt * 8 + 2 * l] = up0;
      ured[warp * MQ + nt * 8 + 2 * l + 1] = up1;
    }
  }

  if (mean_only & 1) {
    __syncthreads();
    if (tid < cnt) {
      double u = 0.0;
      for (int ww = 0; ww < NW; ++ww) u += ured[ww * MQ + tid];
      pair_u[s_pair[tid]] = u;
    }
    return;
  }

  PMK_CYC(q_init = clock64() - q_total;)
  // ---- right-looking blocked TRSM:  s = L^-1 k  (mixtureGP.jl:311), ||s||^2 on the fly ----------
  double vacc[OT][2];
#pragma unroll
  for (int k = 0; k < OT; ++k) vacc[k][0] = vacc[k][1] = 0.0;

  // Per-warp cp.async ring of L tiles.  What a warp consumes in step 3 -- for J, for ct: the tiles
  // (t_i, 4J+ct) of its active row tiles -- does not depend on the solve, so the copies run DEPTH-1
  // groups ahead of the DMMAs (across the block barriers too) without costing a register.  One commit
  // group = one (J, ct) = up to NT tiles at fixed ring positions; every lane copies and later reads
  // back exactly its own 16 bytes of each fragment-major tile.
  extern __shared__ __align__(16) unsigned char pmk_dyn_smem[];
  constexpr int NG = NT / GI;        // groups per (J, ct): GI row tiles each
  double2* ring = reinterpret_cast<double2*>(pmk_dyn_smem) + (size_t)warp * (DEPTH * GI * 32) + lane;
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
  // producer state, kept deliberately cheap: one global pointer per owned row tile (advanced by one tile =
  // 512 B per (J, ct) group), a bit mask of the row tiles that exist (t < ntl), and the row-tile threshold
  // 4*pJ+4 below which a tile is already solved.
  const double2* srcb = Lp + lane;   // + (4*pJ + pct) tiles, advanced by one tile per (J, ct)
  int toff[NT];                      // tri(t_i) * 32: start of row tile t_i (double2 units)
  unsigned exists = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int t = warp + NW * i;
    toff[i] = (int)tri(t) * 32;
    if (t < ntl) exists |= 1u << i;
  }
  int pthr = 4, pct = 0, pslot = 0;  // next group: column tile 4*pJ + pct with pthr = 4*pJ + 4
  const int pthr_end = 4 * nblk;     // pJ + 1 < nblk  <=>  pthr < 4 * nblk
  auto p_issue = [&](int ih) {       // ih: which GI-sized part of the row tiles (compile-time at every call site)
    if (pthr < pthr_end) {
      const uint32_t dst = ring_u32 + (uint32_t)(pslot * (GI * 512));
#pragma unroll
      for (int ii = 0; ii < GI; ++ii) {
        const int i = ih * GI + ii;
        if (((exists >> i) & 1u) && warp + NW * i >= pthr) cp_async16_u32(dst + ii * 512, srcb + toff[i]);
      }
      if (ih == NG - 1) {
        srcb += 32;
        if (++pct == 4) { pct = 0; pthr += 4; }
      }
    }
    cp_async_commit();               // always commit: keeps the group count in step with the consumer
    pslot = (pslot + 1 == DEPTH) ? 0 : pslot + 1;
  };
  // prologue: DEPTH-1 groups in flight.  The group sequence is (J, ct, ih) with ih fastest.
#pragma unroll
  for (int s_ = 0; s_ < DEPTH - 1; ++s_) p_issue(s_ % NG);
  int cslot = 0;                     // ring slot of the group consumed next

  for (int J = 0; J < nblk; ++J) {
    PMK_CYC(long long qc = clock64();)
    // prefetch this warp's inverse-diagonal-block tiles for step 2 (latency hidden behind the barrier)
    double2 fI[OT][4];
#pragma unroll
    for (int k = 0; k < OT; ++k) {
      const int ot = warp + NW * k;
      const int a = ot / NQT;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        fI[k][b] = (ot < 4 * NQT && b <= a) ? Ip[(size_t)J * (kInvTilesPerBlock * 32) + (a * (a + 1) / 2 + b) * 32 + lane]
                                            : make_double2(0.0, 0.0);
    }
    // 1. owners of block J's four row tiles publish W_J = -acc (their rows are final: W_J = L_JJ S_J)
    double* Wb = Wbuf[J & 1];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int t = warp + NW * i;
      if ((t >> 2) == J) {
        const int a = t & 3;
#pragma unroll
        for (int nt = 0; nt < NQT; ++nt)
          *reinterpret_cast<double2*>(&Wb[(8 * a + g) * LDQ + nt * 8 + 2 * l]) =
              make_double2(-acc[i][nt][0], -acc[i][nt][1]);
      }
    }
    __syncthreads();      // the ONLY block barrier per J (W is double-buffered)
    PMK_CYC({ long long c1 = clock64(); q_pub += c1 - qc; qc = c1; })
    // 2. side job, off the critical path: S_J = inv(L_JJ) * W_J only feeds ||s||^2 (4 x NQT tiles over the warps)
#pragma unroll
    for (int k = 0; k < OT; ++k) {
      const int ot = warp + NW * k;
      PMK_UNIFORM_IF(ot < 4 * NQT) {
        const int a = ot / NQT, nt = ot % NQT;
        double s0 = 0.0, s1 = 0.0, r0 = 0.0, r1 = 0.0;
        const double* Cb = &Wb[l * LDQ + nt * 8 + g];
        // b = 0 always; b = 1..a behind real (uniform) branches
        dmma884(s0, s1, fI[k][0].x, Cb[0]);
        dmma884(r0, r1, fI[k][0].y, Cb[4 * LDQ]);
        PMK_UNIFORM_IF(a >= 1) {
          dmma884(s0, s1, fI[k][1].x, Cb[8 * LDQ]);
          dmma884(r0, r1, fI[k][1].y, Cb[12 * LDQ]);
          PMK_UNIFORM_IF(a >= 2) {
            dmma884(s0, s1, fI[k][2].x, Cb[16 * LDQ]);
            dmma884(r0, r1, fI[k][2].y, Cb[20 * LDQ]);
            PMK_UNIFORM_IF(a >= 3) {
              dmma884(s0, s1, fI[k][3].x, Cb[24 * LDQ]);
              dmma884(r0, r1, fI[k][3].y, Cb[28 * LDQ]);
            }
          }
        }
        s0 += r0;
        s1 += r1;
        vacc[k][0] = fma(s0, s0, vacc[k][0]);
        vacc[k][1] = fma(s1, s1, vacc[k][1]);
      }
    }
    PMK_CYC({ long long c1 = clock64(); q_diag += c1 - qc; qc = c1; })
    // 3. acc[I] += M_IJ * W_J for the row tiles below block J (M_IJ W_J = L_IJ S_J).  Tile guards are REAL branches
    //    (PMK_UNIFORM_IF); within a tile the DMMAs are ordered k-step-major so that consecutive
    //    ones hit different accumulators.
    if (J + 1 < nblk) {
      unsigned active = 0;
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        if (warp + NW * i >= 4 * J + 4) active |= 1u << i;
      }
      active &= exists;
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) {
        double bf[2][NQT];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int nt = 0; nt < NQT; ++nt) bf[ks][nt] = Wb[(8 * ct + 4 * ks + l) * LDQ + nt * 8 + g];
#pragma unroll
        for (int ih = 0; ih < NG; ++ih) {
          cp_async_wait<DEPTH - 2>();        // group (J, ct, ih) has landed
          const double2* rs = ring + cslot * (GI * 32);
          cslot = (cslot + 1 == DEPTH) ? 0 : cslot + 1;
          p_issue((ih + DEPTH - 1) % NG);    // next group goes into the slot consumed one group ago
#pragma unroll
          for (int ii = 0; ii < GI; ++ii) {
            const int i = ih * GI + ii;
            PMK_UNIFORM_IF(active & (1u << i)) {
              const double2 af = rs[ii * 32];
#pragma unroll
              for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.x, bf[0][nt]);
#pragma unroll
              for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.y, bf[1][nt]);
            }
          }
        }
      }
    }
    PMK_CYC(q_upd += clock64() - qc;)
  }
  cp_async_wait<0>();

  // ---- reduce ||s||^2 and finish --------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < OT; ++k) {
    const int ot = warp + NW * k;
    if (ot < 4 * NQT) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double v = vacc[k][e];
        v += __shfl_xor_sync(kFullQ, v, 4);
        v += __shfl_xor_sync(kFullQ, v, 8);
        v += __shfl_xor_sync(kFullQ, v, 16);
        if (g == 0) vred[ot * 8 + 2 * l + e] = v;
      }
    }
  }
  __syncthreads();
  if (tid < cnt) {
    const int nt = tid >> 3, qi = tid & 7;
    double vs = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) vs += vred[(a * NQT + nt) * 8 + qi];
    double u = 0.0;
    for (int ww = 0; ww < NW; ++ww) u += ured[ww * MQ + tid];
    double xq[D];
#pragma unroll
    for (int d = 0; d < D; ++d) xq[d] = s_xq[d * MQ + tid];
    const double kxx = eval_kernel<D>(kp, xq, xq);
    double v = kxx - vs;                               // mixtureGP.jl:312, clamp(., 1e-12, Inf)
    if (!(mean_only & 2) && v < 1e-12) v = 1e-12;   // flag bit1: no clamp (evalqueryGP!, querying.jl:76-78)
    const int64_t gp = s_pair[tid];
    pair_u[gp] = u;
    pair_v[gp] = v;
  }
#ifdef PMK_PROFILE_CYCLES
  if (tid == 0) {
    atomicAdd(&g_query_cycles[0], (unsigned long long)(clock64() - q_total));
    atomicAdd(&g_query_cycles[1], (unsigned long long)q_init);
    atomicAdd(&g_query_cycles[2], (unsigned long long)q_pub);
    atomicAdd(&g_query_cycles[3], (unsigned long long)q_diag);
    atomicAdd(&g_query_cycles[4], (unsigned long long)q_upd);
    atomicAdd(&g_query_cycles[5], 1ull);
  }
#endif
}


// one translation unit per D (pmk_query_d{1,2,3}.cu) instantiates the three size classes
//   class 0: n_pad <=  512, MQ = 32     class 1: n_pad <= 1024, MQ = 16     class 2: n_pad <= 2048, MQ = 8
template <int D, int NW, int NT, int NQT, int DEPTH, int GI>
static void launch_one(unsigned grid, const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp, int mean_only,
                       double* pu, double* pv, cudaStream_t s) {
  constexpr size_t dyn = (size_t)NW * DEPTH * GI * 32 * sizeof(double2);
  static_assert(dyn <= 200 * 1024, "ring does not fit in shared memory");
  static bool configured = false;   // one attribute call per instantiation (per process; all devices are B200)
  auto kern = k_query_pairs<D, NW, NT, NQT, DEPTH, GI>;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    configured = true;
  }
  kern<<<grid, NW * 32, dyn, s>>>(lt, w, q, kp, mean_only, pu, pv);
}

template <int D>
void launch_pairs_d(int cls, unsigned grid, const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp,
                    int mean_only, double* pu, double* pv, cudaStream_t s) {
  constexpr int NW = 16;   // ring bytes = NW * DEPTH * GI * 512
  // class:   0: n_pad <= 512, 32 pairs/CTA | 1: <= 768, 24 | 2: <= 1024, 16 | 3: <= 1536, 16 | 4: <= 2048, 8
  // (the accumulators of n_pad x MQ doubles must fit the register file: NT * NQT * 4 registers per thread)
  if (cls == 0) launch_one<D, NW, 4, 4, 3, 4>(grid, lt, w, q, kp, mean_only, pu, pv, s);
  else if (cls == 1) launch_one<D, NW, 6, 3, 3, 6>(grid, lt, w, q, kp, mean_only, pu, pv, s);
  else if (cls == 2) launch_one<D, NW, 8, 2, 2, 8>(grid, lt, w, q, kp, mean_only, pu, pv, s);
  else if (cls == 3) launch_one<D, NW, 12, 2, 2, 6>(grid, lt, w, q, kp, mean_only, pu, pv, s);
  else launch_one<D, NW, 16, 1, 2, 8>(grid, lt, w, q, kp, mean_only, pu, pv, s);
}

}  // namespace pmk
