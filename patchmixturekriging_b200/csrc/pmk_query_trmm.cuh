// K3 (default solver): fused per-(query, leaf) kernel with the EXPLICIT inverse P = inv(L).
//
//   queryinner! (reference src/RKHS/mixtureGP.jl:296-316):  kq = k(x*, X_p);  u = dot(kq, c);  v = k(x*,x*) - ||L \ kq||^2
//
// Here  s = P kq  (P formed once per fit, pmk_api.cu build_operands): a triangular matrix product with NO dependency
// between row blocks, so nothing ever waits for a block to be "solved" -- the FP64 tensor pipe is fed continuously,
// and the kernel evaluations (FP64 ALU work, which B200 runs next to DMMA) are spread over the whole tile instead of
// sitting in front of it.  k(x*, X_p) still never touches HBM.
//
// Squared-exponential kernel only (the evaluation is inlined into the DMMA warps); every other kernel function takes
// the substitution kernel (pmk_query_trsm.cuh).
//
// Persistent CTAs (one per SM, 512 threads), warp-specialised with setmaxnreg:
//   * warps 0-11 "unified" (152 registers): own the tile (leaf p, MQ = 8*NQT pairs binned to p).  Output rows S (n_pad x MQ)
//     live in registers as DMMA accumulators, row tiles dealt cyclically to the warps.  The tile walks the 32-column blocks
//     J of P:   S[I] += P[I, J] * K_J   for all row tiles I >= J
//     where K_J = k(X_p[32J .. 32J+31], x*) is evaluated just in time, two blocks ahead, by the same warps (3 kernel
//     evaluations per thread and step, interleaved with the DMMA stream by the warp schedulers) into a 3-deep ring in shared
//     memory; the mean dot(kq, c) is accumulated alongside.  One 384-thread barrier per J.
//   * warps 12-14 stream the packed fragment-major P tiles into per-warp rings with 1-D TMA bulk copies completing on
//     mbarriers (one lane per (unified warp, row tile)); the rings run across tiles.
//   * warp 15 stages the tiles two ahead (pair ids, query points, leaf descriptor) and finalises finished tiles
//     (sum of the per-warp partials, k(x*,x*) - ||s||^2, clamp, scatter to the pair arrays), so the unified warps go
//     from one tile straight into the next.
#pragma once
#include "pmk_query_trsm.cuh"

namespace pmk {

static constexpr int kUW = 12;                 // unified (evaluate + DMMA) warps
static constexpr int kUnifiedThreads = kUW * 32;
static constexpr int kTrmmThreads = 512;
static constexpr int kSB = 2;                  // staged tiles: the one being worked on and the next

__device__ __forceinline__ void unified_bar() { asm volatile("bar.sync 2, %0;" ::"n"(kUnifiedThreads) : "memory"); }

__device__ __forceinline__ double q_lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ double2 q_lds_v2f64(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool q_mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// tuning switches (tools/k3_experiments.sh); measured on c3_mini, 15.84 ms with all three off:
//   PARK   suspend-time hint on the loader warps' waits            no effect
//   EARLY  non-blocking mbarrier.test_wait ahead of the evaluations  +1.2 ms (test_wait is slow)
//   ASMLDS 32-bit shared addresses through inline ld.shared          +0.2 ms (volatile asm pins the schedule)
#ifndef PMK_TRMM_PARK
#define PMK_TRMM_PARK 0
#endif
#ifndef PMK_TRMM_EARLY
#define PMK_TRMM_EARLY 0
#endif
#ifndef PMK_TRMM_ASMLDS
#define PMK_TRMM_ASMLDS 0
#endif
// blocking wait of the loader warps: lets the hardware park the thread (suspend-time hint) instead of spinning through
// the issue slots the DMMA warps need
__device__ __forceinline__ void q_mbar_wait_parked(uint32_t bar, uint32_t parity) {
  if (!PMK_TRMM_PARK) {
    q_mbar_wait(bar, parity);
    return;
  }
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(2000)
      : "memory");
}

struct TileDesc {
  int p, npad, n, valid;
};

// exp(t) for t <= 0: the cross-covariance needs ~1e-16 relative accuracy but none of libm's special cases.  Cody-Waite
// reduction, degree-13 Taylor polynomial in Estrin form (short dependency chain: FP64 latency, not throughput, limits
// a warp that has only three evaluations in flight), exponent patched in.  Max error 2 ulp vs libm on [-699, 0]
// (tests/test_host.py checks the same arithmetic in numpy); arguments below -699 return exp(-699) ~ 1e-304.
__device__ __forceinline__ double exp_neg(double t) {
  const double L2E = 1.4426950408889634, LN2H = 6.93147180369123816490e-01, LN2L = 1.90821492927058770002e-10;
  t = fmax(t, -699.0);
  const double nf = rint(t * L2E);
  double r = fma(nf, -LN2H, t);
  r = fma(nf, -LN2L, r);
  const double c2 = 0.5, c3 = 1.0 / 6, c4 = 1.0 / 24, c5 = 1.0 / 120, c6 = 1.0 / 720, c7 = 1.0 / 5040, c8 = 1.0 / 40320,
               c9 = 1.0 / 362880, c10 = 1.0 / 3628800, c11 = 1.0 / 39916800, c12 = 1.0 / 479001600, c13 = 1.0 / 6227020800.0;
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double p01 = 1.0 + r, p23 = fma(c3, r, c2), p45 = fma(c5, r, c4), p67 = fma(c7, r, c6), p89 = fma(c9, r, c8),
               pab = fma(c11, r, c10), pcd = fma(c13, r, c12);
  const double q0 = fma(p23, r2, p01), q1 = fma(p67, r2, p45), q2 = fma(pab, r2, p89);
  const double sres = fma(fma(pcd, r4, q2), r8, fma(q1, r4, q0));
  const int n = (int)nf;
  return __hiloint2double(__double2hiint(sres) + (n << 20), __double2loint(sres));
}

template <int D, int NT, int NQT, int CG, int GI, int DEPTH, int NPMAX>
__global__ void __launch_bounds__(kTrmmThreads, 1)
k_query_trmm(LeafTable lt, PairWork w, QueryPlan q, KParams kp, int flags, double* __restrict__ pair_u,
             double* __restrict__ pair_v) {
  constexpr int MQ = 8 * NQT;
  constexpr int LDQ = MQ + 4;                 // == 4 or 12 (mod 16): conflict-free fragment loads
  constexpr int NG = 4 / CG;
  constexpr int NSUB = NT / GI;
  constexpr int SLOT_BYTES = GI * CG * 512;
  constexpr int RPL = MQ > 16 ? 1 : (MQ > 8 ? 2 : 4);   // rows evaluated at once by one warp (lanes = RPL x MQ)
  constexpr bool STAGE_X = NPMAX > 0;         // the leaf's training inputs and alpha staged in shared memory by TMA
  constexpr int RING_BYTES = kUW * DEPTH * SLOT_BYTES;
  static_assert(CG == 1 || CG == 2 || CG == 4, "CG divides the 4 column tiles of a block");
  static_assert(NT % GI == 0 && (GI == NT || CG == 1), "row-tile subsets only with single-column groups");
  constexpr int NKB = 3;                      // K_J ring: block J in use, J+1 ready, J+2 being written
  __shared__ double s_xq[kSB][D * MQ];
  __shared__ int64_t s_pair[kSB][MQ];
  __shared__ TileDesc s_desc[kSB];
  __shared__ double vred[2][kUW * MQ];
  __shared__ double ured[2][kUW * 32];
  __shared__ __align__(8) uint64_t full_bar[kUW * DEPTH], empty_bar[kUW * DEPTH];
  __shared__ __align__(8) uint64_t stage_full[kSB], stage_empty[kSB], tile_done[2], fin_done[2];

  extern __shared__ __align__(128) unsigned char pmk_dyn_smem[];   // P rings: [unified warp][DEPTH][GI][CG][512 B],
                                                                   // then s_X[kSB][D+1][NPMAX] (inputs SoA, alpha)
  double* s_X = reinterpret_cast<double*>(pmk_dyn_smem + RING_BYTES);
  // K_J ring: rows = training points of block J, cols = queries
  double* Kb = reinterpret_cast<double*>(pmk_dyn_smem + RING_BYTES + (size_t)kSB * (D + 1) * NPMAX * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = w.tile_off[w.n_class_leaves];
  if (tid == 0) {
    for (int k = 0; k < kUW * DEPTH; ++k) {
      q_mbar_init(q_smem_u32(&full_bar[k]), 1);
      q_mbar_init(q_smem_u32(&empty_bar[k]), 1);
    }
    for (int k = 0; k < kSB; ++k) {
      q_mbar_init(q_smem_u32(&stage_full[k]), 32);           // every lane of warp 15
      q_mbar_init(q_smem_u32(&stage_empty[k]), kUW + 3);      // lane 0 of every reader warp (warp 15 itself reads last)
    }
    for (int k = 0; k < 2; ++k) {
      q_mbar_init(q_smem_u32(&tile_done[k]), kUW);
      q_mbar_init(q_smem_u32(&fin_done[k]), 32);
    }

    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t ring0 = q_smem_u32(pmk_dyn_smem);

  if (warp >= kUW) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");   // the whole loader warpgroup, one instruction
  if (warp == kUW + 3) {
    // ============================ warp 15: stage tiles ahead, finalise finished tiles =====================
    int lo = 0;
    const int64_t my_tiles = n_tiles > (int64_t)blockIdx.x ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    auto stage = [&](int64_t kt) {      // tile kt of this CTA into buffer kt % kSB (the buffer is free: see the loop below)
      const int64_t tile = blockIdx.x + kt * (int64_t)gridDim.x;
      const int sb = (int)(kt % kSB);
      while (lo + 1 < w.n_class_leaves && w.tile_off[lo + 1] <= tile) ++lo;     // tiles come in increasing order
      const int p = w.class_leaves[lo];
      const int64_t gleaf = w.leaf_base + p;
      const int64_t pstart = w.leaf_pair_start[gleaf] + (tile - w.tile_off[lo]) * MQ;
      const int64_t pend = w.leaf_pair_start[gleaf + 1];
      const int cnt = (int)((pend - pstart) < (int64_t)MQ ? (pend - pstart) : (int64_t)MQ);
      const int npad = lt.npad[p];
      if (lane < MQ) {
        const int qi = lane < cnt ? lane : cnt - 1;
        const int64_t gp = w.sorted_pair[pstart + qi];
        s_pair[sb][lane] = lane < cnt ? gp : (int64_t)-1;
        const int64_t j = q.pair_q[gp];
#pragma unroll
        for (int d = 0; d < D; ++d) s_xq[sb][d * MQ + lane] = q.Xq[j * D + d];
      }
      const uint32_t fb = q_smem_u32(&stage_full[sb]);
      if (lane == 0) {
        TileDesc dsc;
        dsc.p = p;
        dsc.npad = npad;
        dsc.n = lt.n[p];
        dsc.valid = 1;
        s_desc[sb] = dsc;
        if (STAGE_X) {
          const uint32_t bytes = (uint32_t)npad * 8u;
          q_mbar_expect_tx(fb, (D + 1) * bytes);
          const uint32_t dst = q_smem_u32(s_X + (size_t)sb * (D + 1) * (NPMAX > 0 ? NPMAX : 1));
#pragma unroll
          for (int d = 0; d < D; ++d)
            q_bulk_g2s(dst + d * (NPMAX * 8), lt.xs + d * lt.xstride + lt.xoff[p], bytes, fb);
          q_bulk_g2s(dst + D * (NPMAX * 8), lt.alpha + lt.xoff[p], bytes, fb);
        } else {
          q_mbar_arrive(fb);
        }
      } else {
        q_mbar_arrive(fb);
      }
    };
    for (int64_t kt = 0; kt < kSB && kt < my_tiles; ++kt) stage(kt);
    for (int64_t kf = 0; kf < my_tiles; ++kf) {
      // ---- finalise tile kf, then reuse its staging buffer for tile kf + kSB
      const int sb = (int)(kf % kSB), vb = (int)(kf & 1);
      q_mbar_wait_parked(q_smem_u32(&tile_done[vb]), (uint32_t)((kf >> 1) & 1));
      if (lane < MQ) {
        double vs = 0.0, u = 0.0;
#pragma unroll
        for (int ww = 0; ww < kUW; ++ww) vs += vred[vb][ww * MQ + lane];
#pragma unroll
        for (int ww = 0; ww < kUW; ++ww)
#pragma unroll
          for (int ro = 0; ro < RPL; ++ro) u += ured[vb][ww * 32 + ro * MQ + lane];
        double xq[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xq[d] = s_xq[sb][d * MQ + lane];
        const double kxx = eval_kernel<D>(kp, xq, xq);
        double v = kxx - vs;                               // mixtureGP.jl:312, clamp(., 1e-12, Inf)
        if (!(flags & 2) && v < 1e-12) v = 1e-12;        // flag bit1: no clamp (evalqueryGP!, querying.jl:76-78)
        const int64_t gp = s_pair[sb][lane];
        if (gp >= 0) {
          pair_u[gp] = u;
          pair_v[gp] = v;
        }
      }
      q_mbar_arrive(q_smem_u32(&fin_done[vb]));
      __syncwarp();
      if (kf + kSB < my_tiles) {
        q_mbar_wait_parked(q_smem_u32(&stage_empty[sb]), (uint32_t)((kf / kSB) & 1));   // producers and unified warps are done with it
        stage(kf + kSB);
      }
    }
    return;
  }

  if (warp >= kUW) {
    // ============================ warps 12-14: P tile stream ================================================
    constexpr int LPC = NT <= 8 ? 8 : (NT <= 16 ? 16 : 32);   // lanes per unified warp (one lane per row tile)
    constexpr int CPR = 32 / LPC;
    constexpr int ROUNDS = 4 / CPR;
    const int pw = warp - kUW;
    const int i = lane % LPC, sub = lane / LPC;
    uint32_t slot[ROUNDS], ph[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) slot[r] = ph[r] = 0;
    int64_t kt = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++kt) {
      const int sb = (int)(kt % kSB);
      q_mbar_wait(q_smem_u32(&stage_full[sb]), (uint32_t)((kt / kSB) & 1));
      const int p = s_desc[sb].p;
      const int n_pts = s_desc[sb].n;
      __syncwarp();
      if (lane == 0) q_mbar_arrive(q_smem_u32(&stage_empty[sb]));
      // only the row / column tiles that hold real points take part: the identity padding of the fit (n_pad is a multiple
      // of 32) contributes exact zeros, so the pair kernel works on ceil(n/8) tiles
      const int ntl = (n_pts + 7) >> 3, nblk = (ntl + 3) >> 2;
      const char* Pp = reinterpret_cast<const char*>(lt.P + lt.loff[p]);
      for (int J = 0; J < nblk; ++J) {
#pragma unroll 1
        for (int cg = 0; cg < NG; ++cg) {
          const int thr = 4 * J + CG * cg;        // first column tile of the group: row tiles t >= thr take part
          bool act[ROUNDS];
          unsigned msk[ROUNDS];
#pragma unroll
          for (int r = 0; r < ROUNDS; ++r) {
            const int t = 4 * pw + r * CPR + sub + kUW * i;
            act[r] = (i < NT) && (t < ntl) && (t >= thr);
            const unsigned bal = __ballot_sync(kFullQ, act[r]);
            msk[r] = LPC == 32 ? bal : ((bal >> (sub * LPC)) & ((1u << (LPC & 31)) - 1u));
          }
#pragma unroll
          for (int ih = 0; ih < NSUB; ++ih) {
            constexpr unsigned SUBMASK = (1u << GI) - 1u;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
              const int cw = 4 * pw + r * CPR + sub;
              const unsigned sm = msk[r] & (SUBMASK << (ih * GI));
              const bool mine = (i / GI) == ih;
              if (sm != 0 && mine && (i == ih * GI || act[r])) {
                const int t = cw + kUW * i;
                // column tiles of the group that exist for row tile t (the diagonal block is triangular): c <= t
                uint32_t total = 0;
#pragma unroll
                for (int ii = 0; ii < GI; ++ii) {
                  if (sm & (1u << (ih * GI + ii))) {
                    const int tt = cw + kUW * (ih * GI + ii);
                    const int nc = tt - thr + 1 < CG ? tt - thr + 1 : CG;
                    total += (uint32_t)nc * 512u;
                  }
                }
                const uint32_t fb = q_smem_u32(&full_bar[cw * DEPTH + slot[r]]);
                q_mbar_wait_parked(q_smem_u32(&empty_bar[cw * DEPTH + slot[r]]), ph[r] ^ 1u);
                if (i == ih * GI) q_mbar_expect_tx(fb, (PMK_K3_X & 4) ? 0u : total);
                if (act[r] && !(PMK_K3_X & 4)) {
                  const int nc = t - thr + 1 < CG ? t - thr + 1 : CG;
                  q_bulk_g2s(ring0 + (uint32_t)((cw * DEPTH + slot[r]) * SLOT_BYTES + (i - ih * GI) * (CG * 512)),
                             Pp + (tri(t) + (size_t)thr) * 512, (uint32_t)nc * 512u, fb);
                }
              }
              if (sm != 0) {
                if (++slot[r] == DEPTH) { slot[r] = 0; ph[r] ^= 1u; }
              }
            }
          }
        }
      }
    }
    return;
  }

  // ================================== unified warps =============================================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 152;\n");
  const int g = lane >> 2, l = lane & 3;
  // shared-memory operands are addressed with 32-bit shared addresses + immediates (the generic-pointer form made the
  // compiler rebuild the shared window base for every load once registers got tight)
  const uint32_t ring_u32 = ring0 + (uint32_t)(warp * (DEPTH * SLOT_BYTES) + lane * 16);
  const uint32_t kb_u32 = q_smem_u32(Kb) + (uint32_t)((l * LDQ + g) * 8);     // this lane's B fragment inside a K block
  const uint32_t my_full = q_smem_u32(&full_bar[warp * DEPTH]);
  const uint32_t my_empty = q_smem_u32(&empty_bar[warp * DEPTH]);
  uint32_t slot = 0, ph = 0;
  const int64_t xstride = lt.xstride;
  const int ej = lane % MQ, ero = lane / MQ;            // evaluation role: query ej, row offset ero (valid if ero < RPL)
  const bool e_on = lane < RPL * MQ;

  int64_t kt = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++kt) {
    const int sb = (int)(kt % kSB), vb = (int)(kt & 1);
    PMK_CYC(long long c_t0 = clock64(), c_eval = 0, c_bar = 0, c_full = 0, c_stage = 0, c_end = 0;)
    q_mbar_wait(q_smem_u32(&stage_full[sb]), (uint32_t)((kt / kSB) & 1));
    PMK_CYC(c_stage = clock64() - c_t0;)
    const int p = s_desc[sb].p, n = s_desc[sb].n;
    const int ntl = (n + 7) >> 3, nblk = (ntl + 3) >> 2;      // tiles with real points only (see the producer)
    const double* __restrict__ xs = lt.xs + lt.xoff[p];
    const double* __restrict__ al = lt.alpha + lt.xoff[p];
    const double* sX = s_X + (size_t)sb * (D + 1) * (NPMAX > 0 ? NPMAX : 1);
    double xq[D];
#pragma unroll
    for (int d = 0; d < D; ++d) xq[d] = s_xq[sb][d * MQ + ej];
    double usum = 0.0;

    // K_Jb = k(X[32 Jb + r], x*_j) for this warp's rows r of block Jb, into the ring; mean partials on the way
    auto eval_block = [&](int Jb) {
      const uint32_t kb = (uint32_t)Jb % NKB;
      double* Kd = Kb + kb * (32 * LDQ);
#pragma unroll
      for (int kk = 0; kk < (32 + kUW * RPL - 1) / (kUW * RPL); ++kk) {
        const int r = warp * RPL + ero + kUW * RPL * kk;
        if (e_on && r < 32) {
          const int row = 32 * Jb + r;
          double kv = 0.0;
          if (row < n) {
            double xr[D];
#pragma unroll
            for (int d = 0; d < D; ++d) xr[d] = STAGE_X ? sX[d * NPMAX + row] : xs[d * xstride + row];
            // evalkernel(xq, X[i]), mixtureGP.jl:304, squared exponential only (kernel.jl:350-357), fully inline: an
            // out-of-line call here would spill the accumulators.  exp(-a |x - z|^2) without the reference's
            // sqrt / re-square round trip: <= 2 ulp from it, far inside the 1e-9 contract of the posterior.
            double s2 = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
              const double dd = xq[d] - xr[d];
              s2 = fma(dd, dd, s2);
            }
            kv = (PMK_K3_X & 1) ? s2 : exp_neg(-kp.p * s2);
            usum = fma(kv, STAGE_X ? sX[D * NPMAX + row] : al[row], usum);   // dot(kq, c)    mixtureGP.jl:308
          }
          Kd[r * LDQ + ej] = kv;
        }
      }
    };

    PMK_CYC(long long c_a = clock64();)
    unified_bar();                    // everyone has left the previous tile: the K ring may be overwritten
    eval_block(0);
    if (nblk > 1) eval_block(1);
    PMK_CYC(c_eval += clock64() - c_a;)

    double acc[NT][NQT][2];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int nt = 0; nt < NQT; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;
    unsigned exists = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i)
      if (warp + kUW * i < ntl) exists |= 1u << i;

    for (int J = 0; J < nblk; ++J) {
      PMK_CYC(c_a = clock64();)
      // one 384-thread barrier per step: K_J (written during step J-2) and K_{J+1} are complete and step J-1 is over
      // everywhere, so K_{J+2} may overwrite K_{J-1}.  (An mbarrier hand-over without any block barrier measured 5 % slower.)
      if (!(PMK_K3_X & 8)) unified_bar();
      // is this step's first operand group already here?  (asked now, needed after the evaluations)
      const bool early = PMK_TRMM_EARLY ? q_mbar_test(my_full + slot * 8, ph) : false;
      PMK_CYC({ long long c_b = clock64(); c_bar += c_b - c_a; c_a = c_b; })
      if (J + 2 < nblk) eval_block(J + 2);
      PMK_CYC(c_eval += clock64() - c_a;)
      const uint32_t kj = kb_u32 + ((uint32_t)J % NKB) * (32 * LDQ * 8);
      const double* Kjg = Kb + (J % NKB) * (32 * LDQ);
      bool landed = early;
#pragma unroll 1
      for (int cg = 0; cg < NG; ++cg) {
        unsigned gact = 0;            // row tiles that take part in this column group: t >= 4J + CG cg
#pragma unroll
        for (int i = 0; i < NT; ++i)
          if (warp + kUW * i >= 4 * J + CG * cg) gact |= 1u << i;
        gact &= exists;
#pragma unroll
        for (int ih = 0; ih < NSUB; ++ih) {
          constexpr unsigned SUBMASK = (1u << GI) - 1u;
          PMK_UNIFORM_IF((gact & (SUBMASK << (ih * GI))) != 0) {
            PMK_CYC(c_a = clock64();)
            if (!landed) q_mbar_wait(my_full + slot * 8, ph);   // this warp's operand group (J, cg, ih) has landed
            landed = false;
            PMK_CYC(c_full += clock64() - c_a;)
            const uint32_t rs = ring_u32 + slot * SLOT_BYTES;
            const double2* rsg = reinterpret_cast<const double2*>(pmk_dyn_smem) + (size_t)warp * (DEPTH * SLOT_BYTES / 16) + lane +
                                 slot * (SLOT_BYTES / 16);
#pragma unroll
            for (int c = 0; c < CG; ++c) {
              const int ct = cg * CG + c;
              double bf[2][NQT];
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int nt = 0; nt < NQT; ++nt) bf[ks][nt] = PMK_TRMM_ASMLDS ? q_lds_f64(kj + (uint32_t)(((8 * ct + 4 * ks) * LDQ + nt * 8) * 8))
                                                                              : Kjg[(8 * ct + 4 * ks + l) * LDQ + nt * 8 + g];
#pragma unroll
              for (int ii = 0; ii < GI; ++ii) {
                const int i = ih * GI + ii;
                // P is lower triangular: row tile t meets column tile 4J+ct only if t >= 4J+ct
                PMK_UNIFORM_IF((gact & (1u << i)) && (c == 0 || warp + kUW * i >= 4 * J + ct)) {
                  const double2 af = PMK_TRMM_ASMLDS ? q_lds_v2f64(rs + (uint32_t)((ii * CG + c) * 512)) : rsg[(ii * CG + c) * 32];
#pragma unroll
                  for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.x, bf[0][nt]);
#pragma unroll
                  for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.y, bf[1][nt]);
                }
              }
            }
            __syncwarp();
            if (lane == 0) q_mbar_arrive(my_empty + slot * 8);   // slot free: the producer may refill it
            if (++slot == DEPTH) { slot = 0; ph ^= 1u; }
          }
        }
      }
    }

    // ---- ||s||^2 partials of this warp's rows, mean partials; hand the tile to warp 15 ---------------------------
    PMK_CYC(c_a = clock64();)
    if (kt >= 2) q_mbar_wait(q_smem_u32(&fin_done[vb]), (uint32_t)(((kt >> 1) - 1) & 1));
#pragma unroll
    for (int nt = 0; nt < NQT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < NT; ++i) v = fma(acc[i][nt][e], acc[i][nt][e], v);
        v += __shfl_xor_sync(kFullQ, v, 4);
        v += __shfl_xor_sync(kFullQ, v, 8);
        v += __shfl_xor_sync(kFullQ, v, 16);
        if (g == 0) vred[vb][warp * MQ + nt * 8 + 2 * l + e] = v;
      }
    }
    ured[vb][warp * 32 + lane] = e_on ? usum : 0.0;
    __syncwarp();
    if (lane == 0) {
      q_mbar_arrive(q_smem_u32(&tile_done[vb]));
      q_mbar_arrive(q_smem_u32(&stage_empty[sb]));
    }
#ifdef PMK_PROFILE_CYCLES
    if (lane == 0 && (warp == 0 || warp == 11)) {      // counters 0-5: warp 0 (most row tiles); 6-7: warp 11's waits
      const long long c_now = clock64();
      if (warp == 0) {
        atomicAdd(&g_query_cycles[0], (unsigned long long)(c_now - c_t0));
        atomicAdd(&g_query_cycles[1], (unsigned long long)c_eval);
        atomicAdd(&g_query_cycles[2], (unsigned long long)c_bar);
        atomicAdd(&g_query_cycles[3], (unsigned long long)c_full);
        atomicAdd(&g_query_cycles[4], (unsigned long long)(c_stage + (c_now - c_a)));
        atomicAdd(&g_query_cycles[5], 1ull);
      } else {
        atomicAdd(&g_query_cycles[6], (unsigned long long)c_bar);
        atomicAdd(&g_query_cycles[7], (unsigned long long)c_full);
      }
    }
#endif
  }
}

template <int D, int NT, int NQT, int CG, int GI, int DEPTH, int NPMAX>
static void launch_trmm_one(const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp, int flags,
                            double* pu, double* pv, cudaStream_t s) {
  constexpr size_t dyn = (size_t)kUW * DEPTH * GI * CG * 512 + (size_t)kSB * (D + 1) * NPMAX * 8 + (size_t)3 * 32 * (8 * NQT + 4) * 8;
  static_assert(dyn <= 214 * 1024, "rings, staged inputs and the K ring do not fit in shared memory");
  auto kern = k_query_trmm<D, NT, NQT, CG, GI, DEPTH, NPMAX>;
  const int n_sm = device_sm_count();
  static DeviceOnce once;
  once.run([&] { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn); });
  kern<<<n_sm, kTrmmThreads, dyn, s>>>(lt, w, q, kp, flags, pu, pv);   // persistent: tiles are strided over the CTAs
}

#ifndef PMK_TRMM_CG0
#define PMK_TRMM_CG0 2
#endif
#ifndef PMK_TRMM_DEPTH0
#define PMK_TRMM_DEPTH0 2
#endif
#ifndef PMK_TRMM_CG1
#define PMK_TRMM_CG1 1
#endif
#ifndef PMK_TRMM_DEPTH1
#define PMK_TRMM_DEPTH1 3
#endif

template <int D>
void launch_trmm_d(int cls, const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp, int flags, double* pu,
                   double* pv, cudaStream_t s) {
  // same size classes and tile shapes as the substitution kernel (launch_pairs_d)
  // last argument: capacity of the staged-input buffers (0 = read the leaf's inputs from global memory)
  if (cls == 0) launch_trmm_one<D, 6, 4, PMK_TRMM_CG0, 6, PMK_TRMM_DEPTH0, 512>(lt, w, q, kp, flags, pu, pv, s);
  else if (cls == 1) launch_trmm_one<D, 8, 3, PMK_TRMM_CG1, 8, PMK_TRMM_DEPTH1, (D <= 2 ? 768 : 0)>(lt, w, q, kp, flags, pu, pv, s);
  else if (cls == 2) launch_trmm_one<D, 11, 2, 1, 11, 2, 0>(lt, w, q, kp, flags, pu, pv, s);
  else if (cls == 3) launch_trmm_one<D, 16, 1, 1, 8, 3, 0>(lt, w, q, kp, flags, pu, pv, s);
  else launch_trmm_one<D, 22, 1, 1, 11, 3, 0>(lt, w, q, kp, flags, pu, pv, s);
}

}  // namespace pmk
