// K3: fused per-(query, leaf) kernel -- cross-covariance k(x*, X_p), predictive mean
// dot(k, alpha), and latent variance k(x*,x*) - ||L^-1 k||^2 -- in one pass, so k(x*, X_p) never
// touches HBM.  Replaces queryinner! (reference src/RKHS/mixtureGP.jl:296-316) for all pairs
// of a leaf at once (the reference does one dtrsv per pair).
//
// Persistent, warp-specialised CTAs (one per SM, 512 threads):
//   * 12 CONSUMER warps (setmaxnreg 160) own the tile of work (leaf p, MQ = 8*NQT pairs binned to p).  The
//     n_pad x MQ cross-covariance tile lives in REGISTERS as DMMA accumulators for the whole tile (row tiles dealt
//     cyclically to the warps so the shrinking triangular work stays balanced); a right-looking blocked TRSM
//     walks the 32-row blocks J:
//         W_J = C_J                             (rows of block J are final; W_J = L_JJ S_J, published via smem)
//         C_I -= M_IJ * W_J   for all I > J     (DMMA; M_IJ = L_IJ inv(L_JJ) precomputed by k_make_M)
//         S_J = inv(L_JJ) * W_J                 (DMMA side job, only for ||s||^2 -- not on the critical path)
//     One block barrier per J.  Only W_J (32 x MQ, double-buffered) and inv(L_JJ) sit in shared memory; ||s||^2 and
//     dot(k, alpha) are reduced on the fly.
//   * 4 PRODUCER warps (setmaxnreg 32) stream the packed fragment-major M tiles from L2/HBM into per-consumer-warp
//     rings with 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) completing on mbarriers: one lane per (consumer warp,
//     row tile), one copy per (J, CG column tiles) = CG*512 contiguous bytes.  The consumers' instruction stream holds
//     no address arithmetic for the operand stream at all: wait(full) - LDS.128 - 8 DMMA ... - arrive(empty).
//     The ring runs across work tiles, so the next tile's first operands are in flight during the epilogue.
#pragma once
#include "pmk_internal.cuh"

namespace pmk {

static constexpr unsigned kFullQ = 0xffffffffu;
static constexpr int kCW = 12;                 // consumer warps (3 per SM sub-partition)
static constexpr int kPW = 4;                  // producer warps (one warpgroup: setmaxnreg is per warpgroup)
static constexpr int kConsumerThreads = kCW * 32;
static constexpr int kK3Threads = (kCW + kPW) * 32;

// timing experiments only (tools/k3_experiments.sh; results are WRONG with any bit set): 1 = no kernel evaluations,
// 2 = no ||s||^2 side job, 4 = no M-tile copies, 8 = no block barrier
#ifndef PMK_K3_X
#define PMK_K3_X 0
#endif

// per-phase cycle counters of k_query_pairs (thread 0 of every CTA): total, init (cross-covariance), publish+barrier,
// diagonal solve+barrier, update, #tiles -- read through pmk_debug_counters
static __device__ unsigned long long g_query_cycles[8];
static void read_query_cycles_tu(unsigned long long* out, bool reset) {
  cudaMemcpyFromSymbol(out, g_query_cycles, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_query_cycles, z, sizeof z);
  }
}

__device__ __forceinline__ uint32_t q_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void q_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void q_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void q_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void q_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void q_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }

// tile -> (class leaf slot) by binary search in the exclusive scan of tiles per class leaf
__device__ __forceinline__ int find_tile_leaf(const PairWork& w, int64_t tile) {
  int lo = 0, hi = w.n_class_leaves;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (w.tile_off[mid] <= tile) lo = mid; else hi = mid;
  }
  return lo;
}

template <int D, int NT, int NQT, int CG, int GI, int DEPTH>
__global__ void __launch_bounds__(kK3Threads, 1)
k_query_pairs(LeafTable lt, PairWork w, QueryPlan q, KParams kp, int mean_only, double* __restrict__ pair_u,
              double* __restrict__ pair_v) {
  constexpr int MQ = 8 * NQT;
  constexpr int LDQ = MQ + 4;                 // == 4 or 12 (mod 16) for MQ in {8,16,24,32}: conflict-free fragment loads
  constexpr int OT = (4 * NQT + kCW - 1) / kCW; // diagonal-solve output tiles per warp
  constexpr int NG = 4 / CG;                  // column groups per block J
  constexpr int NSUB = NT / GI;               // row-tile subsets per column group (big leaves: a slot holds GI < NT tiles)
  constexpr int SLOT_BYTES = GI * CG * 512;   // one operand group of one consumer warp: GI row tiles x CG column tiles
  static_assert(CG == 1 || CG == 2 || CG == 4, "CG divides the 4 column tiles of a block");
  static_assert(NT % GI == 0 && (GI == NT || CG == 1), "row-tile subsets only with single-column groups");
  __shared__ __align__(16) double Wbuf[2][32 * LDQ];          // W_J = L_JJ S_J of two consecutive blocks
  __shared__ __align__(16) double Ibuf[2][kInvDoublesPerBlock]; // inv(L_JJ) of two consecutive blocks
  __shared__ double s_xq[D * MQ];
  __shared__ int64_t s_pair[MQ];
  __shared__ double ured[kCW * MQ];
  __shared__ double vred[4 * NQT * 8];
  __shared__ __align__(8) uint64_t full_bar[kCW * DEPTH];
  __shared__ __align__(8) uint64_t empty_bar[kCW * DEPTH];
  extern __shared__ __align__(128) unsigned char pmk_dyn_smem[];   // rings: [consumer warp][DEPTH][NT][CG][512 B]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = w.tile_off[w.n_class_leaves];
  if (tid == 0) {
    for (int k = 0; k < kCW * DEPTH; ++k) {
      q_mbar_init(q_smem_u32(&full_bar[k]), 1);
      q_mbar_init(q_smem_u32(&empty_bar[k]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t ring0 = q_smem_u32(pmk_dyn_smem);

  if (warp >= kCW) {
    // ================================ producers ================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");
    if (mean_only & 1) return;
    constexpr int LPC = NT <= 8 ? 8 : (NT <= 16 ? 16 : 32);   // lanes per consumer warp (one lane per row tile)
    constexpr int CPR = 32 / LPC;             // consumer warps served per round
    constexpr int ROUNDS = 4 / CPR;
    const int pw = warp - kCW;
    if (4 * pw >= kCW) return;                // producer warp pw serves consumer warps 4pw .. 4pw+3
    const int i = lane % LPC, sub = lane / LPC;
    uint32_t slot[ROUNDS], ph[ROUNDS];        // ring position of this lane's consumer warp (both sides count groups alike)
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) slot[r] = ph[r] = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int lo_p = find_tile_leaf(w, tile);
      const int p = w.class_leaves[lo_p];
      const int npad = lt.npad[p];
      const int nblk = npad >> 5, ntl = npad >> 3;
      const char* Mp = reinterpret_cast<const char*>(lt.M + lt.loff[p]);
      const int J0 = (mean_only & 4) ? (int)((tile - w.tile_off[lo_p]) * MQ) >> 5 : 0;   // inversion: first non-zero block
      for (int J = J0; J + 1 < nblk; ++J) {
        const int thr = 4 * J + 4;
        bool act[ROUNDS];
        unsigned msk[ROUNDS];             // active row tiles of this lane's consumer warp (bit i)
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
          const int t = 4 * pw + r * CPR + sub + kCW * i;
          act[r] = (i < NT) && (t < ntl) && (t >= thr);
          const unsigned bal = __ballot_sync(kFullQ, act[r]);
          msk[r] = LPC == 32 ? bal : ((bal >> (sub * LPC)) & ((1u << (LPC & 31)) - 1u));
        }
#pragma unroll 1
        for (int cg = 0; cg < NG; ++cg) {
#pragma unroll
          for (int ih = 0; ih < NSUB; ++ih) {
            constexpr unsigned SUBMASK = (1u << GI) - 1u;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
              const int cw = 4 * pw + r * CPR + sub;
              const int kk = __popc(msk[r] & (SUBMASK << (ih * GI)));
              const bool mine = (i / GI) == ih;
              if (kk > 0 && mine && (i == ih * GI || act[r])) {
                const uint32_t fb = q_smem_u32(&full_bar[cw * DEPTH + slot[r]]);
                q_mbar_wait(q_smem_u32(&empty_bar[cw * DEPTH + slot[r]]), ph[r] ^ 1u);
                if (i == ih * GI) q_mbar_expect_tx(fb, (PMK_K3_X & 4) ? 0u : (uint32_t)(kk * CG * 512));
                if (act[r] && !(PMK_K3_X & 4)) {
                  const int t = cw + kCW * i;
                  q_bulk_g2s(ring0 + (uint32_t)((cw * DEPTH + slot[r]) * SLOT_BYTES + (i - ih * GI) * (CG * 512)),
                             Mp + (tri(t) + (size_t)(4 * J + CG * cg)) * 512, CG * 512, fb);
                }
              }
              // every lane tracks the group counter of ITS consumer warp (a warp whose subset is exhausted skips the group)
              if (kk > 0) {
                if (++slot[r] == DEPTH) { slot[r] = 0; ph[r] ^= 1u; }
              }
            }
          }
        }
      }
    }
    return;
  }

  // ================================== consumers ==================================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 160;\n");
  const int g = lane >> 2, l = lane & 3;
  const double2* ring = reinterpret_cast<const double2*>(pmk_dyn_smem) + (size_t)warp * (DEPTH * SLOT_BYTES / 16) + lane;
  const uint32_t my_full = q_smem_u32(&full_bar[warp * DEPTH]);
  const uint32_t my_empty = q_smem_u32(&empty_bar[warp * DEPTH]);
  uint32_t slot = 0, ph = 0;
  const int64_t xstride = lt.xstride;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int lo = find_tile_leaf(w, tile);
    const int p = w.class_leaves[lo];
    const int64_t gleaf = w.leaf_base + p;
    const bool invert = (mean_only & 4) != 0;      // RHS = MQ identity columns: the tile computes columns of inv(L)
    const int col0 = invert ? (int)(tile - w.tile_off[lo]) * MQ : 0;
    const int64_t pstart = invert ? 0 : w.leaf_pair_start[gleaf] + (tile - w.tile_off[lo]) * MQ;
    const int64_t pend = invert ? 0 : w.leaf_pair_start[gleaf + 1];
    const int cnt = (int)((pend - pstart) < (int64_t)MQ ? (pend - pstart) : (int64_t)MQ);
    const int n = lt.n[p], npad = lt.npad[p];
    const int nblk = npad >> 5, ntl = npad >> 3;
    const double* __restrict__ xs = lt.xs + lt.xoff[p];
    const double* __restrict__ al = lt.alpha + lt.xoff[p];
    const double* __restrict__ Ip = lt.Linv + lt.ioff[p];

    PMK_CYC(long long q_total = clock64(), q_init = 0, q_pub = 0, q_diag = 0, q_upd = 0;)
    if (!invert && tid < MQ) {
      const int qi = tid < cnt ? tid : cnt - 1;
      const int64_t gp = w.sorted_pair[pstart + qi];
      s_pair[tid] = tid < cnt ? gp : (int64_t)-1;
      const int64_t j = q.pair_q[gp];
#pragma unroll
      for (int d = 0; d < D; ++d) s_xq[d * MQ + tid] = q.Xq[j * D + d];
    }
    // inv(L_00) for the first side job (cp.async: every thread waits for its own 16 bytes before the block barrier)
    if (!(mean_only & 1) && tid < kInvDoublesPerBlock / 2)
      cp_async16_u32(q_smem_u32(&Ibuf[(col0 >> 5) & 1][2 * tid]), Ip + (size_t)(col0 >> 5) * kInvDoublesPerBlock + 2 * tid);
    cp_async_commit();
    consumer_bar();

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
        const int t = warp + kCW * i;
        const int row = 8 * t + g;
        double k0 = 0.0, k1 = 0.0;
        if (invert) {
          const int col = col0 + nt * 8 + 2 * l;
          k0 = row == col ? 1.0 : 0.0;
          k1 = row == col + 1 ? 1.0 : 0.0;
        } else if ((t < ntl) && (row < n)) {
          double xr[D];
#pragma unroll
          for (int d = 0; d < D; ++d) xr[d] = xs[d * xstride + row];
          const double a_row = al[row];
          if (PMK_K3_X & 1) {
            k0 = xq0[0] + xr[0];
            k1 = xq1[0] + xr[0];
          } else {
            k0 = eval_kernel<D>(kp, xq0, xr);        // evalkernel(xq, X[i])   mixtureGP.jl:304
            k1 = eval_kernel<D>(kp, xq1, xr);
          }
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
      consumer_bar();
      if (tid < cnt) {
        double u = 0.0;
        for (int ww = 0; ww < kCW; ++ww) u += ured[ww * MQ + tid];
        pair_u[s_pair[tid]] = u;
      }
      cp_async_wait<0>();
      continue;     // the gather barrier of the next tile orders these reads before its writes
    }

    PMK_CYC(q_init = clock64() - q_total;)
    // ---- right-looking blocked TRSM:  s = L^-1 k  (mixtureGP.jl:311), ||s||^2 on the fly ----------
    double vacc[OT][2];
#pragma unroll
    for (int k = 0; k < OT; ++k) vacc[k][0] = vacc[k][1] = 0.0;
    unsigned exists = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i)
      if (warp + kCW * i < ntl) exists |= 1u << i;

    double* __restrict__ Pout = invert ? lt.P + lt.loff[p] : nullptr;
    for (int J = invert ? (col0 >> 5) : 0; J < nblk; ++J) {
      PMK_CYC(long long qc = clock64();)
      // 1. owners of block J's four row tiles publish W_J = -acc (their rows are final: W_J = L_JJ S_J)
      double* Wb = Wbuf[J & 1];
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        const int t = warp + kCW * i;
        if ((t >> 2) == J) {
          const int a = t & 3;
#pragma unroll
          for (int nt = 0; nt < NQT; ++nt)
            *reinterpret_cast<double2*>(&Wb[(8 * a + g) * LDQ + nt * 8 + 2 * l]) =
                make_double2(-acc[i][nt][0], -acc[i][nt][1]);
        }
      }
      cp_async_wait<0>();      // this thread's 16 bytes of inv(L_JJ)
      if (!(PMK_K3_X & 8)) consumer_bar();      // the ONLY block barrier per J (W and inv(L_JJ) are double-buffered)
      if (J + 1 < nblk && tid < kInvDoublesPerBlock / 2)
        cp_async16_u32(q_smem_u32(&Ibuf[(J + 1) & 1][2 * tid]), Ip + (size_t)(J + 1) * kInvDoublesPerBlock + 2 * tid);
      cp_async_commit();
      PMK_CYC({ long long c1 = clock64(); q_pub += c1 - qc; qc = c1; })
      // 2. side job, off the critical path: S_J = inv(L_JJ) * W_J only feeds ||s||^2 (4 x NQT tiles over the warps)
      const double2* Ib = reinterpret_cast<const double2*>(Ibuf[J & 1]) + lane;
#pragma unroll
      for (int k = 0; k < OT; ++k) {
        const int ot = warp + kCW * k;
        PMK_UNIFORM_IF(!(PMK_K3_X & 2) && ot < 4 * NQT) {
          const int a = ot / NQT, nt = ot % NQT;
          double s0 = 0.0, s1 = 0.0, r0 = 0.0, r1 = 0.0;
          const double* Cb = &Wb[l * LDQ + nt * 8 + g];
          const double2* Ia = Ib + (a * (a + 1) / 2) * 32;
          // b = 0 always; b = 1..a behind real (uniform) branches
          double2 f = Ia[0];
          dmma884(s0, s1, f.x, Cb[0]);
          dmma884(r0, r1, f.y, Cb[4 * LDQ]);
          PMK_UNIFORM_IF(a >= 1) {
            f = Ia[32];
            dmma884(s0, s1, f.x, Cb[8 * LDQ]);
            dmma884(r0, r1, f.y, Cb[12 * LDQ]);
            PMK_UNIFORM_IF(a >= 2) {
              f = Ia[64];
              dmma884(s0, s1, f.x, Cb[16 * LDQ]);
              dmma884(r0, r1, f.y, Cb[20 * LDQ]);
              PMK_UNIFORM_IF(a >= 3) {
                f = Ia[96];
                dmma884(s0, s1, f.x, Cb[24 * LDQ]);
                dmma884(r0, r1, f.y, Cb[28 * LDQ]);
              }
            }
          }
          s0 += r0;
          s1 += r1;
          vacc[k][0] = fma(s0, s0, vacc[k][0]);
          vacc[k][1] = fma(s1, s1, vacc[k][1]);
          PMK_UNIFORM_IF(invert) {
            // S_J tile (a, nt) = rows 32J+8a.., columns col0+8nt.. of inv(L): store it in the packed fragment-major layout
            const int rt = 4 * J + a, cth = (col0 >> 3) + nt;
            if (cth <= rt) {
              double* tp = Pout + (tri(rt) + cth) * 64;
              const int q0 = 2 * l, q1 = 2 * l + 1;
              tp[(g * 4 + (q0 & 3)) * 2 + (q0 >> 2)] = s0;
              tp[(g * 4 + (q1 & 3)) * 2 + (q1 >> 2)] = s1;
            }
          }
        }
      }
      PMK_CYC({ long long c1 = clock64(); q_diag += c1 - qc; qc = c1; })
      // 3. acc[I] += M_IJ * W_J for the row tiles below block J (M_IJ W_J = L_IJ S_J).  Tile guards are REAL branches
      //    (PMK_UNIFORM_IF); within a tile the DMMAs are ordered k-step-major so that consecutive
      //    ones hit different accumulators.
      unsigned active = 0;
#pragma unroll
      for (int i = 0; i < NT; ++i)
        if (warp + kCW * i >= 4 * J + 4) active |= 1u << i;
      active &= exists;
      if (J + 1 < nblk && active != 0) {
#pragma unroll 1
        for (int cg = 0; cg < NG; ++cg) {
#pragma unroll
          for (int ih = 0; ih < NSUB; ++ih) {
            constexpr unsigned SUBMASK = (1u << GI) - 1u;
            PMK_UNIFORM_IF(NSUB == 1 || (active & (SUBMASK << (ih * GI))) != 0) {
              q_mbar_wait(my_full + slot * 8, ph);              // this warp's operand group (J, cg, ih) has landed
              const double2* rs = ring + slot * (SLOT_BYTES / 16);
#pragma unroll
              for (int c = 0; c < CG; ++c) {
                const int ct = cg * CG + c;
                double bf[2][NQT];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                  for (int nt = 0; nt < NQT; ++nt) bf[ks][nt] = Wb[(8 * ct + 4 * ks + l) * LDQ + nt * 8 + g];
#pragma unroll
                for (int ii = 0; ii < GI; ++ii) {
                  const int i = ih * GI + ii;
                  PMK_UNIFORM_IF(active & (1u << i)) {
                    const double2 af = rs[(ii * CG + c) * 32];
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
      PMK_CYC(q_upd += clock64() - qc;)
    }

    // ---- reduce ||s||^2 and finish --------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < OT; ++k) {
      const int ot = warp + kCW * k;
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
    cp_async_wait<0>();
    consumer_bar();
    if (!invert && tid < cnt) {
      const int nt = tid >> 3, qi = tid & 7;
      double vs = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a) vs += vred[(a * NQT + nt) * 8 + qi];
      double u = 0.0;
      for (int ww = 0; ww < kCW; ++ww) u += ured[ww * MQ + tid];
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
    // (the next tile's gather writes s_xq / s_pair from the same threads that just read them; its block barrier orders
    //  every other shared buffer)
  }
}


// one translation unit per D (pmk_query_d{1,2,3}.cu) instantiates the size classes
template <int D, int NT, int NQT, int CG, int GI, int DEPTH>
static void launch_one(unsigned grid, const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp, int mean_only,
                       double* pu, double* pv, cudaStream_t s) {
  constexpr size_t dyn = (size_t)kCW * DEPTH * GI * CG * 512;
  static_assert(dyn <= 200 * 1024, "rings do not fit in shared memory");
  auto kern = k_query_pairs<D, NT, NQT, CG, GI, DEPTH>;
  const int n_sm = device_sm_count();
  static DeviceOnce once;
  once.run([&] { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn); });
  const unsigned ctas = grid < (unsigned)n_sm ? grid : (unsigned)n_sm;   // persistent: tiles are strided over the CTAs
  kern<<<ctas, kK3Threads, dyn, s>>>(lt, w, q, kp, mean_only, pu, pv);
}

// ring geometry of the two headline classes (overridable for tuning builds): column tiles per group, ring slots
#ifndef PMK_K3_CG0
#define PMK_K3_CG0 2
#endif
#ifndef PMK_K3_DEPTH0
#define PMK_K3_DEPTH0 2
#endif
#ifndef PMK_K3_CG1
#define PMK_K3_CG1 1
#endif
#ifndef PMK_K3_DEPTH1
#define PMK_K3_DEPTH1 3
#endif

template <int D>
void launch_pairs_d(int cls, unsigned grid, const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp,
                    int mean_only, double* pu, double* pv, cudaStream_t s) {
  // class:   0: n_pad <= 512, 32 pairs/tile | 1: <= 768, 24 | 2: <= 1024, 16 | 3: <= 1536, 8 | 4: <= 2048, 8
  // (the accumulators of n_pad x MQ doubles must fit the consumers' registers: NT * NQT * 4 per thread)
  // template arguments: NT row tiles per warp, NQT query tiles, CG column tiles x GI row tiles per operand group,
  // DEPTH ring slots per consumer warp
  if (cls == 0) launch_one<D, 6, 4, PMK_K3_CG0, 6, PMK_K3_DEPTH0>(grid, lt, w, q, kp, mean_only, pu, pv, s);
  else if (cls == 1) launch_one<D, 8, 3, PMK_K3_CG1, 8, PMK_K3_DEPTH1>(grid, lt, w, q, kp, mean_only, pu, pv, s);
  else if (cls == 2) launch_one<D, 11, 2, 1, 11, 3>(grid, lt, w, q, kp, mean_only, pu, pv, s);
  else if (cls == 3) launch_one<D, 16, 1, 1, 8, 3>(grid, lt, w, q, kp, mean_only, pu, pv, s);
  else launch_one<D, 22, 1, 1, 11, 3>(grid, lt, w, q, kp, mean_only, pu, pv, s);
}

}  // namespace pmk
