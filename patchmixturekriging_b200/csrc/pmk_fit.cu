// Fit path: fused Gram + batched blocked Cholesky (K1+K2) and the triangular solves for alpha.
//
// Replaces the body of fitmixtureGP! (reference src/RKHS/mixtureGP.jl:92-115):
//     U = constructkernelmatrix(X, θ); U[i,i] += σ²; c = U\y; L = cholesky(U).L
// One CTA per BSP leaf.  Left-looking blocked Cholesky with 32-column panels:
//   panel J:  C = K[J:, J] - L[J:, 0:J] * L[J, 0:J]^T
//             (k_gram_tiles has written the lower tiles of K + sigma2*I into the leaf's L tile slots, in the order the
//              accumulators want them; the update runs on DMMA.8x8x4 with both operands read as packed fragment
//              tiles, 512 B coalesced per warp load)
//             diagonal block: Cholesky + explicit 32x32 inverse by warp 0 (registers + shuffles)
//             L[J+1:, J] = C * inv(L_JJ)^T on DMMA, straight from the accumulators, stored once in packed-tile form.
// Several CTAs are resident per SM so that one leaf's serial diagonal-block phase overlaps the
// other leaves' DMMA phases.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "pmk_internal.cuh"

namespace pmk {

static constexpr unsigned kFull = 0xffffffffu;

// per-phase cycle counters of k_chol (warp 0 of every CTA), read by tools via pmk_debug_counters
__device__ unsigned long long g_chol_cycles[8];

static constexpr int LD = 36;   // smem row stride (doubles), == 4 (mod 16): fragment loads are conflict-free

// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void k_pack_leaves(LeafTable lt, const int64_t* __restrict__ leaf_off, const double* __restrict__ X,
                              const double* __restrict__ y) {
  const int p = blockIdx.x;
  const int n = lt.n[p], npad = lt.npad[p];
  const int64_t src = leaf_off[p], dst = lt.xoff[p];
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
#pragma unroll
    for (int d = 0; d < D; ++d) lt.xs[d * lt.xstride + dst + i] = i < n ? X[(src + i) * D + d] : 0.0;
    lt.y[dst + i] = i < n ? y[src + i] : 0.0;
    lt.alpha[dst + i] = 0.0;
  }
  if (threadIdx.x == 0) lt.info[p] = 0;
}

// ---------------------------------------------------------------------------------------------
// Cholesky + inverse of one 32x32 diagonal block by a single warp, in shared memory (lane = row
// for the factorisation, lane = column for the inverse).  Rolled loops, few registers: this is the
// serial phase of a panel and is hidden behind the DMMA phases of the other CTAs on the SM.
//   Dbuf (stride LDD): in = the block (lower part valid), out = L_JJ (upper zeroed)
//   Ibuf (stride LD) : out = inv(L_JJ) (upper zeroed)
// returns 0 or the (block-local, 1-based) order of the first non-positive pivot.
static constexpr int LDD = 33;
#ifndef PMK_FACTOR_UNROLLED
#define PMK_FACTOR_UNROLLED 1      /* measured (tools/factor_bench.cu, k_chol on C3): the rolled loop nest is slower, 40 k vs 26 k cycles alone */
#endif
__device__ __noinline__ int factor_block32(double* Dbuf, double* Ibuf, int lane) {
  // Step j does column j of the left-looking (dot-product form) Cholesky -- lane i >= j owns L[i][j] -- and,
  // fused into the same step, row j of X = inv(L) by forward substitution (lane c owns column c of X):
  //     x_jc = (delta_jc - sum_{k<j} L[j][k] x_kc) / L[j][j].
  // Both need row j of L (k < j), loaded once; the dependency chains are independent, so they overlap.
  // ROLLED loops, four partial sums per chain: the fully unrolled form of round 1 (4 k instructions of straight-line code,
  // executed once per call) ran at the speed of the instruction fetch -- 38 k cycles per block measured with the cycle
  // counters, with or without DMMA warps on the SM -- while this loop nest lives in the instruction cache.
  const double* rowp = Dbuf + lane * LDD;
  const double* xcol = Ibuf + lane;
  int info = 0;
#if PMK_FACTOR_UNROLLED
#pragma unroll
#else
#pragma unroll 1
#endif
  for (int j = 0; j < 32; ++j) {
    const double* lj = Dbuf + j * LDD;
    double s0 = rowp[j], s1 = 0.0, s2 = 0.0, s3 = 0.0;
    double t0 = (j == lane) ? 1.0 : 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    int k = 0;
#if PMK_FACTOR_UNROLLED
#pragma unroll
#else
#pragma unroll 1
#endif
    for (; k + 4 <= j; k += 4) {
      const double l0 = lj[k], l1 = lj[k + 1], l2 = lj[k + 2], l3 = lj[k + 3];
      s0 = fma(-rowp[k], l0, s0);
      s1 = fma(-rowp[k + 1], l1, s1);
      s2 = fma(-rowp[k + 2], l2, s2);
      s3 = fma(-rowp[k + 3], l3, s3);
      t0 = fma(-l0, xcol[k * LD], t0);
      t1 = fma(-l1, xcol[(k + 1) * LD], t1);
      t2 = fma(-l2, xcol[(k + 2) * LD], t2);
      t3 = fma(-l3, xcol[(k + 3) * LD], t3);
    }
    for (; k < j; ++k) {
      const double l0 = lj[k];
      s0 = fma(-rowp[k], l0, s0);
      t0 = fma(-l0, xcol[k * LD], t0);
    }
    const double s = (s0 + s1) + (s2 + s3);
    const double d = __shfl_sync(kFull, s, j);
    if (!(d > 0.0) && info == 0) info = j + 1;    // uniform: d is a broadcast
    const double inv = rsqrt(d);                  // 1/ajj (dpotf2 scales the column by 1/ajj)
    __syncwarp();
    Dbuf[lane * LDD + j] = lane > j ? s * inv : (lane == j ? d * inv : 0.0);
    Ibuf[j * LD + lane] = (j >= lane) ? ((t0 + t1) + (t2 + t3)) * inv : 0.0;
    __syncwarp();
  }
  return info;
}

// trailing-update inner loop of k_chol for NV valid row tiles (NV is warp-uniform).
// The warp's own operand tiles L[t_r, ct] stream through a per-warp cp.async ring kDepth-1 column tiles ahead
// (register-free prefetch: without it every iteration exposed an L2 round trip and a leaf took ~3.5 M cycles);
// the panel's row tiles L[J, ct] are shared by all warps of the CTA and come through L1.
#ifndef PMK_CHOL_DEPTH
#define PMK_CHOL_DEPTH 3
#endif
static constexpr int kCholDepth = PMK_CHOL_DEPTH;
// Register prefetch of the panel-row fragments one column tile ahead costs 16 registers at a cap of 80: with it ptxas spills
// 484 B (some of it inside the update loop), without it 280 B -- and the other five warps of the sub-partition hide the
// L1 / L2 latency anyway (C3 k_chol, same box: 11.55 -> 10.82 ms).
#ifndef PMK_CHOL_BPREFETCH
#define PMK_CHOL_BPREFETCH 0
#endif
#ifndef PMK_CHOL_FUSED_SOLVE
#define PMK_CHOL_FUSED_SOLVE 1
#endif
#ifndef PMK_CHOL_CPREFETCH
#define PMK_CHOL_CPREFETCH 1
#endif
// zv (optional): the forward-solve vector z = L^-1 y of the finished columns; zacc[r] then collects this lane's share of
// L[row tile r, 0:nct] z[0:8 nct] (rows g, columns l and 4 + l of every tile) on the way -- the A fragments are in registers anyway.
template <int R, int NV, bool WITH_Z = false>
__device__ __forceinline__ void chol_kloop(double (&acc)[R][4][2], const double2* __restrict__ Lp, const int (&bo)[4],
                                           const int (&ao)[R], int nct, uint32_t ring_u32, const double2* ring,
                                           const double* __restrict__ zv = nullptr, double* zacc = nullptr) {
  auto issue = [&](int ct, int slot) {
    if (ct < nct) {
#pragma unroll
      for (int r = 0; r < NV; ++r) cp_async16_u32(ring_u32 + (uint32_t)((slot * R + r) * 512), Lp + ao[r] + ct * 32);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < kCholDepth - 1; ++s) issue(s, s);
  int cslot = 0, fslot = kCholDepth - 1;
#if PMK_CHOL_BPREFETCH
  double2 bn[4];                                   // panel-row fragments, loaded one column tile ahead
#pragma unroll
  for (int b = 0; b < 4; ++b) bn[b] = nct > 0 ? Lp[bo[b]] : make_double2(0.0, 0.0);
#endif
  for (int ct = 0; ct < nct; ++ct) {
    double2 bf[4];
#if PMK_CHOL_BPREFETCH
#pragma unroll
    for (int b = 0; b < 4; ++b) bf[b] = bn[b];
    if (ct + 1 < nct) {
#pragma unroll
      for (int b = 0; b < 4; ++b) bn[b] = Lp[bo[b] + (ct + 1) * 32];
    }
#else
#pragma unroll
    for (int b = 0; b < 4; ++b) bf[b] = Lp[bo[b] + ct * 32];
#endif
    cp_async_wait<kCholDepth - 2>();
    const double2* rs = ring + cslot * (R * 32);
    issue(ct + kCholDepth - 1, fslot);      // refill the slot consumed one iteration ago
    fslot = cslot;
    cslot = (cslot + 1 == kCholDepth) ? 0 : cslot + 1;
    double zl = 0.0, zh = 0.0;
    if (WITH_Z) {
      zl = __ldcg(zv + 8 * ct + (threadIdx.x & 3));          // written earlier in this launch by another warp: read through L2
      zh = __ldcg(zv + 8 * ct + 4 + (threadIdx.x & 3));
    }
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      const double2 af = rs[r * 32];
#pragma unroll
      for (int b = 0; b < 4; ++b) dmma884(acc[r][b][0], acc[r][b][1], af.x, bf[b].x);
#pragma unroll
      for (int b = 0; b < 4; ++b) dmma884(acc[r][b][0], acc[r][b][1], af.y, bf[b].y);
      if (WITH_Z) zacc[r] = fma(af.y, zh, fma(af.x, zl, zacc[r]));
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// K1 for the fit: the lower tiles of K + sigma2*I (identity on the padding) written into the leaf's L tile
// slots in C-FRAGMENT-major order (lane (g,l) holds {M[g][2l], M[g][2l+1]}), i.e. exactly as k_chol's
// accumulators want them.  A tile slot goes K (here) -> raw panel block C (k_chol, parked) -> L (final,
// A-fragment-major).  Full-occupancy elementwise kernel: the evaluations (sqrt + exp chains) run at FP64
// pipe throughput here instead of stalling the low-occupancy factorisation (measured: evaluating them
// inside k_chol took 31 % of its cycles).  One warp per row tile.
// The leaf's points are staged ONCE per CTA in shared memory by 1-D TMA bulk copies (cp.async.bulk -> UBLKCP, one per
// coordinate, completing on an mbarrier; north star (1): "points staged in shared memory via TMA"); every evaluation then
// reads its column point with broadcast LDS instead of going through L1.
#ifndef PMK_GRAMT_MINB
#define PMK_GRAMT_MINB 4      // 64 registers instead of 110: 2 -> 4 resident CTAs per SM (C3: 1.80 -> 1.19 ms; 5: 1.17 with more spills)
#endif
__device__ __forceinline__ uint32_t f_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void f_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(f_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(f_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void f_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "F_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra F_DONE;\n"
      "bra F_WAIT;\n"
      "F_DONE:\n"
      "}\n" ::"r"(f_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int D>
__global__ void __launch_bounds__(256, PMK_GRAMT_MINB)
k_gram_tiles(LeafTable lt, const int* __restrict__ order, KParams kp, double sigma2, int smem_npad) {
  __shared__ double s_exp[64];
  __shared__ __align__(8) uint64_t s_bar;
  extern __shared__ __align__(128) unsigned char pmk_gram_smem[];
  double* sx = reinterpret_cast<double*>(pmk_gram_smem);          // [D][smem_npad]
  const int p = order[blockIdx.x];
  const int n = lt.n[p], npad = lt.npad[p], ntl = npad >> 3;
  if ((int)blockIdx.y * 8 >= ntl) return;                          // whole CTA past the leaf's last row tile
  if (threadIdx.x < 64) s_exp[threadIdx.x] = c_exp2_64[threadIdx.x];
  if (threadIdx.x == 0) {
    f_mbar_init(&s_bar, 1);
    // only the columns this CTA's row tiles can meet: 0 .. 8 (8 blockIdx.y + 8) - 1
    const int ncol = min(npad, 64 * ((int)blockIdx.y + 1));
    f_mbar_expect_tx(&s_bar, (uint32_t)(D * ncol * 8));
#pragma unroll
    for (int d = 0; d < D; ++d) f_bulk_g2s(sx + d * smem_npad, lt.xs + d * lt.xstride + lt.xoff[p], (uint32_t)(ncol * 8), &s_bar);
  }
  __syncthreads();
  f_mbar_wait(&s_bar, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.y * 8 + warp;
  if (t >= ntl) return;
  const int g = lane >> 2, l = lane & 3;
  double2* Lp = reinterpret_cast<double2*>(lt.L + lt.loff[p]) + tri(t) * 32 + lane;
  const int row = 8 * t + g;
  double xr[D];
#pragma unroll
  for (int d = 0; d < D; ++d) xr[d] = sx[d * smem_npad + row];
  // one entry, every case: lower triangle of K + sigma2*I, identity on the padding
  auto entry = [&](int col) {
    double v = 0.0;
    if (col <= row) {
      if (row < n) {          // col <= row < n
        double xc[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xc[d] = sx[d * smem_npad + col];
        v = eval_kernel<D>(kp, xr, xc);               // evalkernel(X[i], X[j]), i >= j  (RKHS.jl:21-25)
        if (row == col) v = __dadd_rn(v, sigma2);     // mixtureGP.jl:102-104
      } else {
        v = (row == col) ? 1.0 : 0.0;                 // identity padding
      }
    }
    return v;
  };
  // Squared exponential, tiles strictly below the diagonal, real rows: no case distinctions, exp(-a |x - z|^2) with the
  // same table-driven exp as the pair kernel's cross-covariance (<= 2 ulp from the sqrt / re-square / libm-exp form of
  // kernel.jl:350-357), four column tiles = eight independent evaluations per thread in flight.
  int c = 0;
  if (kp.kind == PMK_KERNEL_SQEXP && 8 * t + 7 < n) {
    for (; c + 4 <= t; c += 4) {
      double arg[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int col = 8 * (c + (k >> 1)) + 2 * l + (k & 1);
        double s2 = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const double dd = xr[d] - sx[d * smem_npad + col];
          s2 = fma(dd, dd, s2);
        }
        arg[k] = -kp.p * s2;
      }
#pragma unroll
      for (int k = 0; k < 8; k += 2)
        Lp[(c + (k >> 1)) * 32] = make_double2(exp_neg_tab(arg[k], s_exp), exp_neg_tab(arg[k + 1], s_exp));
    }
#ifndef PMK_GRAMT_NO_TAIL
    for (; c < t; ++c) {          // the up to three remaining tiles strictly below the diagonal, same arithmetic
      double arg[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int col = 8 * c + 2 * l + k;
        double s2 = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const double dd = xr[d] - sx[d * smem_npad + col];
          s2 = fma(dd, dd, s2);
        }
        arg[k] = -kp.p * s2;
      }
      Lp[c * 32] = make_double2(exp_neg_tab(arg[0], s_exp), exp_neg_tab(arg[1], s_exp));
    }
#endif
  }
  for (; c <= t; ++c) Lp[c * 32] = make_double2(entry(8 * c + 2 * l), entry(8 * c + 2 * l + 1));
}

// ---------------------------------------------------------------------------------------------
// K2: blocked left-looking Cholesky of one leaf per CTA (see the file header).  Per 32-column panel J:
//   A. warps 0-3: diagonal block  D = K_JJ - L[J,0:J] L[J,0:J]^T  -> shared memory (named barrier of 4 warps)
//   B. warp 0 factors + inverts D (serial, latency-bound) and raises a shared-memory flag WHILE all other warps (and
//      warp 0 afterwards) pull row-tile groups off a shared counter and compute C = K[t,J] - L[t,0:J] L[J,0:J]^T on DMMA;
//      a group keeps C in its accumulators, waits for the flag and stores L[t,J] = C inv(L_JJ)^T (DMMA), final
//      A-fragment-major tiles (PMK_CHOL_FUSED_SOLVE; with 0 the raw C tiles are parked in their L slots and
//   C. after a CTA barrier all warps compute L[t,J] = C inv(L_JJ)^T from the parked tiles -- the round-1 kernel).
// Lookahead of the diagonal block (measured on C3: k_chol 11.68 -> 11.41 ms; with the diagonal block's update removed
// altogether -- wrong results, timing only -- 10.13 ms: the kernel is bound by the left-looking re-reads of the finished
// columns, 34 GB per C3 fit, not by the length of the per-panel critical path).
// Loads of tiles that are read once per panel (the parked K / C tiles): L2 only, so they do not evict the panel's row tiles
// L[J, 0:J] -- shared by all warps of the CTA and re-read for every unit -- from L1.
#ifndef PMK_CHOL_LDCG
#define PMK_CHOL_LDCG 1
#endif
__device__ __forceinline__ double2 ld_once(const double2* p) {
#if PMK_CHOL_LDCG
  return __ldcg(p);
#else
  return *p;
#endif
}
#ifndef PMK_CHOL_LOOKAHEAD
#define PMK_CHOL_LOOKAHEAD 1
#endif
#ifndef PMK_CHOL_MINB
#define PMK_CHOL_MINB 3      // resident leaves per SM the register allocation targets (4 forces 64 registers: measured below)
#endif
template <int NW, int R, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
k_chol(LeafTable lt, const int* __restrict__ order) {
  __shared__ double Dbuf[32 * LDD];
  __shared__ double Ibuf[32 * LD];
  __shared__ double s_zp[2][32];     // forward solve of alpha: L[block, older panels] z, folded in by the lookahead warps (by block parity)
  __shared__ double s_zn[32];        // ... and L[block, newest panel] z from phase A
  __shared__ int s_fail, s_next, s_next2[2], s_ready;
  const int p = order[blockIdx.x];
  const int npad = lt.npad[p];
  const int nblk = npad >> 5, ntl = npad >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, l = lane & 3;
  double2* Lp = reinterpret_cast<double2*>(lt.L + lt.loff[p]);
  double2* Ip = reinterpret_cast<double2*>(lt.Linv + lt.ioff[p]);
  if (threadIdx.x == 0) {
    s_fail = 0;
    s_next = 4;
    s_next2[0] = 4;
    s_next2[1] = 8;
    s_ready = 0;
  }
  if (threadIdx.x < 64) s_zp[threadIdx.x >> 5][threadIdx.x & 31] = 0.0;
  if (threadIdx.x < 32) s_zn[threadIdx.x] = 0.0;
  __syncthreads();
  // z = L^-1 y (the forward half of the solve for alpha) is formed panel by panel on the operand tiles the factorisation streams
  // anyway: z_J = inv(L_JJ) (y_J - L[J, 0:J] z[0:J]); it lives in lt.alpha until k_solve_alpha's backward sweep overwrites it.
  const int64_t xo = lt.xoff[p];
  const double* zv = lt.alpha + xo;
  extern __shared__ __align__(16) unsigned char pmk_chol_smem[];
  const double2* ring = reinterpret_cast<const double2*>(pmk_chol_smem) + (size_t)warp * (kCholDepth * R * 32) + lane;
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
  const int src_lo = (lane & ~3) | (l >> 1);
  const int src_hi = (lane & ~3) | (2 + (l >> 1));

  // L[t, J] = C inv(L_JJ)^T for one row tile: cf = the raw block C in C-fragment order (four column tiles), trow = the
  // row tile's slots of panel J; inv(L_JJ) is read from Ibuf.  Final tiles are stored A-fragment-major.
  auto solve_store = [&](const double2 (&cf)[4], double2* trow) {
    double alo[4], ahi[4];
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {     // C-fragment (cols 2l, 2l+1) -> A-fragment (cols l, l+4) inside the quad
      const double v0 = __shfl_sync(kFull, cf[kb].x, src_lo);
      const double v1 = __shfl_sync(kFull, cf[kb].y, src_lo);
      const double w0 = __shfl_sync(kFull, cf[kb].x, src_hi);
      const double w1 = __shfl_sync(kFull, cf[kb].y, src_hi);
      alo[kb] = (l & 1) ? v1 : v0;
      ahi[kb] = (l & 1) ? w1 : w0;
    }
    double* tile_row = reinterpret_cast<double*>(trow);
    double o0[4], o1[4], p0[4], p1[4];
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) o0[cb] = o1[cb] = p0[cb] = p1[cb] = 0.0;
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {     // kb outermost: the four output tiles advance in lock step
#pragma unroll
      for (int cb = kb; cb < 4; ++cb) dmma884(o0[cb], o1[cb], alo[kb], Ibuf[(8 * cb + g) * LD + 8 * kb + l]);
#pragma unroll
      for (int cb = kb; cb < 4; ++cb) dmma884(p0[cb], p1[cb], ahi[kb], Ibuf[(8 * cb + g) * LD + 8 * kb + 4 + l]);
    }
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      // C-fragment (row g, cols 2l, 2l+1) -> packed A-fragment-major tile
      double* tile = tile_row + cb * 64;
      const int q0 = 2 * l, q1 = 2 * l + 1;
      tile[(g * 4 + (q0 & 3)) * 2 + (q0 >> 2)] = o0[cb] + p0[cb];
      tile[(g * 4 + (q1 & 3)) * 2 + (q1 >> 2)] = o1[cb] + p1[cb];
    }
  };

  PMK_CYC(long long c_total = clock64(), c_diag = 0, c_work = 0, c_factor = 0, c_solve = 0, c_wait = 0;)
  for (int J = 0; J < nblk; ++J) {
    const int t0 = 4 * J;
    int bo[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) bo[b] = (int)tri(t0 + b) * 32 + lane;
    PMK_CYC(long long c0 = clock64();)
    // ---- A: diagonal block ----------------------------------------------------------------------
    if (warp < 4) {
      double acc[R][4][2];
      int ao[R];
#pragma unroll
      for (int r = 0; r < R; ++r) ao[r] = (int)tri(t0 + warp) * 32 + lane;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double2 kt = (b <= warp) ? ld_once(Lp + ao[0] + (t0 + b) * 32) : make_double2(0.0, 0.0);
        acc[0][b][0] = -kt.x;
        acc[0][b][1] = -kt.y;
      }
      // Only the newest panel's four column tiles are left to subtract: the contributions of panels 0 .. J-2 were folded
      // into the parked tiles during panel J-1's phase B (lookahead below), off the critical path.
      double zacc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) zacc[r] = 0.0;
      if (!PMK_CHOL_LOOKAHEAD) {
        chol_kloop<R, 1, true>(acc, Lp, bo, ao, 4 * J, ring_u32, ring, zv, zacc);
      } else if (J > 0) {
        int bo2[4], ao2[R];
#pragma unroll
        for (int b = 0; b < 4; ++b) bo2[b] = bo[b] + (t0 - 4) * 32;
#pragma unroll
        for (int r = 0; r < R; ++r) ao2[r] = ao[r] + (t0 - 4) * 32;
        chol_kloop<R, 1, true>(acc, Lp, bo2, ao2, 4, ring_u32, ring, zv + 8 * (t0 - 4), zacc);
      }
      {
        double zs = zacc[0];
        zs += __shfl_xor_sync(kFull, zs, 1);
        zs += __shfl_xor_sync(kFull, zs, 2);
        if (l == 0) s_zn[8 * warp + g] = zs;
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        Dbuf[(8 * warp + g) * LDD + 8 * b + 2 * l + 0] = -acc[0][b][0];
        Dbuf[(8 * warp + g) * LDD + 8 * b + 2 * l + 1] = -acc[0][b][1];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");      // the four diagonal-row warps only
      PMK_CYC({ long long c1 = clock64(); c_diag += c1 - c0; c0 = c1; })
      if (warp == 0) {
        const int info = factor_block32(Dbuf, Ibuf, lane);
        if (info != 0) {
          if (lane == 0) {
            lt.info[p] = 32 * J + info;
            s_fail = 1;
          }
        } else {
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            for (int b = 0; b <= a; ++b) {
              const int rd = (8 * a + g) * LDD + 8 * b + l;
              const int ri = (8 * a + g) * LD + 8 * b + l;
              Lp[(tri(t0 + a) + t0 + b) * 32 + lane] = make_double2(Dbuf[rd], Dbuf[rd + 4]);
              Ip[(size_t)J * (kInvTilesPerBlock * 32) + (a * (a + 1) / 2 + b) * 32 + lane] =
                  make_double2(Ibuf[ri], Ibuf[ri + 4]);
            }
          }
        }
#if PMK_CHOL_FUSED_SOLVE
        __syncwarp();
        __threadfence_block();
        if (lane == 0) *(volatile int*)&s_ready = J + 1;     // inv(L_JJ) is in Ibuf (or s_fail is set): the units' epilogues may run
#endif
        if (info == 0) {      // z_J = inv(L_JJ) (y_J - L[J, 0:J] z): row `lane` of the lower-triangular inverse
          const double older = PMK_CHOL_LOOKAHEAD ? s_zp[J & 1][lane] : 0.0;
          const double rl = lt.y[xo + 32 * J + lane] - older - s_zn[lane];
          double zz = 0.0;
#pragma unroll 8
          for (int k = 0; k < 32; ++k) zz = fma(Ibuf[lane * LD + k], __shfl_sync(kFull, rl, k), zz);   // the inverse's upper part is zero
          lt.alpha[xo + 32 * J + lane] = zz;
        }
        PMK_CYC({ long long c1 = clock64(); c_factor += c1 - c0; c0 = c1; })
      }
    }
    // ---- lookahead: the NEXT panel's diagonal block minus the contributions of the finished panels 0 .. J-1, by two of
    // the warps that have no part in phase A; parked in the block's own tile slots (C-fragment-major, like K)
    if (PMK_CHOL_LOOKAHEAD && J > 0 && J + 1 < nblk && (warp == 4 || warp == 5)) {
      static_assert(R == 2, "the lookahead deals the next diagonal block as two pairs of row tiles");
      const int tb = t0 + 4 + 2 * (warp - 4);
      int bo2[4], ao[R];
#pragma unroll
      for (int b = 0; b < 4; ++b) bo2[b] = (int)tri(t0 + 4 + b) * 32 + lane;
      double acc[R][4][2];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        ao[r] = (int)tri(tb + r) * 32 + lane;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double2 kt = (t0 + 4 + b <= tb + r) ? ld_once(Lp + ao[r] + (t0 + 4 + b) * 32) : make_double2(0.0, 0.0);
          acc[r][b][0] = -kt.x;
          acc[r][b][1] = -kt.y;
        }
      }
      double zacc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) zacc[r] = 0.0;
      chol_kloop<R, R, true>(acc, Lp, bo2, ao, 4 * J, ring_u32, ring, zv, zacc);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        double zs = zacc[r];
        zs += __shfl_xor_sync(kFull, zs, 1);
        zs += __shfl_xor_sync(kFull, zs, 2);
        if (l == 0) s_zp[(J + 1) & 1][8 * (2 * (warp - 4) + r) + g] = zs;      // rows of block J + 1, columns of panels 0 .. J-1
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (t0 + 4 + b <= tb + r) Lp[ao[r] + (t0 + 4 + b) * 32] = make_double2(-acc[r][b][0], -acc[r][b][1]);
    }
    // ---- B: off-diagonal row tiles, dynamically dealt; raw C parked in the tile slots -------------
    for (;;) {
      int tb = 0;
#if PMK_CHOL_FUSED_SOLVE
      if (lane == 0) tb = atomicAdd(&s_next2[J & 1], R);
#else
      if (lane == 0) tb = atomicAdd(&s_next, R);
#endif
      tb = __shfl_sync(kFull, tb, 0);
      if (tb >= ntl) break;
      double acc[R][4][2];
      int ao[R];
      int nv = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool tv = tb + r < ntl;
        nv += tv ? 1 : 0;
        ao[r] = (int)tri(tv ? tb + r : tb) * 32 + lane;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double2 kt = tv ? ld_once(Lp + ao[r] + (t0 + b) * 32) : make_double2(0.0, 0.0);
          acc[r][b][0] = -kt.x;
          acc[r][b][1] = -kt.y;
        }
      }
      if (nv == R) chol_kloop<R, R>(acc, Lp, bo, ao, 4 * J, ring_u32, ring);
      else chol_kloop<R, 1>(acc, Lp, bo, ao, 4 * J, ring_u32, ring);     // R == 2: one valid tile
#if PMK_CHOL_FUSED_SOLVE
      // The raw block never leaves the registers: as soon as warp 0 has published inv(L_JJ) -- long before a unit's update loop
      // ends, except in the first panels -- the unit multiplies it in and stores the FINAL tiles (no parked C: one write and
      // one read of the whole factor less per fit, and one CTA barrier per panel instead of two).
      if (lane == 0)
        while (*(volatile int*)&s_ready <= J) __nanosleep(40);
      __syncwarp();
      __threadfence_block();
      if (*(volatile int*)&s_fail) break;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r < nv) {
          double2 cf[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) cf[b] = make_double2(-acc[r][b][0], -acc[r][b][1]);
          solve_store(cf, Lp + (tri(tb + r) + t0) * 32);
        }
      }
#else
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r < nv) {
#pragma unroll
          for (int b = 0; b < 4; ++b) Lp[ao[r] + (t0 + b) * 32] = make_double2(-acc[r][b][0], -acc[r][b][1]);
        }
      }
#endif
    }
    PMK_CYC({ long long c1 = clock64(); c_work += c1 - c0; c0 = c1; })
#if PMK_CHOL_FUSED_SOLVE
    if (threadIdx.x == 0) s_next2[(J + 1) & 1] = t0 + 8;      // panel J+1's first off-diagonal row tile (that counter is idle during panel J)
    __syncthreads();     // panel J complete and visible before panel J+1 reads it; Ibuf free for the next factor
    PMK_CYC({ long long c1 = clock64(); c_wait += c1 - c0; c0 = c1; })
    if (s_fail) return;
    continue;
#endif
    __syncthreads();     // inverse block ready, every C tile parked
    PMK_CYC({ long long c1 = clock64(); c_wait += c1 - c0; c0 = c1; })
    if (s_fail) return;
    // ---- C: L[t, J] = C * inv(L_JJ)^T -----------------------------------------------------------
    {
      int t = t0 + 4 + warp;
      double2 cf[4];
#if PMK_CHOL_CPREFETCH
      double2 cn[4];
      if (t < ntl) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) cf[kb] = ld_once(Lp + (tri(t) + t0 + kb) * 32 + lane);
      }
#endif
      for (; t < ntl; t += NW) {
        double2* trow = Lp + (tri(t) + t0) * 32;
#if PMK_CHOL_CPREFETCH
        if (t + NW < ntl) {               // next row tile's parked block, one iteration ahead
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) cn[kb] = ld_once(Lp + (tri(t + NW) + t0 + kb) * 32 + lane);
        }
#else
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) cf[kb] = ld_once(trow + kb * 32 + lane);
#endif
        solve_store(cf, trow);
#if PMK_CHOL_CPREFETCH
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) cf[kb] = cn[kb];
#endif
      }
    }
    if (threadIdx.x == 0) s_next = t0 + 8;     // first off-diagonal row tile of the next panel
    PMK_CYC({ long long c1 = clock64(); c_solve += c1 - c0; c0 = c1; })
    __syncthreads();     // panel J complete and visible before panel J+1 reads it
    PMK_CYC({ long long c1 = clock64(); c_wait += c1 - c0; })
  }
#ifdef PMK_PROFILE_CYCLES
  if (threadIdx.x == 0) {
    atomicAdd(&g_chol_cycles[0], (unsigned long long)(clock64() - c_total));
    atomicAdd(&g_chol_cycles[1], (unsigned long long)c_diag);
    atomicAdd(&g_chol_cycles[2], (unsigned long long)c_work);
    atomicAdd(&g_chol_cycles[3], (unsigned long long)c_factor);
    atomicAdd(&g_chol_cycles[4], (unsigned long long)c_solve);
    atomicAdd(&g_chol_cycles[5], (unsigned long long)c_wait);
    atomicAdd(&g_chol_cycles[6], 1ull);
  }
#endif
}

// ---------------------------------------------------------------------------------------------
// K2, level-synchronous form (leaves of n_pad >= kCholLevelsMinNpad; the one-CTA-per-leaf kernel above serves the smaller ones).
// All leaves advance panel by panel -- 64 columns at a time -- and two kinds of launches alternate, so that the serial chain of a
// diagonal block and the DMMA work never share an SM:
//     k_chol_factor64 (Jp)  : one warp per leaf on the 64x64 diagonal block [D00; D10 D11] the panel kernels have kept up to date:
//                             L00 = chol(D00), L10 = D10 inv(L00)^T, L11 = chol(D11 - L10 L10^T), both 32x32 inverses, and the two
//                             blocks of z = L^-1 y (the forward half of the solve for alpha);
//     k_chol_panel64 (Jp)   : for the row tiles below the block: [C0 C1] = K[t, 64 columns] - L[t, 0:8Jp) L[Jp rows, 0:8Jp)^T with eight
//                             accumulator tiles per row tile, then X0 = C0 inv(L00)^T, X1 = (C1 - X0 L10^T) inv(L11)^T, stored once as
//                             final packed tiles; plus the RIGHT-LOOKING extras that leave k_chol_factor64 nothing to accumulate:
//                             a CTA = one 64-row block I of the leaf, its four warps publish their final tiles in shared memory (over
//                             the dead operand buffers) and subtract X X^T from the block's 36 diagonal tiles (still holding K_II
//                             minus the earlier panels, C-fragment-major as k_gram_tiles wrote them), nine tiles per warp, and
//                             S_I += X z_J goes to lt.alpha.
// k_chol_panel64 runs the loop of the pair kernel -- operands in SHARED memory, moved by TMA: one CTA = four compute warps + one
// producer warp, all on one (leaf, 64-row block below panel Jp); warp w owns row tiles tb0 + 2w, + 1.
//  * B operand = the eight row tiles of the panel's own block, L[Jp rows, 0:8Jp): the producer stages them in chunks of KC column
//    tiles with 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) into a two-deep ring; the compute warps walk the chunks in lock step
//    (they all have the same loop length) and hand a buffer back through an mbarrier.
//  * A operand = the warp's own rows, contiguous in the packed layout, streamed by its lane 0 through a private ring of bulk copies
//    (four column tiles x two rows per slot); the two raw K blocks of the panel follow the finished columns in the same rows, so the
//    last two chunks of the stream deliver them and C = K - sum is formed at the end: nothing is loaded in front of the loop.
//  * inner loop per column tile: 2 + 8 LDS.128, 32 DMMA.
// History and measurements (32-column panels, CTA shapes, a persistent form, what bounds it): profiles/chol_r02_notes.md.
__device__ __forceinline__ void f_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(f_smem_u32(bar)) : "memory");
}
static constexpr int kPanCW = 4;      // column tiles per chunk of a warp's own rows
static constexpr int kPanDepth = 2;   // slots of a warp's own-row ring
#ifndef PMK_P64_KC
#define PMK_P64_KC 2         // column tiles per staged chunk of the diagonal block's eight row tiles
#endif
#ifndef PMK_P64_MINB
#define PMK_P64_MINB 3
#endif
static constexpr int kP64Tiles = 36;     // staged with the inverses: inv(L00) (10 tiles), inv(L11) (10), L10 (16)

__global__ void __launch_bounds__(32)
k_chol_factor64(LeafTable lt, const int* __restrict__ order, int Jp, int with_z) {
  __shared__ double Dbuf[32 * LDD];
  __shared__ double I0[32 * LD];
  __shared__ double I1[32 * LD];
  __shared__ double rbuf[32];
  const int p = order[blockIdx.x];
  const int ntl = lt.npad[p] >> 3;
  const int t0 = 8 * Jp;
  if (t0 >= ntl || lt.info[p] != 0) return;
  const int lane = threadIdx.x;
  const int g = lane >> 2, l = lane & 3;
  double2* Lp = reinterpret_cast<double2*>(lt.L + lt.loff[p]);
  double2* Ip = reinterpret_cast<double2*>(lt.Linv + lt.ioff[p]);
  const int64_t xo = lt.xoff[p];
  // a factored block (dense in Dbuf / Ib) -> packed tiles of L and of Linv block J
  auto write_block = [&](const double* Ib, int tt, int J) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      for (int b = 0; b <= a; ++b) {
        const int rd = (8 * a + g) * LDD + 8 * b + l;
        const int ri = (8 * a + g) * LD + 8 * b + l;
        Lp[(tri(tt + a) + tt + b) * 32 + lane] = make_double2(Dbuf[rd], Dbuf[rd + 4]);
        Ip[(size_t)J * (kInvTilesPerBlock * 32) + (a * (a + 1) / 2 + b) * 32 + lane] = make_double2(Ib[ri], Ib[ri + 4]);
      }
    }
  };
  // ---- first 32x32 block
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) {
      const double2 d = ld_once(Lp + (tri(t0 + a) + t0 + b) * 32 + lane);      // C-fragment-major: {M[g][2l], M[g][2l+1]}
      Dbuf[(8 * a + g) * LDD + 8 * b + 2 * l] = d.x;
      Dbuf[(8 * a + g) * LDD + 8 * b + 2 * l + 1] = d.y;
    }
  __syncwarp();
  int info = factor_block32(Dbuf, I0, lane);
  if (info != 0) {
    if (lane == 0) lt.info[p] = 8 * t0 + info;
    return;
  }
  write_block(I0, t0, 2 * Jp);
  double z0 = 0.0;
  if (with_z) {
    const double rl = lt.y[xo + 8 * t0 + lane] - lt.alpha[xo + 8 * t0 + lane];
#pragma unroll 8
    for (int k = 0; k < 32; ++k) z0 = fma(I0[lane * LD + k], __shfl_sync(kFull, rl, k), z0);      // the inverse's upper part is zero
    lt.alpha[xo + 8 * t0 + lane] = z0;
  }
  if (t0 + 4 >= ntl) return;
  // ---- L10 = D10 inv(L00)^T, final tiles stored; S1 += L10 z0
  const int src_lo = (lane & ~3) | (l >> 1);
  const int src_hi = (lane & ~3) | (2 + (l >> 1));
  double2 xa[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double alo[4], ahi[4];
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      const double2 cf = ld_once(Lp + (tri(t0 + 4 + a) + t0 + kb) * 32 + lane);
      const double v0 = __shfl_sync(kFull, cf.x, src_lo);
      const double v1 = __shfl_sync(kFull, cf.y, src_lo);
      const double w0 = __shfl_sync(kFull, cf.x, src_hi);
      const double w1 = __shfl_sync(kFull, cf.y, src_hi);
      alo[kb] = (l & 1) ? v1 : v0;
      ahi[kb] = (l & 1) ? w1 : w0;
    }
    double o0[4], o1[4], p0[4], p1[4];
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) o0[cb] = o1[cb] = p0[cb] = p1[cb] = 0.0;
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
      for (int cb = kb; cb < 4; ++cb) dmma884(o0[cb], o1[cb], alo[kb], I0[(8 * cb + g) * LD + 8 * kb + l]);
#pragma unroll
      for (int cb = kb; cb < 4; ++cb) dmma884(p0[cb], p1[cb], ahi[kb], I0[(8 * cb + g) * LD + 8 * kb + 4 + l]);
    }
    double zpart = 0.0;
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      const double f0 = o0[cb] + p0[cb], f1 = o1[cb] + p1[cb];       // row g, columns 2l, 2l + 1
      zpart = fma(f1, __shfl_sync(kFull, z0, 8 * cb + 2 * l + 1), fma(f0, __shfl_sync(kFull, z0, 8 * cb + 2 * l), zpart));
      const double v0 = __shfl_sync(kFull, f0, src_lo);
      const double v1 = __shfl_sync(kFull, f1, src_lo);
      const double w0 = __shfl_sync(kFull, f0, src_hi);
      const double w1 = __shfl_sync(kFull, f1, src_hi);
      xa[a][cb] = make_double2((l & 1) ? v1 : v0, (l & 1) ? w1 : w0);
      Lp[(tri(t0 + 4 + a) + t0 + cb) * 32 + lane] = xa[a][cb];
    }
    zpart += __shfl_xor_sync(kFull, zpart, 1);
    zpart += __shfl_xor_sync(kFull, zpart, 2);
    if (l == 0) rbuf[8 * a + g] = zpart;
  }
  // ---- second block: D11 - L10 L10^T
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) {
      const double2 d = ld_once(Lp + (tri(t0 + 4 + a) + t0 + 4 + b) * 32 + lane);
      double o0 = 0.0, o1 = 0.0, p0 = 0.0, p1 = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        dmma884(o0, o1, xa[a][c].x, xa[b][c].x);
        dmma884(p0, p1, xa[a][c].y, xa[b][c].y);
      }
      Dbuf[(8 * a + g) * LDD + 8 * b + 2 * l] = d.x - (o0 + p0);
      Dbuf[(8 * a + g) * LDD + 8 * b + 2 * l + 1] = d.y - (o1 + p1);
    }
  __syncwarp();
  info = factor_block32(Dbuf, I1, lane);
  if (info != 0) {
    if (lane == 0) lt.info[p] = 8 * t0 + 32 + info;
    return;
  }
  write_block(I1, t0 + 4, 2 * Jp + 1);
  if (with_z) {
    const double rl = lt.y[xo + 8 * t0 + 32 + lane] - lt.alpha[xo + 8 * t0 + 32 + lane] - rbuf[lane];
    double z1 = 0.0;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) z1 = fma(I1[lane * LD + k], __shfl_sync(kFull, rl, k), z1);
    lt.alpha[xo + 8 * t0 + 32 + lane] = z1;
  }
}

template <int NW>
__global__ void __launch_bounds__((NW + 1) * 32, PMK_P64_MINB)
k_chol_panel64(LeafTable lt, const int* __restrict__ order, int Jp) {
  constexpr int R = 2, CW = 4, KC = PMK_P64_KC;
  static_assert(NW == 4, "one CTA = one 64-row block = four warps of two row tiles");
  static_assert(CW % KC == 0, "a B chunk never straddles two A chunks");
  constexpr int kRing = kPanDepth * R * CW * 512;
  static_assert(2 * 8 * KC * 512 + NW * kRing >= 8 * 8 * 512, "the staged final tiles reuse the operand buffers");
  extern __shared__ __align__(128) unsigned char pmk_chol_smem[];
  __shared__ __align__(8) uint64_t bfull[2], bempty[2], ifull, afull[NW * kPanDepth];
  const int p = order[blockIdx.y];
  const int ntl = lt.npad[p] >> 3;
  const int t0 = 8 * Jp;
  const int tb0 = t0 + 8 + blockIdx.x * (NW * R);
  if (tb0 >= ntl || lt.info[p] != 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_active = min(NW, (ntl - tb0) >> 1);          // 4, or 2 for a half block at the end of the leaf
  unsigned char* Bbuf = pmk_chol_smem;                      // [2][8][KC][512]
  unsigned char* Aring = Bbuf + 2 * 8 * KC * 512;           // [NW][kPanDepth][R][CW][512]
  unsigned char* Ibuf = Aring + NW * kRing;                 // [36][512]: inv(L00), inv(L11), L10
  if (threadIdx.x == 0) {
    for (int k = 0; k < 2; ++k) {
      f_mbar_init(&bfull[k], 1);
      f_mbar_init(&bempty[k], n_active);
    }
    f_mbar_init(&ifull, 1);
    for (int k = 0; k < NW * kPanDepth; ++k) f_mbar_init(&afull[k], 1);
  }
  __syncthreads();
  const char* Lbytes = reinterpret_cast<const char*>(lt.L + lt.loff[p]);
  const int nchunks = t0 / KC;
  if (warp == NW) {
    if (lane != 0) {
      // the block's diagonal tiles (last touched a launch ago) are read-modified-written at the very end: pull them into L2 now
      const int nrt = n_active * R;
      for (int a = 0; a < nrt; ++a) {
        const char* row = Lbytes + (tri(tb0 + a) + (size_t)tb0) * 512;
        for (int off = (lane - 1) * 128; off < (a + 1) * 512; off += 31 * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(row + off));
      }
    }
    if (lane == 0) {
      f_mbar_expect_tx(&ifull, kP64Tiles * 512);
      f_bulk_g2s(Ibuf, reinterpret_cast<const char*>(lt.Linv + lt.ioff[p]) + (size_t)(2 * Jp) * (kInvTilesPerBlock * 512), 2 * kInvTilesPerBlock * 512, &ifull);
#pragma unroll
      for (int a = 0; a < 4; ++a)
        f_bulk_g2s(Ibuf + (20 + 4 * a) * 512, Lbytes + (tri(t0 + 4 + a) + (size_t)t0) * 512, 4 * 512, &ifull);
      for (int kc = 0; kc < nchunks; ++kc) {
        const int buf = kc & 1;
        if (kc >= 2) f_mbar_wait(&bempty[buf], (uint32_t)(((kc >> 1) - 1) & 1));
        f_mbar_expect_tx(&bfull[buf], 8 * KC * 512);
#pragma unroll
        for (int b = 0; b < 8; ++b)
          f_bulk_g2s(Bbuf + (size_t)((buf * 8 + b) * KC) * 512, Lbytes + (tri(t0 + b) + (size_t)kc * KC) * 512, KC * 512, &bfull[buf]);
      }
    }
    return;
  }
  if (warp >= n_active) return;
  const int g = lane >> 2, l = lane & 3;
  const int tb = tb0 + warp * R;
  const int n_a = 2 * Jp + 2;                             // chunks of the own-row stream: the finished columns, then K0, K1
  unsigned char* myring = Aring + (size_t)warp * kRing;
  uint64_t* myfull = &afull[warp * kPanDepth];
  auto issue_a = [&](int a) {
    const int slot = a % kPanDepth;
    f_mbar_expect_tx(&myfull[slot], R * CW * 512);
#pragma unroll
    for (int r = 0; r < R; ++r)
      f_bulk_g2s(myring + (size_t)((slot * R + r) * CW) * 512, Lbytes + (tri(tb + r) + (size_t)a * CW) * 512, CW * 512, &myfull[slot]);
  };
  if (lane == 0)
    for (int a = 0; a < kPanDepth && a < n_a; ++a) issue_a(a);
  double acc[R][8][2];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[r][b][0] = acc[r][b][1] = 0.0;
  const double2* Bl = reinterpret_cast<const double2*>(Bbuf) + lane;
  const double2* Al = reinterpret_cast<const double2*>(myring) + lane;
  for (int a = 0; a < 2 * Jp; ++a) {
    const int slot = a % kPanDepth;
    f_mbar_wait(&myfull[slot], (uint32_t)((a / kPanDepth) & 1));
#pragma unroll
    for (int ct = 0; ct < CW; ++ct) {
      const int kc = (a * CW + ct) / KC, buf = kc & 1;
      const int pos = ct % KC;
      if (pos == 0) f_mbar_wait(&bfull[buf], (uint32_t)((kc >> 1) & 1));
      double2 af[R];
#pragma unroll
      for (int r = 0; r < R; ++r) af[r] = Al[((slot * R + r) * CW + ct) * 32];
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        double2 bf[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = Bl[((buf * 8 + 4 * hb + b) * KC + pos) * 32];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int b = 0; b < 4; ++b) dmma884(acc[r][4 * hb + b][0], acc[r][4 * hb + b][1], af[r].x, bf[b].x);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int b = 0; b < 4; ++b) dmma884(acc[r][4 * hb + b][0], acc[r][4 * hb + b][1], af[r].y, bf[b].y);
      }
      if (pos == KC - 1) {
        __syncwarp();                                     // every lane has read the B chunk
        if (lane == 0) f_mbar_arrive(&bempty[buf]);
      }
    }
    __syncwarp();                                         // every lane has read the slot
    if (lane == 0 && a + kPanDepth < n_a) issue_a(a + kPanDepth);
  }
  // ---- the two K blocks (last two chunks of the stream): C = K - sum
#pragma unroll
  for (int hb = 0; hb < 2; ++hb) {
    const int a = 2 * Jp + hb;
    const int slot = a % kPanDepth;
    f_mbar_wait(&myfull[slot], (uint32_t)((a / kPanDepth) & 1));
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double2 kt = Al[((slot * R + r) * CW + b) * 32];
        acc[r][4 * hb + b][0] = kt.x - acc[r][4 * hb + b][0];
        acc[r][4 * hb + b][1] = kt.y - acc[r][4 * hb + b][1];
      }
  }
  const int64_t xo = lt.xoff[p];
  double2* Lp = reinterpret_cast<double2*>(lt.L + lt.loff[p]);
  f_mbar_wait(&ifull, 0);
  const double2* Il = reinterpret_cast<const double2*>(Ibuf) + lane;
  const int src_lo = (lane & ~3) | (l >> 1);
  const int src_hi = (lane & ~3) | (2 + (l >> 1));
  double2 xs[R][8];                   // the final tiles, A-fragment form (= packed storage form)
  // C-fragment values of a row tile's four column tiles -> X = C inv^T with the packed inverse tiles at Il + ioff tiles;
  // out: C-fragment values f0 / f1 and the A-fragment form
  auto solve4 = [&](const double (&c0)[4], const double (&c1)[4], int itile0, double (&f0)[4], double (&f1)[4], double2 (&xo4)[4]) {
    double alo[4], ahi[4];
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      const double v0 = __shfl_sync(kFull, c0[kb], src_lo);
      const double v1 = __shfl_sync(kFull, c1[kb], src_lo);
      const double w0 = __shfl_sync(kFull, c0[kb], src_hi);
      const double w1 = __shfl_sync(kFull, c1[kb], src_hi);
      alo[kb] = (l & 1) ? v1 : v0;
      ahi[kb] = (l & 1) ? w1 : w0;
    }
    double o0[4], o1[4], p0[4], p1[4];
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) o0[cb] = o1[cb] = p0[cb] = p1[cb] = 0.0;
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
      for (int cb = kb; cb < 4; ++cb) {
        const double2 iv = Il[(itile0 + cb * (cb + 1) / 2 + kb) * 32];      // {inv[8cb+g][8kb+l], inv[8cb+g][8kb+4+l]}
        dmma884(o0[cb], o1[cb], alo[kb], iv.x);
        dmma884(p0[cb], p1[cb], ahi[kb], iv.y);
      }
    }
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      f0[cb] = o0[cb] + p0[cb];
      f1[cb] = o1[cb] + p1[cb];
      const double v0 = __shfl_sync(kFull, f0[cb], src_lo);
      const double v1 = __shfl_sync(kFull, f1[cb], src_lo);
      const double w0 = __shfl_sync(kFull, f0[cb], src_hi);
      const double w1 = __shfl_sync(kFull, f1[cb], src_hi);
      xo4[cb] = make_double2((l & 1) ? v1 : v0, (l & 1) ? w1 : w0);
    }
  };
#pragma unroll
  for (int r = 0; r < R; ++r) {
    double c0[4], c1[4], f0[4], f1[4], g0[4], g1[4];
    double2 x0[4], x1[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      c0[b] = acc[r][b][0];
      c1[b] = acc[r][b][1];
    }
    solve4(c0, c1, 0, f0, f1, x0);                       // X0 = C0 inv(L00)^T
    // C1 -= X0 L10^T: output column tile nb, k over X0's column tiles; B[k][n] = L10[8 nb + n][8 kb + k] = the packed tile (nb, kb)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const double2 lv = Il[(20 + nb * 4 + kb) * 32];
        dmma884(s0, s1, x0[kb].x, lv.x);
        dmma884(q0, q1, x0[kb].y, lv.y);
      }
      c0[nb] = acc[r][4 + nb][0] - (s0 + q0);
      c1[nb] = acc[r][4 + nb][1] - (s1 + q1);
    }
    solve4(c0, c1, 10, g0, g1, x1);                      // X1 = (C1 - X0 L10^T) inv(L11)^T
    double zpart = 0.0;
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      const double2 za = *reinterpret_cast<const double2*>(lt.alpha + xo + 8 * t0 + 8 * cb + 2 * l);
      const double2 zb = *reinterpret_cast<const double2*>(lt.alpha + xo + 8 * t0 + 32 + 8 * cb + 2 * l);
      zpart = fma(f1[cb], za.y, fma(f0[cb], za.x, zpart));
      zpart = fma(g1[cb], zb.y, fma(g0[cb], zb.x, zpart));
      xs[r][cb] = x0[cb];
      xs[r][4 + cb] = x1[cb];
      Lp[(tri(tb + r) + t0 + cb) * 32 + lane] = x0[cb];
      Lp[(tri(tb + r) + t0 + 4 + cb) * 32 + lane] = x1[cb];
    }
    // S_I += L[I, J] z_J for the rows of this tile (only this warp touches them in this launch)
    zpart += __shfl_xor_sync(kFull, zpart, 1);
    zpart += __shfl_xor_sync(kFull, zpart, 2);
    if (l == 0) lt.alpha[xo + 8 * (tb + r) + g] += zpart;
  }
  // ---- D -= X X^T on the block's diagonal tiles (64x64 lower, or 32x32 for a half block): the final tiles of all warps go to
  // shared memory (over the operand buffers: every warp is past its update loop at the first barrier)
  const int nthr = n_active * 32;
  asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
  double2* St = reinterpret_cast<double2*>(pmk_chol_smem) + lane;        // [row tile of the block][column tile of the panel][32]
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) St[((warp * R + r) * 8 + cb) * 32] = xs[r][cb];
  asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
  // Row a of the block's lower tile triangle has a + 1 tiles, all with the same left operand X[a]: a warp takes rows w and
  // (rows - 1 - w) -- nine tiles each for a full block, five for a half block -- and runs a row's tiles side by side (independent
  // accumulators), the old values of the whole row loaded first.
  const int nrt = n_active * R;                            // row tiles of the block
  auto syrk_row = [&](int a) {
    double2 dold[8];
    double o0[8], o1[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      o0[b] = o1[b] = 0.0;
      dold[b] = make_double2(0.0, 0.0);
      PMK_UNIFORM_IF(b <= a) dold[b] = __ldcg(Lp + (tri(tb0 + a) + tb0 + b) * 32 + lane);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const double2 af = St[(a * 8 + c) * 32];
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        PMK_UNIFORM_IF(b <= a) {
          const double2 bf = St[(b * 8 + c) * 32];
          dmma884(o0[b], o1[b], af.x, bf.x);
          dmma884(o0[b], o1[b], af.y, bf.y);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      PMK_UNIFORM_IF(b <= a) Lp[(tri(tb0 + a) + tb0 + b) * 32 + lane] = make_double2(dold[b].x - o0[b], dold[b].y - o1[b]);
    }
  };
  syrk_row(nrt - 1 - warp);
  syrk_row(warp);
}

// ---------------------------------------------------------------------------------------------
// M_IJ = L_IJ inv(L_JJ) for every strictly-lower 32x32 block (one warp per row tile; full occupancy, HBM-bound:
// reads L once, writes M once).  With M the pair kernel's blocked TRSM needs no diagonal solve between its
// updates:  W_J := L_JJ S_J obeys  W_I = C_I - sum_{J<I} M_IJ W_J.
__global__ void __launch_bounds__(256)
k_make_M(LeafTable lt, int first_leaf) {
  const int p = first_leaf + blockIdx.x;
  const int ntl = lt.npad[p] >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.y * 8 + warp;
  if (t >= ntl) return;
  const int g = lane >> 2, l = lane & 3;
  const double2* __restrict__ Lrow = reinterpret_cast<const double2*>(lt.L + lt.loff[p]) + tri(t) * 32 + lane;
  double* Mrow = lt.M + lt.loff[p] + tri(t) * 64;
  const double* __restrict__ Ib = lt.Linv + lt.ioff[p];
  const int nJ = t >> 2;                         // column blocks strictly left of the row tile's own block
  for (int J = 0; J < nJ; ++J) {
    const double* Iblk = Ib + (size_t)J * kInvDoublesPerBlock;
    double2 af[4];
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) af[kb] = Lrow[(4 * J + kb) * 32];
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      double o0 = 0.0, o1 = 0.0, p0 = 0.0, p1 = 0.0;
#pragma unroll
      for (int kb = cb; kb < 4; ++kb) {          // inv(L_JJ) is lower triangular: rows kb >= cb of column block cb
        // B[k][n] = Linv[8kb + k][8cb + n]; lane holds k = l (and l + 4), n = g
        const int tl = (kb * (kb + 1) / 2 + cb) * 32;
        const double b0 = Iblk[(tl + l * 4 + (g & 3)) * 2 + (g >> 2)];
        const double b1 = Iblk[(tl + (l + 4) * 4 + (g & 3)) * 2 + (g >> 2)];
        dmma884(o0, o1, af[kb].x, b0);
        dmma884(p0, p1, af[kb].y, b1);
      }
      double* tile = Mrow + (4 * J + cb) * 64;
      const int q0 = 2 * l, q1 = 2 * l + 1;
      tile[(g * 4 + (q0 & 3)) * 2 + (q0 >> 2)] = o0 + p0;
      tile[(g * 4 + (q1 & 3)) * 2 + (q1 >> 2)] = o1 + p1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// alpha = L^-T L^-1 y per leaf (the reference solves U\y by LU, mixtureGP.jl:106; same solution up to
// rounding -- see DESIGN.md "alpha").  Blocked substitution: GEMV updates stream the packed tiles,
// diagonal blocks use the stored 32x32 inverses.
__device__ __forceinline__ double linv_elem(const double* __restrict__ Iblk, int i, int k) {
  // element (i,k), i >= k, of one block's inverse (10 packed tiles)
  const int a = i >> 3, b = k >> 3;
  return Iblk[((a * (a + 1) / 2 + b) * 32 + (i & 7) * 4 + (k & 3)) * 2 + ((k & 7) >> 2)];
}

#ifndef PMK_SOLVE_MINB
#define PMK_SOLVE_MINB 6      // 40 registers, 6 resident leaves per SM: the solves are latency-bound (C3: 1.96 -> 1.70 ms; 5: 1.76)
#endif
template <int NW>
__global__ void __launch_bounds__(NW * 32, PMK_SOLVE_MINB)
k_solve_alpha(LeafTable lt, const int* __restrict__ order, const double* rhs_in, double* out, int backward_only) {
  extern __shared__ double sm[];
  const int p = order[blockIdx.x];
  if (lt.info[p] != 0) return;
  const int n = lt.n[p], npad = lt.npad[p];
  const int nblk = npad >> 5, ntl = npad >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, l = lane & 3;
  double* z = sm;                 // npad
  double* red = sm + npad;        // NW * 32
  double* rhs = red + NW * 32;    // 32
  const double2* __restrict__ Lp = reinterpret_cast<const double2*>(lt.L + lt.loff[p]);
  const double* __restrict__ Ib = lt.Linv + lt.ioff[p];
  const int64_t xo = lt.xoff[p];
  for (int i = threadIdx.x; i < npad; i += NW * 32) z[i] = rhs_in[xo + i];
  __syncthreads();
  // ---- forward: z <- L^-1 y  (backward_only: rhs_in already holds z, computed panel by panel by the factorisation)
  for (int J = backward_only ? nblk : 0; J < nblk; ++J) {
    double part[4] = {0.0, 0.0, 0.0, 0.0};
    for (int ct = warp; ct < 4 * J; ct += NW) {
      const double zlo = z[8 * ct + l], zhi = z[8 * ct + 4 + l];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double2 f = Lp[(tri(4 * J + a) + ct) * 32 + lane];
        part[a] = fma(f.x, zlo, part[a]);
        part[a] = fma(f.y, zhi, part[a]);
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      part[a] += __shfl_xor_sync(kFull, part[a], 1);
      part[a] += __shfl_xor_sync(kFull, part[a], 2);
      if (l == 0) red[warp * 32 + 8 * a + g] = part[a];
    }
    __syncthreads();
    if (warp == 0) {
      double r = z[32 * J + lane];
      for (int w = 0; w < NW; ++w) r -= red[w * 32 + lane];
      rhs[lane] = r;
      __syncwarp();
      const double* Iblk = Ib + (size_t)J * kInvDoublesPerBlock;
      double s = 0.0;
      for (int k = 0; k <= lane; ++k) s = fma(linv_elem(Iblk, lane, k), rhs[k], s);
      __syncwarp();
      z[32 * J + lane] = s;
    }
    __syncthreads();
  }
  // ---- backward: z <- L^-T z
  for (int J = nblk - 1; J >= 0; --J) {
    double plo[4] = {0.0, 0.0, 0.0, 0.0}, phi[4] = {0.0, 0.0, 0.0, 0.0};
    for (int t = 4 * J + 4 + warp; t < ntl; t += NW) {
      const double al = z[8 * t + g];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double2 f = Lp[(tri(t) + 4 * J + b) * 32 + lane];
        plo[b] = fma(f.x, al, plo[b]);
        phi[b] = fma(f.y, al, phi[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
      for (int m = 4; m <= 16; m <<= 1) {
        plo[b] += __shfl_xor_sync(kFull, plo[b], m);
        phi[b] += __shfl_xor_sync(kFull, phi[b], m);
      }
      if (g == 0) {
        red[warp * 32 + 8 * b + l] = plo[b];
        red[warp * 32 + 8 * b + 4 + l] = phi[b];
      }
    }
    __syncthreads();
    if (warp == 0) {
      double r = z[32 * J + lane];
      for (int w = 0; w < NW; ++w) r -= red[w * 32 + lane];
      rhs[lane] = r;
      __syncwarp();
      const double* Iblk = Ib + (size_t)J * kInvDoublesPerBlock;
      double s = 0.0;
      for (int k = lane; k < 32; ++k) s = fma(linv_elem(Iblk, k, lane), rhs[k], s);
      __syncwarp();
      z[32 * J + lane] = s;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < npad; i += NW * 32) out[xo + i] = i < n ? z[i] : 0.0;
}

// ---------------------------------------------------------------------------------------------
// One step of iterative refinement of alpha: r = y - (K + sigma2 I) alpha with the Gram entries re-evaluated
// (the factor only knows L L^T, whose distance to K + sigma2 I is the backward error refinement is meant to remove),
// d = L^-T L^-1 r by k_solve_alpha, alpha += d.  One warp per row; FP64 residual.
template <int D>
__global__ void __launch_bounds__(256)
k_alpha_residual(LeafTable lt, const int* __restrict__ order, KParams kp, double sigma2, double* __restrict__ r) {
  const int p = order[blockIdx.x];
  const int n = lt.n[p], npad = lt.npad[p];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t xo = lt.xoff[p];
  const double* __restrict__ xs = lt.xs + xo;
  const double* __restrict__ al = lt.alpha + xo;
  for (int i = blockIdx.y * 8 + warp; i < npad; i += gridDim.y * 8) {
    double acc = 0.0;
    if (i < n) {
      double xi[D];
#pragma unroll
      for (int d = 0; d < D; ++d) xi[d] = xs[d * lt.xstride + i];
      for (int j = lane; j < n; j += 32) {
        double xj[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xj[d] = xs[d * lt.xstride + j];
        double k = (i >= j) ? eval_kernel<D>(kp, xi, xj) : eval_kernel<D>(kp, xj, xi);   // the lower-triangle entry, mirrored (RKHS.jl:21-31)
        if (i == j) k = __dadd_rn(k, sigma2);
        acc = fma(k, al[j], acc);
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(kFull, acc, m);
    if (lane == 0) r[xo + i] = i < n ? lt.y[xo + i] - acc : 0.0;
  }
}

__global__ void k_alpha_add(double* __restrict__ alpha, const double* __restrict__ d, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) alpha[i] += d[i];
}

// (max diag(L) / min diag(L))^2 over the real rows of every leaf, maximum over the leaves: a lower bound of
// cond(K + sigma2 I), free with the factor.  out[0] is raised with an atomic max on the bit pattern (positive doubles order like
// their bit patterns).
__global__ void k_diag_range(LeafTable lt, double* __restrict__ out) {
  const int p = blockIdx.x;
  const int n = lt.n[p];
  const double* __restrict__ Lp = lt.L + lt.loff[p];
  double lo = INFINITY, hi = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = Lp[ltile_elem(i, i)];
    lo = fmin(lo, d);
    hi = fmax(hi, d);
  }
  __shared__ double slo[8], shi[8];
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(kFull, lo, m));
    hi = fmax(hi, __shfl_xor_sync(kFull, hi, m));
  }
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fmin(lo, slo[w]); hi = fmax(hi, shi[w]); }
    if (lt.info[p] == 0 && lo > 0.0) {
      const double ratio = hi / lo;
      atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(ratio * ratio));
    }
  }
}

// dense column-major n x n lower-triangular copy of leaf p's factor (pmk_get_L)
__global__ void k_unpack_L(LeafTable lt, int p, double* __restrict__ out, int which) {
  const int n = lt.n[p];
  const double* __restrict__ Lp = (which ? lt.P : lt.L) + lt.loff[p];
  const int64_t total = (int64_t)n * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx % n), c = (int)(idx / n);
    out[idx] = r >= c ? Lp[ltile_elem(r, c)] : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers (called from pmk_api.cu)
void launch_pack(int D, const LeafTable& lt, const int64_t* d_leaf_off, const double* dX, const double* dy, cudaStream_t s) {
  switch (D) {
    case 1: k_pack_leaves<1><<<lt.n_leaves, 128, 0, s>>>(lt, d_leaf_off, dX, dy); break;
    case 2: k_pack_leaves<2><<<lt.n_leaves, 128, 0, s>>>(lt, d_leaf_off, dX, dy); break;
    case 3: k_pack_leaves<3><<<lt.n_leaves, 128, 0, s>>>(lt, d_leaf_off, dX, dy); break;
    default: break;
  }
}

void launch_gram_tiles(int D, const LeafTable& lt, const int* d_order, int n_order, int max_npad, KParams kp, double sigma2,
                       cudaStream_t s) {
  if (n_order <= 0) return;
  dim3 grid(n_order, (max_npad / 8 + 7) / 8);
  const size_t dyn = (size_t)D * max_npad * sizeof(double);          // the leaf's points, one row of max_npad doubles per coordinate
  static DeviceOnce once;
  once.run([&] {
    const int cap = 3 * PMK_MAX_LEAF_POINTS * (int)sizeof(double);
    cudaFuncSetAttribute(k_gram_tiles<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    cudaFuncSetAttribute(k_gram_tiles<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    cudaFuncSetAttribute(k_gram_tiles<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
  });
  switch (D) {
    case 1: k_gram_tiles<1><<<grid, 256, dyn, s>>>(lt, d_order, kp, sigma2, max_npad); break;
    case 2: k_gram_tiles<2><<<grid, 256, dyn, s>>>(lt, d_order, kp, sigma2, max_npad); break;
    case 3: k_gram_tiles<3><<<grid, 256, dyn, s>>>(lt, d_order, kp, sigma2, max_npad); break;
    default: break;
  }
}

#ifndef PMK_CHOL_NW
#define PMK_CHOL_NW 8
#endif
template <int NW, int MINB>
static void launch_chol_shape(const LeafTable& lt, const int* d_order, int n_order, cudaStream_t s) {
  constexpr int R = 2;
  size_t dyn = (size_t)NW * kCholDepth * R * 32 * sizeof(double2);   // per-warp operand rings
  // tuning knob: PMK_CHOL_CTAS_PER_SM=1|2 pads the dynamic shared memory so that fewer leaves are resident per SM
  static int ctas_per_sm = [] { const char* e = getenv("PMK_CHOL_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  if (ctas_per_sm == 2) dyn = 100 * 1024;
  else if (ctas_per_sm == 1) dyn = 150 * 1024;
  static DeviceOnce once;           // static + dynamic shared memory exceeds the 48 KB default
  once.run([&] { cudaFuncSetAttribute(k_chol<NW, R, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024); });
  k_chol<NW, R, MINB><<<n_order, NW * 32, dyn, s>>>(lt, d_order);
}

// Two CTA shapes of the same kernel: 8 warps x 3 leaves per SM (80 registers) and 6 warps x 4 leaves per SM (64 registers).
// Measured on C3: 4096 leaves 10.25 ms (8 x 3) vs 10.07 ms (6 x 4); 512 leaves (one GPU's share at N = 8) 1.53 vs 1.84 ms -- the
// wider CTA wins when a wave is not full.  The second shape is taken from 12 leaves per SM on; PMK_CHOL_SHAPE=0|1 forces one.
void launch_chol(const LeafTable& lt, const int* d_order, int n_order, cudaStream_t s) {
  if (n_order <= 0) return;
  static int forced = [] { const char* e = getenv("PMK_CHOL_SHAPE"); return e ? atoi(e) : -1; }();
  const int n_sm = device_sm_count();
  const bool six = forced >= 0 ? forced == 1 : n_order >= 12 * n_sm;       // measured: 512 leaves 1.53 (8 x 3) vs 1.84 ms, 4096 leaves 10.25 vs 10.07
  if (six) launch_chol_shape<6, 4>(lt, d_order, n_order, s);
  else launch_chol_shape<PMK_CHOL_NW, PMK_CHOL_MINB>(lt, d_order, n_order, s);
}

void launch_make_M(const LeafTable& lt, int first_leaf, int n_leaves, int max_npad, cudaStream_t s) {
  if (n_leaves <= 0) return;
  dim3 grid(n_leaves, (max_npad / 8 + 7) / 8);
  k_make_M<<<grid, 256, 0, s>>>(lt, first_leaf);
}

// rhs / out: padded per-leaf vectors laid out like lt.y (nullptr = lt.y -> lt.alpha, the fit itself)
// backward_only: rhs already holds z = L^-1 y (the level-synchronous factorisation leaves it in lt.alpha)
void launch_solve(const LeafTable& lt, const int* d_order, int n_order, int max_npad, cudaStream_t s, const double* rhs, double* out,
                  int backward_only) {
  constexpr int NW = 8;
  if (n_order <= 0) return;
  const size_t smem = (size_t)(max_npad + NW * 32 + 32) * sizeof(double);
  k_solve_alpha<NW><<<n_order, NW * 32, smem, s>>>(lt, d_order, rhs ? rhs : lt.y, out ? out : lt.alpha, backward_only);
}

// Level-synchronous factorisation: leaves sorted by size (order), leaves_per_panel[Jp] = how many of them have columns 64 Jp ..
// (a prefix of `order`).  Returns the number of launches.  with_z: k_chol_factor64 also forms z = L^-1 y in lt.alpha.
int launch_chol_levels(const LeafTable& lt, const int* d_order, const std::vector<int>& leaves_per_panel, int max_npad, int with_z,
                         cudaStream_t s) {
  constexpr int PW = 4;
  constexpr size_t dyn = (size_t)2 * 8 * PMK_P64_KC * 512 + (size_t)PW * kPanDepth * 2 * kPanCW * 512 + kP64Tiles * 512;
  static DeviceOnce once;
  once.run([&] { cudaFuncSetAttribute(k_chol_panel64<PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn); });
  int launches = 0;
  const int max_ntl = max_npad / 8;
  const int nJ = (int)leaves_per_panel.size();
  if (nJ == 0 || leaves_per_panel[0] <= 0) return 0;
  k_chol_factor64<<<leaves_per_panel[0], 32, 0, s>>>(lt, d_order, 0, with_z);
  ++launches;
  for (int Jp = 0; Jp < nJ; ++Jp) {
    const int cnt = leaves_per_panel[Jp];
    if (cnt <= 0) break;
    const int rows_below = max_ntl - 8 * Jp - 8;
    if (rows_below > 0) {
      dim3 grid((rows_below + PW * 2 - 1) / (PW * 2), cnt);
      k_chol_panel64<PW><<<grid, (PW + 1) * 32, dyn, s>>>(lt, d_order, Jp);
      ++launches;
    }
    if (Jp + 1 < nJ && leaves_per_panel[Jp + 1] > 0) {
      k_chol_factor64<<<leaves_per_panel[Jp + 1], 32, 0, s>>>(lt, d_order, Jp + 1, with_z);
      ++launches;
    }
  }
  return launches;
}

void launch_alpha_residual(int D, const LeafTable& lt, const int* d_order, int n_order, KParams kp, double sigma2, double* r, cudaStream_t s) {
  if (n_order <= 0) return;
  dim3 grid(n_order, 8);
  switch (D) {
    case 1: k_alpha_residual<1><<<grid, 256, 0, s>>>(lt, d_order, kp, sigma2, r); break;
    case 2: k_alpha_residual<2><<<grid, 256, 0, s>>>(lt, d_order, kp, sigma2, r); break;
    case 3: k_alpha_residual<3><<<grid, 256, 0, s>>>(lt, d_order, kp, sigma2, r); break;
    default: break;
  }
}

void launch_alpha_add(const LeafTable& lt, const double* d, int64_t n, cudaStream_t s) {
  if (n <= 0) return;
  k_alpha_add<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(lt.alpha, d, n);
}

void launch_diag_range(const LeafTable& lt, double* out, cudaStream_t s) {
  if (lt.n_leaves <= 0) return;
  k_diag_range<<<lt.n_leaves, 256, 0, s>>>(lt, out);
}

void read_chol_cycles(unsigned long long* out, bool reset) {
  cudaMemcpyFromSymbol(out, g_chol_cycles, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_chol_cycles, z, sizeof z);
  }
}

void launch_unpack_L(const LeafTable& lt, int p, int n, double* d_out, cudaStream_t s, int which) {
  int64_t total = (int64_t)n * n;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  if (blocks < 1) blocks = 1;
  k_unpack_L<<<blocks, 256, 0, s>>>(lt, p, d_out, which);
}

}  // namespace pmk
