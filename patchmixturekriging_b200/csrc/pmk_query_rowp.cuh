// K3 (default solver): fused per-(query, leaf) kernel, explicit inverse P = inv(L), ROW-PANEL product.
//
//   queryinner! (reference src/RKHS/mixtureGP.jl:296-316):  kq = k(x*, X_p);  u = dot(kq, c);  v = k(x*,x*) - ||L \ kq||^2
//
// s = P kq is a triangular matrix product and only ||s||^2 is wanted, so a row tile of s can be formed completely,
// squared, added up and thrown away.  That turns the tile of work (leaf p, MQ = 8*NQT pairs binned to p) into two phases
// with no dependency between warps inside either:
//
//   E  all 16 compute warps evaluate the cross-covariance K = k(X_p, x*) (n x MQ) ONCE into shared memory, already in DMMA
//      B-fragment order (one 16-byte load per lane = the fragments of two k-steps); dot(kq, c) is reduced on the way.
//      k(x*, X_p) never touches HBM.  The FP64 ALU does not run next to a saturated DMMA pipe on B200
//      (tools/dmma_eval_mix.cu), so evaluating in a phase of its own costs nothing over interleaving it.
//   M  the row tiles of P are handed out in units of R consecutive row tiles, largest first (unit m costs ~R(Rm + (R+1)/2)
//      tile products): the first 16 statically, the rest claimed from a shared-memory counter as a warp's operand stream
//      reaches the end of its unit, so the warps finish within one small unit of each other.  A warp streams ITS rows of P --
//      contiguous in the packed row-tile-major layout -- through its own ring of 1-D TMA bulk copies (cp.async.bulk ->
//      UBLKCP, mbarrier completion, issued by its own lane 0, DEPTH chunks ahead), accumulates S_unit = sum_c P[unit, c] K_c
//      in 2*R*NQT registers and writes the unit's ||S_unit||^2 (one value per query) to shared memory.  No barrier, no shared
//      accumulator, no other warp's data: the inner loop is  wait(full) - LDS.128 - DMMA ...  and nothing else.
//
// Two 512-thread barriers per tile (K complete / K free).  Warp 16 stages tiles two ahead (pair ids, query points, leaf
// descriptor; the leaf's inputs and weights by TMA) and finalises finished tiles (per-unit partials summed in unit order,
// whichever warp produced them -> bit-reproducible; k(x*,x*) - ||s||^2, clamp, scatter to the pair arrays).
#pragma once
#include <type_traits>
#include "pmk_query_trsm.cuh"

namespace pmk {

#ifndef PMK_ROWP_WARPS
#define PMK_ROWP_WARPS 16
#endif
static constexpr int kRW = PMK_ROWP_WARPS;      // compute warps (4 per SM sub-partition)
static constexpr int kRowpCompute = kRW * 32;
static constexpr int kRowpThreads = kRowpCompute + 32;
static constexpr int kRSB = 2;                  // staged tiles: the one being worked on and the next

__device__ __forceinline__ void rowp_bar() { asm volatile("bar.sync 3, %0;" ::"n"(kRowpCompute) : "memory"); }

struct RowpDesc {
  int p, n, ntl, n_units;
};

#ifndef PMK_ROWP_SWP
#define PMK_ROWP_SWP 0
#endif
#ifndef PMK_ROWP_FASTISSUE
#define PMK_ROWP_FASTISSUE 1
#endif
#ifdef PMK_ROWP_MAXNREG
#define PMK_ROWP_BOUNDS __maxnreg__(PMK_ROWP_MAXNREG)
#else
#define PMK_ROWP_BOUNDS __launch_bounds__(kRowpThreads, 1)
#endif
template <int D, int NQT, int R, int CW, int DEPTH, bool STAGE_X>
__global__ void PMK_ROWP_BOUNDS
k_query_rowp(LeafTable lt, PairWork w, QueryPlan q, KParams kp, int flags, int npmax, double* __restrict__ pair_u,
             double* __restrict__ pair_v) {
  constexpr int MQ = 8 * NQT;
  constexpr int SLOT_BYTES = R * CW * 512;
  constexpr int RING_BYTES = kRW * DEPTH * SLOT_BYTES;
  constexpr int EW = (kRW / NQT) * NQT;       // evaluating warps: warp w < EW owns query tile w % NQT
  constexpr int ESTRIDE = kRW / NQT;
  __shared__ double s_xq[kRSB][D * MQ];
  __shared__ int64_t s_pair[kRSB][MQ];
  __shared__ RowpDesc s_desc[kRSB];
  __shared__ int s_next[kRSB];                // next unit (counted from the largest) nobody has claimed yet
  __shared__ double ured[2][kRW * 8];
  __shared__ double s_exp[64];
  __shared__ __align__(8) uint64_t full_bar[kRW * DEPTH];
  __shared__ __align__(8) uint64_t stage_full[kRSB], stage_empty[kRSB], tile_done[2], fin_done[2], x_full, x_empty;

  extern __shared__ __align__(128) unsigned char pmk_dyn_smem[];
  const int ucap = (npmax / 8 + R - 1) / R;                                          // units of the largest leaf
  unsigned char* Kf = pmk_dyn_smem + RING_BYTES;                                    // [column tile][nt][lane] double2
  double* vpart = reinterpret_cast<double*>(Kf + (size_t)npmax * MQ * 8);           // [ucap][MQ]
  double* s_X = vpart + (size_t)ucap * MQ;                                          // [D+1][npmax]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = w.tile_off[w.n_class_leaves];
  const int64_t my_tiles = n_tiles > (int64_t)blockIdx.x ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  if (tid == 0) {
    for (int k = 0; k < kRW * DEPTH; ++k) q_mbar_init(q_smem_u32(&full_bar[k]), 1);
    for (int k = 0; k < kRSB; ++k) {
      q_mbar_init(q_smem_u32(&stage_full[k]), 32);          // every lane of the staging warp
      q_mbar_init(q_smem_u32(&stage_empty[k]), kRW);        // lane 0 of every compute warp
    }
    for (int k = 0; k < 2; ++k) {
      q_mbar_init(q_smem_u32(&tile_done[k]), kRW);
      q_mbar_init(q_smem_u32(&fin_done[k]), 32);
    }
    q_mbar_init(q_smem_u32(&x_full), 1);
    q_mbar_init(q_smem_u32(&x_empty), kRW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 64) s_exp[tid] = c_exp2_64[tid];
  __syncthreads();

  if (warp == kRW) {
    // ============================ staging warp: stage tiles ahead, finalise finished tiles =====================
    int lo = 0;
    auto stage = [&](int64_t kt) {      // tile kt of this CTA into buffer kt % kRSB
      const int64_t tile = blockIdx.x + kt * (int64_t)gridDim.x;
      while (lo + 1 < w.n_class_leaves && w.tile_off[lo + 1] <= tile) ++lo;       // tiles come in increasing order
      const int p = w.class_leaves[lo];
      const int sb = (int)(kt % kRSB);
      const int64_t gleaf = w.leaf_base + p;
      const int64_t pstart = w.leaf_pair_start[gleaf] + (tile - w.tile_off[lo]) * MQ;
      const int64_t pend = w.leaf_pair_start[gleaf + 1];
      const int cnt = (int)((pend - pstart) < (int64_t)MQ ? (pend - pstart) : (int64_t)MQ);
      if (lane < MQ) {
        const int qi = lane < cnt ? lane : cnt - 1;
        const int64_t gp = w.sorted_pair[pstart + qi];
        s_pair[sb][lane] = lane < cnt ? gp : (int64_t)-1;
        const int64_t j = q.pair_q[gp];
#pragma unroll
        for (int d = 0; d < D; ++d) s_xq[sb][d * MQ + lane] = q.Xq[j * D + d];
      }
      if (lane == 0) {
        RowpDesc dsc;
        dsc.p = p;
        dsc.n = lt.n[p];
        dsc.ntl = (dsc.n + 7) >> 3;       // row / column tiles that hold real points (the identity padding contributes zeros)
        dsc.n_units = (dsc.ntl + R - 1) / R;
        s_desc[sb] = dsc;
        s_next[sb] = kRW;                 // units 0 .. kRW-1 (from the largest) go to warps 0 .. kRW-1
      }
      q_mbar_arrive(q_smem_u32(&stage_full[sb]));
    };
    int lo_x = 0;
    auto load_x = [&](int64_t kt) {     // the leaf's inputs and weights of tile kt into the (single) input buffer
      if (!STAGE_X) return;
      if (lane == 0) {
        const int64_t tile = blockIdx.x + kt * (int64_t)gridDim.x;
        while (lo_x + 1 < w.n_class_leaves && w.tile_off[lo_x + 1] <= tile) ++lo_x;
        const int p = w.class_leaves[lo_x];
        const uint32_t bytes = (uint32_t)lt.npad[p] * 8u;
        const uint32_t fb = q_smem_u32(&x_full);
        q_mbar_expect_tx(fb, (D + 1) * bytes);
        const uint32_t dst = q_smem_u32(s_X);
#pragma unroll
        for (int d = 0; d < D; ++d) q_bulk_g2s(dst + d * (npmax * 8), lt.xs + d * lt.xstride + lt.xoff[p], bytes, fb);
        q_bulk_g2s(dst + D * (npmax * 8), lt.alpha + lt.xoff[p], bytes, fb);
      }
    };
    if (my_tiles > 0) {
      stage(0);
      load_x(0);
    }
    if (my_tiles > 1) stage(1);
    for (int64_t kf = 0; kf < my_tiles; ++kf) {
      const int sb = (int)(kf % kRSB), vb = (int)(kf & 1);
      if (STAGE_X && kf + 1 < my_tiles) {
        q_mbar_wait(q_smem_u32(&x_empty), (uint32_t)(kf & 1));       // phase E of tile kf is over: the input buffer is free
        load_x(kf + 1);
      }
      // ---- finalise tile kf, then reuse its staging buffer for tile kf + kRSB
      q_mbar_wait(q_smem_u32(&tile_done[vb]), (uint32_t)((kf >> 1) & 1));
      if (lane < MQ) {
        const int n_units = s_desc[sb].n_units;
        const double* vp = vpart + lane;
        double vs = 0.0, u = 0.0;
        for (int m = 0; m < n_units; ++m) vs += vp[m * MQ];        // unit order: independent of which warp took which unit
#pragma unroll
        for (int ww = 0; ww < EW; ++ww)
          if (ww % NQT == (lane >> 3)) u += ured[vb][ww * 8 + (lane & 7)];
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
      if (kf + kRSB < my_tiles) {
        q_mbar_wait(q_smem_u32(&stage_empty[sb]), (uint32_t)((kf / kRSB) & 1));   // the compute warps are done with it
        stage(kf + kRSB);
      }
    }
    return;
  }

  // ================================== compute warps ==============================================================
  const int g = lane >> 2, l = lane & 3;
  const uint32_t ring_u32 = q_smem_u32(pmk_dyn_smem) + (uint32_t)(warp * (DEPTH * SLOT_BYTES));
  const double2* ring_g = reinterpret_cast<const double2*>(pmk_dyn_smem + (size_t)warp * (DEPTH * SLOT_BYTES)) + lane;
  const double2* Kfl = reinterpret_cast<const double2*>(Kf) + lane;
  const uint32_t my_full = q_smem_u32(&full_bar[warp * DEPTH]);
  uint32_t phbits = 0;                        // parity of the next completion of each ring slot
  const int e_nt = warp % NQT, e_c0 = warp / NQT;

  for (int64_t kt = 0; kt < my_tiles; ++kt) {
    const int sb = (int)(kt % kRSB), vb = (int)(kt & 1);
    PMK_CYC(long long c_t0 = clock64(), c_full = 0, c_a;)
    q_mbar_wait(q_smem_u32(&stage_full[sb]), (uint32_t)((kt / kRSB) & 1));
    const int p = s_desc[sb].p, n = s_desc[sb].n, ntl = s_desc[sb].ntl, n_units = s_desc[sb].n_units;
    const char* Pp = reinterpret_cast<const char*>(lt.P + lt.loff[p]);
    double* vp = vpart;

    // chunk ch of unit m: column tiles [ch CW, ch CW + CW) of the unit's R row tiles (row tile t has columns 0..t)
    auto unit_chunks = [&](int m) {
      const int last = (R * m + R - 1 < ntl ? R * m + R - 1 : ntl - 1);      // last row tile of the unit = its widest
      return (last + CW) / CW;                                               // ceil((last + 1) / CW)
    };
    auto issue = [&](int m, int ch, int slot) {     // lane 0 only
      const int c0 = ch * CW;
      const uint32_t fb = my_full + slot * 8;
      uint32_t total = 0;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int t = R * m + i;
        int nc = t + 1 - c0;
        nc = nc > CW ? CW : nc;
        if (t < ntl && nc > 0) total += (uint32_t)nc * 512u;
      }
      q_mbar_expect_tx(fb, (PMK_K3_X & 4) ? 0u : total);
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int t = R * m + i;
        int nc = t + 1 - c0;
        nc = nc > CW ? CW : nc;
        if (t < ntl && nc > 0 && !(PMK_K3_X & 4))
          q_bulk_g2s(ring_u32 + (uint32_t)(slot * SLOT_BYTES + i * (CW * 512)), Pp + (tri(t) + (size_t)c0) * 512,
                     (uint32_t)nc * 512u, fb);
      }
    };

    // The operand stream of this warp: units claimed largest first (index i counts from the largest: unit n_units-1-i),
    // the first one statically, the following ones from the tile's counter when the stream reaches the end of a unit.
    // Claimed-but-not-yet-consumed units wait in a byte queue (at most DEPTH + 1 of them).
    // A unit's leading chunks are "full" (every row tile meets every column tile): R copies of CW tiles from a running
    // pointer; only its last chunks (the diagonal end) take the general path.
    auto full_chunks = [&](int m) { return R * m + R - 1 < ntl ? (R * m + 1) / CW : 0; };
    uint64_t uq = 0;
    int uqn = 0;
    int im = n_units - 1 - warp, ich = 0, inch = 0, ifull = 0;
    const char* isrc = Pp;
    auto begin_unit = [&]() {
      inch = unit_chunks(im);
      ifull = full_chunks(im);
      isrc = Pp + tri(R * im) * 512;
      uq |= (uint64_t)im << (8 * uqn);
      ++uqn;
    };
    if (im >= 0) begin_unit();
    auto issue_next = [&](int slot) {
      if (im >= 0) {
        if (PMK_ROWP_FASTISSUE && ich < ifull) {
          if (lane == 0) {
            const uint32_t fb = my_full + slot * 8;
            q_mbar_expect_tx(fb, (PMK_K3_X & 4) ? 0u : (uint32_t)SLOT_BYTES);
#pragma unroll
            for (int i = 0; i < R; ++i)
              if (!(PMK_K3_X & 4)) q_bulk_g2s(ring_u32 + (uint32_t)(slot * SLOT_BYTES + i * (CW * 512)),
                         isrc + (uint32_t)((i * R * im + i * (i + 1) / 2) * 512), CW * 512, fb);
          }
        } else if (lane == 0) {
          issue(im, ich, slot);
        }
        isrc += CW * 512;
        if (++ich == inch) {
          int idx = 0;
          if (lane == 0) idx = atomicAdd(&s_next[sb], 1);
          idx = __shfl_sync(kFullQ, idx, 0);
          im = n_units - 1 - idx;
          ich = 0;
          if (im >= 0) begin_unit();
        }
      }
    };
    // prefill the ring with the first chunks (they land during phase E)
#pragma unroll
    for (int s = 0; s < DEPTH; ++s) issue_next(s);

    if (kt >= 2) q_mbar_wait(q_smem_u32(&fin_done[vb]), (uint32_t)(((kt >> 1) - 1) & 1));   // ured[vb] is free again
    if (STAGE_X) q_mbar_wait(q_smem_u32(&x_full), (uint32_t)(kt & 1));
    PMK_CYC(const long long c_e0 = clock64();)

    // ---- phase E: K = k(X_p, x*) into shared memory in B-fragment order, mean partials on the way ------------------
    // lane (g, l) of an evaluating warp owns query 8 e_nt + g and, per column tile c, rows 8c + l and 8c + 4 + l: exactly
    // the double2 it will later load as the B fragments of the tile's two k-steps.  Two column tiles = four evaluations per
    // step, branch-free (rows past n are evaluated on row n-1 and zeroed) and written stage by stage so that the four
    // FP64 dependency chains interleave.
    if (warp < EW) {
      double xq[D];
#pragma unroll
      for (int d = 0; d < D; ++d) xq[d] = s_xq[sb][d * MQ + 8 * e_nt + g];
      const double* __restrict__ xs = lt.xs + lt.xoff[p];
      const double* __restrict__ al = lt.alpha + lt.xoff[p];
      const bool sqexp = kp.kind == PMK_KERNEL_SQEXP;
      const uint32_t s_exp_u32 = q_smem_u32(s_exp);
      double usum = 0.0;
      // one step = column tiles c and c + ESTRIDE.  FULL: both hold real rows only -- no row clamps, no selects.
      auto eval_step = [&](int c, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        double xr[4][D], av[4], kv[4];
        bool ok[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = 8 * (c + (e >> 1) * ESTRIDE) + 4 * (e & 1) + l;
          ok[e] = FULL || row < n;
          const int rc = ok[e] ? row : n - 1;
#pragma unroll
          for (int d = 0; d < D; ++d) xr[e][d] = STAGE_X ? s_X[d * npmax + rc] : xs[d * lt.xstride + rc];
          av[e] = STAGE_X ? s_X[D * npmax + rc] : al[rc];
        }
        if (sqexp) {
          // exp(-a |x - z|^2) without the reference's sqrt / re-square round trip (kernel.jl:350-357): <= 2 ulp from it,
          // far inside the 1e-9 contract of the posterior
          double arg[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            double s2 = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
              const double dd = xq[d] - xr[e][d];
              s2 = fma(dd, dd, s2);
            }
            arg[e] = -kp.p * s2;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) kv[e] = (PMK_K3_X & 1) ? arg[e] : exp_neg_tab_s(arg[e], s_exp_u32);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) kv[e] = eval_kernel<D>(kp, xq, xr[e]);      // evalkernel(xq, X[i]), mixtureGP.jl:304
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!FULL) kv[e] = ok[e] ? kv[e] : 0.0;
          usum = fma(kv[e], av[e], usum);                                          // dot(kq, c)    mixtureGP.jl:308
        }
        reinterpret_cast<double2*>(Kf)[(c * NQT + e_nt) * 32 + lane] = make_double2(kv[0], kv[1]);
        if (FULL || c + ESTRIDE < ntl) reinterpret_cast<double2*>(Kf)[((c + ESTRIDE) * NQT + e_nt) * 32 + lane] = make_double2(kv[2], kv[3]);
      };
      const int nfull = n >> 3;           // column tiles without padding rows
      int c = e_c0;
      for (; c + ESTRIDE < nfull; c += 2 * ESTRIDE) eval_step(c, std::true_type{});
      for (; c < ntl; c += 2 * ESTRIDE) eval_step(c, std::false_type{});
      usum += __shfl_xor_sync(kFullQ, usum, 1);
      usum += __shfl_xor_sync(kFullQ, usum, 2);
      if (l == 0) ured[vb][warp * 8 + g] = usum;
    }
    if (STAGE_X) {
      __syncwarp();
      if (lane == 0) q_mbar_arrive(q_smem_u32(&x_empty));
    }
    PMK_CYC(const long long c_e1 = clock64();)
    rowp_bar();                       // K is complete
    // the unit partials are single-buffered: the previous tile must have been finalised (long ago: a whole phase E)
    if (kt >= 1) q_mbar_wait(q_smem_u32(&fin_done[vb ^ 1]), (uint32_t)(((kt - 1) >> 1) & 1));
    PMK_CYC(const long long c_m0 = clock64();)

    // ---- phase M: this warp's units of S = P K, each squared and summed per query ----------------------------------
    int slot = 0;
    auto load_b = [&](double2 (&bf)[NQT], int c) {
#pragma unroll
      for (int nt = 0; nt < NQT; ++nt) bf[nt] = Kfl[(c * NQT + nt) * 32];
    };
    auto mma_tile = [&](double (&a)[NQT][2], const double2& af, const double2 (&bf)[NQT]) {
#pragma unroll
      for (int nt = 0; nt < NQT; ++nt) dmma884(a[nt][0], a[nt][1], af.x, bf[nt].x);
#pragma unroll
      for (int nt = 0; nt < NQT; ++nt) dmma884(a[nt][0], a[nt][1], af.y, bf[nt].y);
    };
    while (uqn > 0) {
      const int m = (int)(uq & 255u);
      uq >>= 8;
      --uqn;
      const int nch = unit_chunks(m), nfull = full_chunks(m);
      double acc[R][NQT][2];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int nt = 0; nt < NQT; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;
      // ---- full chunks: every row tile of the unit meets every column tile of the chunk.  Software-pipelined: the K
      // fragments (B) of the next column tile are loaded before the current tile's DMMAs -- across the chunk boundary too,
      // they do not depend on the ring -- and the P fragments (A) of the next column tile of the chunk likewise.
      double2 bf[2][NQT];
      if (PMK_ROWP_SWP && nfull > 0) load_b(bf[0], 0);
      for (int ch = 0; ch < nfull; ++ch) {
        PMK_CYC(c_a = clock64();)
        q_mbar_wait(my_full + slot * 8, (phbits >> slot) & 1u);
        PMK_CYC(c_full += clock64() - c_a;)
        phbits ^= 1u << slot;
        const double2* rs = ring_g + slot * (SLOT_BYTES / 16);
        const int c0 = ch * CW;
        double2 af[2][R];
#pragma unroll
        for (int i = 0; i < R; ++i) af[0][i] = rs[(i * CW) * 32];
#pragma unroll
        for (int cc = 0; cc < CW; ++cc) {
          // column c0 + cc + 1 exists in K: it is at most R m + 1 <= the unit's last row tile
          if (PMK_ROWP_SWP) load_b(bf[(cc + 1) & 1], c0 + cc + 1);
          else load_b(bf[cc & 1], c0 + cc);
          if (cc + 1 < CW) {
#pragma unroll
            for (int i = 0; i < R; ++i) af[(cc + 1) & 1][i] = rs[(i * CW + cc + 1) * 32];
          }
#pragma unroll
          for (int i = 0; i < R; ++i) mma_tile(acc[i], af[cc & 1][i], bf[cc & 1]);
        }
        if (PMK_ROWP_SWP && (CW & 1)) {       // odd chunk width: the prefetched fragments sit in the other buffer
#pragma unroll
          for (int nt = 0; nt < NQT; ++nt) bf[0][nt] = bf[1][nt];
        }
        __syncwarp();                 // every lane has read the slot: refill it with the chunk DEPTH ahead
        issue_next(slot);
        if (++slot == DEPTH) slot = 0;
      }
      // ---- the unit's diagonal end (P is lower triangular: row tile t meets column tile c only if c <= t) and short
      // units at the end of the leaf
      for (int ch = nfull; ch < nch; ++ch) {
        PMK_CYC(c_a = clock64();)
        q_mbar_wait(my_full + slot * 8, (phbits >> slot) & 1u);
        PMK_CYC(c_full += clock64() - c_a;)
        phbits ^= 1u << slot;
        const double2* rs = ring_g + slot * (SLOT_BYTES / 16);
        const int c0 = ch * CW;
        if (R * m + R - 1 < ntl) {
          // whole unit: the leading columns of the chunk (c <= R m) meet every row tile, then the triangle
          // c = R m + 1 + j meets the row tiles i > j
          const int ncf = R * m + 1 - c0;           // full columns in this chunk (may be <= 0 or > CW)
#pragma unroll
          for (int cc = 0; cc < CW; ++cc) {
            PMK_UNIFORM_IF(cc < ncf) {
              double2 b[NQT];
              load_b(b, c0 + cc);
#pragma unroll
              for (int i = 0; i < R; ++i) mma_tile(acc[i], rs[(i * CW + cc) * 32], b);
            }
          }
#pragma unroll
          for (int j = 0; j < R - 1; ++j) {
            const int cc = R * m + 1 + j - c0;      // position of the triangle's column j in this chunk
            PMK_UNIFORM_IF(cc >= 0 && cc < CW) {
              double2 b[NQT];
              load_b(b, c0 + cc);
#pragma unroll
              for (int i = j + 1; i < R; ++i) mma_tile(acc[i], rs[(i * CW + cc) * 32], b);
            }
          }
        } else {
          // short unit at the end of the leaf: guarded tile by tile, with real branches
#pragma unroll
          for (int cc = 0; cc < CW; ++cc) {
            const int c = c0 + cc;
            PMK_UNIFORM_IF(c <= R * m + R - 1 && c < ntl) {      // c <= the unit's last row tile
              double2 b[NQT];
              load_b(b, c);
#pragma unroll
              for (int i = 0; i < R; ++i) {
                PMK_UNIFORM_IF(R * m + i < ntl && c <= R * m + i) mma_tile(acc[i], rs[(i * CW + cc) * 32], b);
              }
            }
          }
        }
        __syncwarp();
        issue_next(slot);
        if (++slot == DEPTH) slot = 0;
      }
      // ||S_unit||^2 per query: squares over the unit's rows (this lane: row g of each tile), then over g
#pragma unroll
      for (int nt = 0; nt < NQT; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double v = 0.0;
#pragma unroll
          for (int i = 0; i < R; ++i) v = fma(acc[i][nt][e], acc[i][nt][e], v);
          v += __shfl_xor_sync(kFullQ, v, 4);
          v += __shfl_xor_sync(kFullQ, v, 8);
          v += __shfl_xor_sync(kFullQ, v, 16);
          if (g == 0) vp[m * MQ + nt * 8 + 2 * l + e] = v;
        }
      }
    }
    PMK_CYC(const long long c_m1 = clock64();)

    // ---- hand the tile to the staging warp ---------------------------------------------------------------------------
    __syncwarp();
    if (lane == 0) {
      q_mbar_arrive(q_smem_u32(&tile_done[vb]));
      q_mbar_arrive(q_smem_u32(&stage_empty[sb]));
    }
    PMK_CYC(const long long c_b0 = clock64();)
    rowp_bar();                       // every warp has left phase M: K may be overwritten
#ifdef PMK_PROFILE_CYCLES
    if (lane == 0) {      // summed over all compute warps: total, tile start, phase E, wait K complete, phase M, #, of M: operand waits, wait K free
      const long long c_now = clock64();
      atomicAdd(&g_query_cycles[0], (unsigned long long)(c_now - c_t0));
      atomicAdd(&g_query_cycles[1], (unsigned long long)(c_e0 - c_t0));
      atomicAdd(&g_query_cycles[2], (unsigned long long)(c_e1 - c_e0));
      atomicAdd(&g_query_cycles[3], (unsigned long long)(c_m0 - c_e1));
      atomicAdd(&g_query_cycles[4], (unsigned long long)(c_m1 - c_m0));
      atomicAdd(&g_query_cycles[5], 1ull);
      atomicAdd(&g_query_cycles[6], (unsigned long long)c_full);
      atomicAdd(&g_query_cycles[7], (unsigned long long)(c_now - c_b0));
    }
#endif
  }
}

// Launches the kernel with this ring shape if its shared memory fits (the leaf's inputs staged when they fit too);
// returns false if even the unstaged form does not fit.
template <int D, int NQT, int R, int CW, int DEPTH>
static bool launch_rowp_one(const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp, int flags, int npmax,
                            double* pu, double* pv, cudaStream_t s) {
  constexpr size_t ring = (size_t)kRW * DEPTH * R * CW * 512;
  const size_t kbytes = (size_t)npmax * 8 * NQT * 8;
  const size_t vbytes = (size_t)((npmax / 8 + R - 1) / R) * 8 * NQT * 8;
  const size_t xbytes = (size_t)(D + 1) * npmax * 8;
  auto k_staged = k_query_rowp<D, NQT, R, CW, DEPTH, true>;
  auto k_plain = k_query_rowp<D, NQT, R, CW, DEPTH, false>;
  const int n_sm = device_sm_count();
  cudaFuncAttributes fa{};
  cudaFuncGetAttributes(&fa, k_staged);
  const size_t kMaxDyn = (size_t)232448 - fa.sharedSizeBytes - 64;   // 227 KB per CTA minus the kernel's static part
  static DeviceOnce once;
  once.run([&] {
    cudaFuncSetAttribute(k_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDyn);
    cudaFuncSetAttribute(k_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDyn);
  });
  // persistent: tiles are strided over the CTAs
  if (ring + kbytes + vbytes + xbytes <= kMaxDyn)
    k_staged<<<n_sm, kRowpThreads, ring + kbytes + vbytes + xbytes, s>>>(lt, w, q, kp, flags, npmax, pu, pv);
  else if (ring + kbytes + vbytes <= kMaxDyn)
    k_plain<<<n_sm, kRowpThreads, ring + kbytes + vbytes, s>>>(lt, w, q, kp, flags, npmax, pu, pv);
  else
    return false;
  return true;
}

#ifndef PMK_ROWP_CW
#define PMK_ROWP_CW 4
#endif
#ifndef PMK_ROWP_DEPTH
#define PMK_ROWP_DEPTH 1
#endif
#ifndef PMK_ROWP_CW3          // chunk width of the classes with MQ = 8 (n_pad > 1024)
#define PMK_ROWP_CW3 3
#endif

// npmax: largest n_pad among the class's leaves (sizes the K buffer); same size classes / MQ as the other pair kernels.
// Measured on c3_mini (16 warps): chunks of 2 column tiles x 2 slots 10.71 ms, 3 x 1 10.43, 4 x 1 10.47, 1 x 4 11.51;
// 12 warps 4 x 1 10.62, 5 x 1 10.53; 8 warps 8 x 1 11.24 -- the per-chunk cost (barrier wait, refill) outweighs a second
// slot: the other three warps of the sub-partition hide the refill latency.  The K buffer (n_pad x MQ doubles) has
// priority over the rings: a class whose largest leaf leaves no room for 4-tile chunks runs with 2-tile chunks.
template <int D>
bool launch_rowp_d(int cls, const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp, int flags, int npmax,
                   double* pu, double* pv, cudaStream_t s) {
  if (cls == 0)
    return launch_rowp_one<D, 4, 2, PMK_ROWP_CW, PMK_ROWP_DEPTH>(lt, w, q, kp, flags, npmax, pu, pv, s) ||
           launch_rowp_one<D, 4, 2, 2, 1>(lt, w, q, kp, flags, npmax, pu, pv, s);
  if (cls == 1)
    return launch_rowp_one<D, 3, 2, PMK_ROWP_CW, PMK_ROWP_DEPTH>(lt, w, q, kp, flags, npmax, pu, pv, s) ||
           launch_rowp_one<D, 3, 2, 2, 1>(lt, w, q, kp, flags, npmax, pu, pv, s);
  // large classes, 4 row tiles per unit.  C4 (2 M queries, class <= 1024 / class <= 1536, ms): 1-tile chunks x 2 slots 64.9 / 150.1,
  // 2 x 1 57.2 / 119.3, 1 x 1 75.9 / 188.4, 3 x 1 (MQ = 8 classes only: K leaves the room) - / 102.5
  if (cls == 2)
    return launch_rowp_one<D, 2, 4, 2, 1>(lt, w, q, kp, flags, npmax, pu, pv, s) ||
           launch_rowp_one<D, 2, 4, 1, 2>(lt, w, q, kp, flags, npmax, pu, pv, s);
  return launch_rowp_one<D, 1, 4, PMK_ROWP_CW3, 1>(lt, w, q, kp, flags, npmax, pu, pv, s) ||
         launch_rowp_one<D, 1, 4, 2, 1>(lt, w, q, kp, flags, npmax, pu, pv, s) ||
         launch_rowp_one<D, 1, 4, 1, 2>(lt, w, q, kp, flags, npmax, pu, pv, s);
}

}  // namespace pmk
