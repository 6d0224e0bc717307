// Shared device helpers of libpmk_b200: kernel-function evaluation (bit-faithful to the
// reference's operation order wherever a comparison or a parity check depends on it),
// the FP64 tensor-core wrapper (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no f64
// kind, so this is the FP64 MMA path sm_100a exposes) and the packed-tile layout of L.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include "../../include/pmk.h"

// Per-phase cycle counters (tools/chol_phases.py, tools/query_phases.py) are compiled in only with
// -DPMK_PROFILE_CYCLES: reading the clock between phases costs ~10 % in the pair kernel.
#ifdef PMK_PROFILE_CYCLES
#define PMK_CYC(...) __VA_ARGS__
#else
#define PMK_CYC(...)
#endif

namespace pmk {

// Host side: run a launcher's one-time setup once per DEVICE.  cudaFuncSetAttribute is per device, and pmk_multi drives several
// devices from several host threads of one process, so "static bool configured" is neither enough nor safe.
struct DeviceOnce {
  std::mutex mu;
  unsigned long long done = 0;      // one bit per device ordinal
  template <class F>
  void run(F&& f) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    if (!((done >> (dev & 63)) & 1ull)) {
      f();
      done |= 1ull << (dev & 63);
    }
  }
};
inline int device_sm_count() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

struct KParams {
  int kind;     // pmk_kernel_id
  double p;     // eps_sq / a / eps
};

// ---------------------------------------------------------------------------------------
// DMMA 8x8x4: D(8x8) += A(8x4) * B(4x8), FP64.  Fragment layout (verified on B200 by
// tools/fp64_peak.cu):  A: lane holds A[lane>>2][lane&3];  B: lane holds B[lane&3][lane>>2];
// C/D: lane holds C[lane>>2][2*(lane&3)+{0,1}].
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// A warp-uniform condition that ptxas must turn into a BRANCH, never into predication.
// Measured on B200 (profiles/ncu_r01_v0_summary.md): a predicated-off DMMA still occupies the FP64
// tensor pipe for its full 16 cycles, so if-converted "skip this tile" guards cost as much as the
// work they skip.  The opaque trip count makes the guarded block a loop body, which cannot be
// if-converted.
__device__ __forceinline__ int opaque_int(int x) {
  int y;
  asm volatile("mov.b32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
#define PMK_UNIFORM_IF(cond) \
  _Pragma("unroll 1") for (int pmk_r_ = pmk::opaque_int((cond) ? 1 : 0); pmk_r_ > 0; --pmk_r_)

// cp.async (LDGSTS) helpers: 16-byte asynchronous global -> shared copies, used as register-free prefetch
// rings for the packed L tiles (each lane copies, and later reads back, exactly its own 16 bytes of a tile).
__device__ __forceinline__ void cp_async16_u32(uint32_t smem_addr, const void* gmem) {
  // .cg: cache in L2 only -- a tile is consumed once per CTA; other CTAs of the same leaf find it in L2
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------
// Packed storage of a leaf's lower-triangular factor L (n_pad x n_pad, n_pad % 32 == 0).
// 8x8 tiles, row-tile-major, lower tiles only: tile (t, c), c <= t, sits at tile index
// t(t+1)/2 + c.  Inside a tile the 64 doubles are stored "fragment-major": lane l of a warp
// owns the double2 {M[l>>2][l&3], M[l>>2][4+(l&3)]} = its A-fragment for two consecutive
// DMMA k-steps (and, because the SYRK/TRSM second operand is a transposed row block of L,
// also its B-fragment).  One LDG.128 per lane = one fully coalesced 512-byte tile.
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ size_t tri(int t) { return (size_t)t * (size_t)(t + 1) / 2; }
// index in DOUBLES of element (r, c), r >= c (or same diagonal tile), within the leaf's storage
__host__ __device__ __forceinline__ size_t ltile_elem(int r, int c) {
  int t = r >> 3, ct = c >> 3, ri = r & 7, ci = c & 7;
  return ((tri(t) + ct) * 32 + ri * 4 + (ci & 3)) * 2 + (ci >> 2);
}
// doubles needed for an n_pad x n_pad packed lower factor
__host__ __device__ __forceinline__ size_t ltile_doubles(int npad) { return tri(npad >> 3) * 64; }
// inverse diagonal blocks: per 32-row block, 10 lower tiles (a >= b) at a(a+1)/2 + b
static constexpr int kInvTilesPerBlock = 10;
static constexpr int kInvDoublesPerBlock = kInvTilesPerBlock * 64;

// ---------------------------------------------------------------------------------------
// kernel functions.  Products/sums that the oracle performs as separate IEEE operations
// are written with __dmul_rn/__dadd_rn so nvcc cannot contract them into FMAs.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double k_tau(int kind, double a, double tau) {
  switch (kind) {
    case PMK_KERNEL_SQEXP:                                   // kernel.jl:350-357
      return exp(__dmul_rn(-a, __dmul_rn(tau, tau)));
    case PMK_KERNEL_SPLINE34: {                              // kernel.jl:299-313
      double r = __dmul_rn(tau, a);
      double t = __dsub_rn(1.0, r);
      if (t < 0.0) return 0.0;
      double t2 = __dmul_rn(t, t), t4 = __dmul_rn(t2, t2), t6 = __dmul_rn(t4, t2);
      double poly = __dadd_rn(__dadd_rn(__dmul_rn(35.0, __dmul_rn(r, r)), __dmul_rn(18.0, r)), 3.0);
      return __ddiv_rn(__dmul_rn(poly, t6), 3.0);
    }
    case PMK_KERNEL_SPLINE12: {                              // kernel.jl:316-330
      double r = __dmul_rn(tau, a);
      double t = __dsub_rn(1.0, r);
      if (t < 0.0) return 0.0;
      double t3 = __dmul_rn(__dmul_rn(t, t), t);
      return __dmul_rn(__dadd_rn(__dmul_rn(3.0, r), 1.0), t3);
    }
    case PMK_KERNEL_SPLINE32: {                              // kernel.jl:333-347
      double r = __dmul_rn(tau, a);
      double t = __dsub_rn(1.0, r);
      if (t < 0.0) return 0.0;
      double t2 = __dmul_rn(t, t), t4 = __dmul_rn(t2, t2);
      return __dmul_rn(__dadd_rn(__dmul_rn(4.0, r), 1.0), t4);
    }
    case PMK_KERNEL_RQ: {                                    // kernel.jl:360-366
      double sa = __dsqrt_rn(a);
      double sd = __dsqrt_rn(__dadd_rn(a, __dmul_rn(tau, tau)));
      double num = __dmul_rn(__dmul_rn(sa, sa), sa);
      double den = __dmul_rn(__dmul_rn(sd, sd), sd);
      return __ddiv_rn(num, den);
    }
    default:
      return 0.0;
  }
}

__device__ __forceinline__ double k_bb_1d(int kind, double e, double x, double z) {
  switch (kind) {
    case PMK_KERNEL_BB10:                                    // kernel.jl:156-158
      return __dsub_rn(fmin(x, z), __dmul_rn(x, z));
    case PMK_KERNEL_BB20: {                                  // kernel.jl:218-225
      double s = __dadd_rn(__dmul_rn(x, x), __dmul_rn(z, z));
      if (z < x)
        return __dmul_rn(__dmul_rn(__dmul_rn(-1.0 / 6.0, z), __dsub_rn(1.0, x)), __dsub_rn(s, __dmul_rn(2.0, x)));
      return __dmul_rn(__dmul_rn(__dmul_rn(-1.0 / 6.0, x), __dsub_rn(1.0, z)), __dsub_rn(s, __dmul_rn(2.0, z)));
    }
    case PMK_KERNEL_BB1EPS: {                                // kernel.jl:168-174
      double den = e * sinh(e);
      double num = sinh(e * fmin(x, z)) * sinh(e * (1.0 - fmax(x, z)));
      return num / den;
    }
    case PMK_KERNEL_BB2EPS: {                                // kernel.jl:176-193
      double mn = fmin(x, z), mx = fmax(x, z), ad = fabs(x - z), s = x + z;
      double em1 = exp(2.0 * e) - 1.0;
      double mult = exp(-e * s) / (4.0 * (e * e * e) * (em1 * em1));
      double t1 = exp(2.0 * e) * (2.0 * e - e * s - 1.0);
      double t2 = exp(4.0 * e) * (e * s + 1.0);
      double t3 = exp(2.0 * e * (1.0 + s)) * (2.0 * e - e * s + 1.0);
      double t4 = exp(2.0 * e * s) * (e * s - 1.0);
      double t5 = exp(2.0 * e * (2.0 + mn)) * (-e * ad - 1.0);
      double t6 = exp(2.0 * e * mx) * (-e * ad + 1.0);
      double t7 = exp(2.0 * e * (1.0 + mn)) * (1.0 - 2.0 * e + e * ad);
      double t8 = exp(2.0 * e * (1.0 + mx)) * (1.0 + 2.0 * e - e * ad);
      return mult * (((((((t1 + t2) + t3) + t4) + t5) + t6) + t7) + t8);
    }
    default:
      return 0.0;
  }
}

__device__ __forceinline__ bool kernel_is_stationary(int kind) {
  return kind == PMK_KERNEL_SQEXP || kind == PMK_KERNEL_SPLINE34 || kind == PMK_KERNEL_SPLINE12 ||
         kind == PMK_KERNEL_SPLINE32 || kind == PMK_KERNEL_RQ;
}

// evalkernel(x, z, theta) for two D-vectors: kernel.jl:277-287 (tau = norm(x1-x2): sequential sum of
// squares, then sqrt; tau^2 re-squared from the rounded tau), kernel.jl:196-198 (tensor product).
template <int D>
__device__ __forceinline__ double stationary_tau(const double* x, const double* z) {
  const double d0 = __dsub_rn(x[0], z[0]);
  if (D == 1) return fabs(d0);
  double s = __dmul_rn(d0, d0);
#pragma unroll
  for (int d = 1; d < D; ++d) {
    const double dd = __dsub_rn(x[d], z[d]);
    s = __dadd_rn(s, __dmul_rn(dd, dd));
  }
  return __dsqrt_rn(s);
}

// every kernel except the squared exponential: out of line, so the hot kernels inline only the
// SqExp fast path (keeps code size and register pressure of the DMMA kernels down).  Points are
// passed by value so they travel in registers, not through local memory.
template <int D>
struct Pt {
  double v[D];
};

template <int D>
__device__ __noinline__ double eval_kernel_generic(int kind, double p, Pt<D> x, Pt<D> z) {
  if (kernel_is_stationary(kind)) return k_tau(kind, p, stationary_tau<D>(x.v, z.v));
  double out = k_bb_1d(kind, p, x.v[0], z.v[0]);
#pragma unroll
  for (int d = 1; d < D; ++d) out = __dmul_rn(out, k_bb_1d(kind, p, x.v[d], z.v[d]));
  return out;
}

template <int D>
__device__ __forceinline__ double eval_kernel(const KParams& kp, const double* x, const double* z) {
  if (kp.kind == PMK_KERNEL_SQEXP) {
    const double tau = stationary_tau<D>(x, z);
    return exp(__dmul_rn(-kp.p, __dmul_rn(tau, tau)));       // kernel.jl:350-357
  }
  Pt<D> xx, zz;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    xx.v[d] = x[d];
    zz.v[d] = z[d];
  }
  return eval_kernel_generic<D>(kp.kind, kp.p, xx, zz);
}

// N independent evaluations, written stage by stage (all distances, then all square roots, then all
// exponentials) so that the FP64 dependency chains of different evaluations interleave: the hot kernels
// evaluate dozens of independent entries per thread, and one evaluation alone is a ~50-instruction chain.
template <int D, int N>
__device__ __forceinline__ void eval_kernel_batch(const KParams& kp, const double (&xa)[N][D], const double (&xb)[N][D],
                                                  const bool (&valid)[N], double (&out)[N]) {
  if (kp.kind == PMK_KERNEL_SQEXP) {
    double arg[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double tau = stationary_tau<D>(xa[i], xb[i]);
      arg[i] = __dmul_rn(-kp.p, __dmul_rn(tau, tau));
    }
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = valid[i] ? exp(arg[i]) : 0.0;
    return;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = valid[i] ? eval_kernel<D>(kp, xa[i], xb[i]) : 0.0;
}

// 2^(j/64), j = 0..63, correctly rounded (copied to shared memory at kernel start)
static __constant__ double c_exp2_64[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0, 0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0,
    0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0, 0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0, 0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0,
    0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0, 0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0, 0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0,
    0x1.6247eb03a5585p+0, 0x1.6623882552225p+0, 0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0, 0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0,
    0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0, 0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0, 0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0,
    0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0, 0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

// exp(t) for t <= 0, table-driven: t = (64 k + j) ln2/64 + r, |r| <= ln2/128, exp(t) = 2^k 2^(j/64) (1 + p(r)) with a
// degree-5 p.  The integer 64k + j falls out of the low word of t*64/ln2 + 1.5*2^52 (no rint, no float->int conversion,
// both of which run at a quarter of the FP64 rate); 10 FP64 operations in all against 19 for exp_neg.  Max error 1.3 ulp
// against the correctly rounded value on [-699, 0] (tests/test_host.py restates the arithmetic); arguments below -699
// return exp(-699) ~ 1e-304.
__device__ __forceinline__ double exp_neg_tab(double t, const double* __restrict__ tab) {
  const double C = 0x1.71547652b82fep+6, HI = 0x1.62e42fee00000p-7, LO = 0x1.a39ef35793c76p-39, MAGIC = 0x1.8p52;
  t = fmax(t, -699.0);
  double kd = fma(t, C, MAGIC);
  const int ki = __double2loint(kd);
  kd -= MAGIC;
  double r = fma(kd, -HI, t);
  r = fma(kd, -LO, r);
  const double T = tab[ki & 63];
  const double r2 = r * r;
  const double q1 = fma(1.0 / 6, r, 0.5), q2 = fma(1.0 / 120, r, 1.0 / 24);
  const double pr = fma(fma(q2, r2, q1), r2, r);
  const double res = fma(T, pr, T);
  return __hiloint2double(__double2hiint(res) + ((ki >> 6) << 20), __double2loint(res));
}

// The same function for the pair kernel's phase E, which is bound by its instruction count, not by the FP64 pipe: the table comes
// as a 32-bit shared-memory address (no generic-to-shared conversion in the loop), and the clamp at -699 is ONE unsigned integer
// minimum on the high word of t (for t <= -0.0 the high words order like the magnitudes; 0xC085D800 is the high word of -699.0)
// instead of a double-precision maximum (a DSETP on the FP64 pipe and three selects).  Same bits as exp_neg_tab for -699 <= t <= 0.
__device__ __forceinline__ double exp_neg_tab_s(double t, uint32_t tab_u32) {
  const double C = 0x1.71547652b82fep+6, HI = 0x1.62e42fee00000p-7, LO = 0x1.a39ef35793c76p-39, MAGIC = 0x1.8p52;
  t = __hiloint2double((int)min((unsigned)__double2hiint(t), 0xC085D800u), __double2loint(t));
  double kd = fma(t, C, MAGIC);
  const int ki = __double2loint(kd);
  kd -= MAGIC;
  double r = fma(kd, -HI, t);
  r = fma(kd, -LO, r);
  double T;
  asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(tab_u32 + (uint32_t)((ki & 63) << 3)));
  const double r2 = r * r;
  const double q1 = fma(1.0 / 6, r, 0.5), q2 = fma(1.0 / 120, r, 1.0 / 24);
  const double pr = fma(fma(q2, r2, q1), r2, r);
  const double res = fma(T, pr, T);
  return __hiloint2double(__double2hiint(res) + ((ki >> 6) << 20), __double2loint(res));
}

// R row tiles per unit, CW column tiles per chunk, DEPTH chunks in flight per warp.
// Dynamic shared memory: [kRW][DEPTH][R][CW][512 B] rings | K fragments: npmax * MQ doubles | per-unit ||S||^2:
// [2][ucap][MQ] | (STAGE_X) inputs + alpha: [D+1][npmax]
// sequential dot product, no contraction (contract of partition.jl:69,254,285 and mixtureGP.jl:361)
template <int D>
__device__ __forceinline__ double dot_seq(const double* v, const double* x) {
  double s = __dmul_rn(v[0], x[0]);
#pragma unroll
  for (int d = 1; d < D; ++d) s = __dadd_rn(s, __dmul_rn(v[d], x[d]));
  return s;
}

}  // namespace pmk
