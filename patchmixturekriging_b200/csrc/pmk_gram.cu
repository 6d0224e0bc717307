// K1: standalone Gram / cross-Gram kernel.  Replaces constructkernelmatrix(X, θ)
// (reference src/RKHS/RKHS.jl:4-34: lower triangle by evalkernel(X[i],X[j]), then mirrored) and
// constructkernelmatrix(X, Z, θ) (RKHS.jl:95-110).  HBM-write-bound: 8 n m bytes out, points in.
// A CTA owns a 128 x 64 tile; the tile's row/column points are staged into shared memory with
// 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier; output is written
// column-major with 16-byte stores, 1 KB contiguous per column segment.
// (The fit path does NOT use this kernel: k_chol evaluates the Gram entries straight into its
//  accumulators; this one serves the public constructkernelmatrix / U_set surface.)
#include <cstdlib>

#include "pmk_internal.cuh"

namespace pmk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int D>
__global__ void k_aos_to_soa(const double* __restrict__ X, int64_t n, int64_t stride, double* __restrict__ xs) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= stride) return;
#pragma unroll
  for (int d = 0; d < D; ++d) xs[d * stride + i] = i < n ? X[i * D + d] : 0.0;
}

static constexpr int TM = 128, TN = 64;

static constexpr int TMP = TM + 1;     // padded row length of the mirror tile: transposed reads are bank-conflict-free

// symmetric: 0 = cross-Gram; 1 = Gram, every entry evaluated by its own tile (the lower triangle as evalkernel(X[i], X[j]), the
// upper one with the arguments swapped = the mirrored value, RKHS.jl:27-31); 2 = Gram, tiles strictly above the diagonal are
// skipped and written by the tile below it instead: that tile keeps its 128 x 64 values in shared memory and stores them a
// second time transposed (512-byte column segments), so every kernel value is evaluated once -- half the FP64 work of a
// kernel that is bound by the sqrt + exp chains, not by its 8 n^2 bytes.
template <int D>
__global__ void __launch_bounds__(256, 3)
k_gram(const double* __restrict__ xr, int64_t xr_stride, int n, const double* __restrict__ xc, int64_t xc_stride, int m,
       KParams kp, double sigma2, int symmetric, int fast_exp, double* __restrict__ K) {
  __shared__ double s_exp[64];
  __shared__ __align__(16) double s_r[D][TM];
  __shared__ __align__(16) double s_c[D][TN];
  __shared__ __align__(8) uint64_t bar;
  extern __shared__ __align__(16) double s_t[];     // symmetric == 2: [TN][TMP] mirror tile
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * TM, j0 = blockIdx.y * TN;
  if (symmetric == 2 && j0 >= i0 + TM) return;       // strictly above the diagonal: the mirrored tile writes it
  const bool mirror = symmetric == 2 && i0 >= j0 + TN;   // strictly below: evaluate once, store twice
  if (tid == 0) mbar_init(&bar, 1);
  if (tid < 64) s_exp[tid] = c_exp2_64[tid];
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, (uint32_t)(D * (TM + TN) * sizeof(double)));
#pragma unroll
    for (int d = 0; d < D; ++d) {
      tma_load_1d(&s_r[d][0], xr + d * xr_stride + i0, TM * sizeof(double), &bar);
      tma_load_1d(&s_c[d][0], xc + d * xc_stride + j0, TN * sizeof(double), &bar);
    }
  }
  mbar_wait(&bar, 0);

  const int r2 = (tid & 63) * 2;
  const int i = i0 + r2;
  if (i < n) {
    double xa[D], xb[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      xa[d] = s_r[d][r2];
      xb[d] = s_r[d][r2 + 1];
    }
    const bool two = (i + 1 < n);
    const bool vec = two && ((n & 1) == 0);
    for (int jj = tid >> 6; jj < TN; jj += 4) {
      const int j = j0 + jj;
      if (j >= m) break;
      double xz[D];
#pragma unroll
      for (int d = 0; d < D; ++d) xz[d] = s_c[d][jj];
      double k0, k1 = 0.0;
      if (fast_exp) {
        // PMK_OPT_GRAM_FAST_EXP (squared exponential only): exp(-eps_sq |x - z|^2) without the reference's sqrt / re-square
        // round trip and with the table-driven exp of the fit and query kernels -- <= 2 ulp from kernel.jl:350-357, symmetric
        // by construction, a third of the FP64 operations: the kernel moves from FP64-ALU-bound towards its HBM writes
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const double da = xa[d] - xz[d], db = xb[d] - xz[d];
          sa = fma(da, da, sa);
          sb = fma(db, db, sb);
        }
        k0 = exp_neg_tab(-kp.p * sa, s_exp);
        k1 = exp_neg_tab(-kp.p * sb, s_exp);
        if (symmetric && i == j) k0 = __dadd_rn(k0, sigma2);
        if (symmetric && i + 1 == j) k1 = __dadd_rn(k1, sigma2);
      } else {
        // reference evaluates the lower triangle as evalkernel(X[i], X[j]) (i >= j) and mirrors it
        k0 = (symmetric && i < j) ? eval_kernel<D>(kp, xz, xa) : eval_kernel<D>(kp, xa, xz);
        if (symmetric && i == j) k0 = __dadd_rn(k0, sigma2);
        if (two) {
          k1 = (symmetric && i + 1 < j) ? eval_kernel<D>(kp, xz, xb) : eval_kernel<D>(kp, xb, xz);
          if (symmetric && i + 1 == j) k1 = __dadd_rn(k1, sigma2);
        }
      }
      double* dst = K + (int64_t)j * n + i;
      if (vec) {
        *reinterpret_cast<double2*>(dst) = make_double2(k0, k1);
      } else {
        dst[0] = k0;
        if (two) dst[1] = k1;
      }
      if (mirror) {
        s_t[jj * TMP + r2] = k0;
        s_t[jj * TMP + r2 + 1] = k1;
      }
    }
  }
  if (!mirror) return;
  __syncthreads();
  // K[j, i] = K[i, j]: column i of the mirrored block is 64 consecutive doubles; a warp stores one column per step
  const int warp = tid >> 5, lane = tid & 31;
  const int j = j0 + 2 * lane;
  const bool vec = (n & 1) == 0;
  for (int ii = warp; ii < TM; ii += 8) {
    const int ic = i0 + ii;
    if (ic >= n) break;
    if (j >= m) continue;
    const double v0 = s_t[(2 * lane) * TMP + ii], v1 = s_t[(2 * lane + 1) * TMP + ii];
    double* dst = K + (int64_t)ic * n + j;
    if (vec && j + 1 < m) {
      *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
    } else {
      dst[0] = v0;
      if (j + 1 < m) dst[1] = v1;
    }
  }
}

void launch_aos_to_soa(int D, const double* dX, int64_t n, int64_t stride, double* xs, cudaStream_t s) {
  const unsigned B = (unsigned)((stride + 255) / 256);
  switch (D) {
    case 1: k_aos_to_soa<1><<<B, 256, 0, s>>>(dX, n, stride, xs); break;
    case 2: k_aos_to_soa<2><<<B, 256, 0, s>>>(dX, n, stride, xs); break;
    case 3: k_aos_to_soa<3><<<B, 256, 0, s>>>(dX, n, stride, xs); break;
    default: break;
  }
}

// xr/xc must be readable up to the next multiple of TM / TN points (callers pad).
void launch_gram(int D, const double* xr, int64_t xr_stride, int n, const double* xc, int64_t xc_stride, int m, KParams kp,
                 double sigma2, int symmetric, double* dK, cudaStream_t s, int fast_exp) {
  fast_exp = (fast_exp && kp.kind == PMK_KERNEL_SQEXP) ? 1 : 0;
  if (n <= 0 || m <= 0) return;
  dim3 grid((n + TM - 1) / TM, (m + TN - 1) / TN);
  // Gram matrices of more than a few tiles: evaluate the lower tiles only and mirror them (mode 2); PMK_GRAM_NO_MIRROR=1
  // keeps every tile evaluating its own entries (the round-1 kernel) for A/B timing.
  static const bool no_mirror = [] { const char* e = getenv("PMK_GRAM_NO_MIRROR"); return e && atoi(e) != 0; }();
  const int mode = (symmetric && n == m && n > 2 * TM && !no_mirror) ? 2 : (symmetric ? 1 : 0);
  const size_t dyn = mode == 2 ? (size_t)TN * TMP * sizeof(double) : 0;
  static DeviceOnce once;
  once.run([&] {
    cudaFuncSetAttribute(k_gram<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TN * TMP * sizeof(double)));
    cudaFuncSetAttribute(k_gram<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TN * TMP * sizeof(double)));
    cudaFuncSetAttribute(k_gram<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TN * TMP * sizeof(double)));
  });
  switch (D) {
    case 1: k_gram<1><<<grid, 256, dyn, s>>>(xr, xr_stride, n, xc, xc_stride, m, kp, sigma2, mode, fast_exp, dK); break;
    case 2: k_gram<2><<<grid, 256, dyn, s>>>(xr, xr_stride, n, xc, xc_stride, m, kp, sigma2, mode, fast_exp, dK); break;
    case 3: k_gram<3><<<grid, 256, dyn, s>>>(xr, xr_stride, n, xc, xc_stride, m, kp, sigma2, mode, fast_exp, dK); break;
    default: break;
  }
}

// ---------------------------------------------------------------------------------------------
// The FP64 roofline denominator, measured where the bench runs (pmk_measure_fp64_peak): a register-resident loop of
// independent mma.sync.m8n8k4.f64 (-> DMMA.8x8x4) chains, 32 warps per SM x 8 accumulator pairs each -- nothing but the tensor
// pipe is exercised.  512 flops per warp-level DMMA.
__global__ void __launch_bounds__(1024) k_dmma_peak(double* __restrict__ out, int iters, double a, double b) {
  double c0[8], c1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    c0[i] = threadIdx.x * 1e-9;
    c1[i] = i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;      // never true: keeps the loop alive
}

// returns the flops one launch performs
double launch_dmma_peak(double* d_out, int iters, cudaStream_t s) {
  const int n_sm = device_sm_count();
  k_dmma_peak<<<n_sm, 1024, 0, s>>>(d_out, iters, 1.0000001, 0.9999999);
  return (double)n_sm * 32.0 * (double)iters * 8.0 * 512.0;
}

}  // namespace pmk
