// Can dedicated "evaluation" warps compute the squared-exponential cross-covariance (exp(-a |x-z|^2), FP64 ALU work)
// next to 12 warps that saturate the FP64 tensor pipe?  Measures, per SM: DMMA cycles per instruction per sub-partition
// (16 = peak) and kernel evaluations per 1000 cycles, for NE evaluation warps with ILP independent evaluations in flight.
// The query kernel needs ~250 evaluations per 1000 cycles per SM (16384 per 66.5k-cycle tile).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/dmma_eval_mix.cu -o tools/dmma_eval_mix
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// exp(t) for t <= 0, Estrin evaluation of the degree-11 polynomial (short dependency chain), no special cases beyond underflow
__device__ __forceinline__ double exp_neg_estrin(double t) {
  const double L2E = 1.4426950408889634, LN2H = 6.93147180369123816490e-01, LN2L = 1.90821492927058770002e-10;
  t = fmax(t, -700.0);
  const double nf = rint(t * L2E);
  double r = fma(nf, -LN2H, t);
  r = fma(nf, -LN2L, r);
  // 1 + r + r^2/2 + ... + r^11/11!
  const double c2 = 0.5, c3 = 1.0 / 6, c4 = 1.0 / 24, c5 = 1.0 / 120, c6 = 1.0 / 720, c7 = 1.0 / 5040, c8 = 1.0 / 40320,
               c9 = 1.0 / 362880, c10 = 1.0 / 3628800, c11 = 1.0 / 39916800;
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double p01 = 1.0 + r, p23 = fma(c3, r, c2), p45 = fma(c5, r, c4), p67 = fma(c7, r, c6), p89 = fma(c9, r, c8),
               pab = fma(c11, r, c10);
  const double q0 = fma(p23, r2, p01), q1 = fma(p67, r2, p45), q2 = fma(pab, r2, p89);
  const double s = fma(q2, r8, fma(q1, r4, q0));
  const int n = (int)nf;
  return __hiloint2double(__double2hiint(s) + (n << 20), __double2loint(s));
}

template <int ILP, int EXPKIND>
__global__ void __launch_bounds__(512, 1) k_mix(double* out, long long* cyc, int iters_mma, int n_eval_warps, int evals_per_lane,
                                                const double* __restrict__ pts) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, l = lane & 3;
  constexpr int NT = 6, NQT = 4, LDQ = 36;
  double* Wb = smem;
  double2* ring = reinterpret_cast<double2*>(smem + 32 * LDQ) + (size_t)(warp % 12) * (NT * 32) + lane;
  double* Kout = smem + 32 * LDQ + 12 * NT * 64;      // 32 x LDQ
  for (int k = tid; k < 32 * LDQ + 12 * NT * 64 + 32 * LDQ; k += blockDim.x) smem[k] = 1e-3 * (k % 17);
  __syncthreads();
  const long long t0 = clock64();
  double s = 0;
  if (warp < 12) {
    double acc[NT][NQT][2];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int nt = 0; nt < NQT; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;
    for (int it = 0; it < iters_mma; ++it) {
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) {
        double bf[2][NQT];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int nt = 0; nt < NQT; ++nt) bf[ks][nt] = Wb[(8 * ct + 4 * ks + l) * LDQ + nt * 8 + g];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const double2 af = ring[i * 32];
#pragma unroll
          for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.x, bf[0][nt]);
#pragma unroll
          for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.y, bf[1][nt]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int nt = 0; nt < NQT; ++nt) s += acc[i][nt][0] + acc[i][nt][1];
    if (lane == 0) cyc[warp] = clock64() - t0;
  } else if (warp < 12 + n_eval_warps) {
    // lane = query; rows broadcast from shared memory (staged training points)
    const double xq0 = pts[lane], xq1 = pts[32 + lane];
    double u = 0.0;
    for (int e = 0; e < evals_per_lane; e += ILP) {
      double kv[ILP];
#pragma unroll
      for (int k = 0; k < ILP; ++k) {
        const int row = (e + k) & 511;
        const double d0 = xq0 - smem[row], d1 = xq1 - smem[512 + row];
        const double s2 = fma(d1, d1, d0 * d0);
        kv[k] = EXPKIND ? exp_neg_estrin(-408.0 * s2) : exp(-408.0 * s2);
      }
#pragma unroll
      for (int k = 0; k < ILP; ++k) {
        Kout[((e + k) & 31) * LDQ + lane] = kv[k];
        u = fma(kv[k], smem[1024 + ((e + k) & 511)], u);
      }
    }
    s = u;
    if (lane == 0) cyc[warp] = clock64() - t0;
  }
  if (s == 123.456) out[0] = s;
}

template <int ILP, int EXPKIND>
static void run(int n_eval, double* d_out, long long* d_cyc, const double* d_pts, int nsm, bool mma) {
  const int iters = mma ? 700 : 0;                         // ~ one tile's worth x 10
  const int evals_per_lane = n_eval ? 16384 * 10 / 32 / n_eval : 0;   // 10 tiles' worth of evaluations shared by the eval warps
  const size_t smem = (32 * 36 + 12 * 6 * 64 + 32 * 36) * sizeof(double);
  CK(cudaFuncSetAttribute(k_mix<ILP, EXPKIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long h[16];
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaMemset(d_cyc, 0, sizeof h));
    k_mix<ILP, EXPKIND><<<nsm, 512, smem>>>(d_out, d_cyc, iters, n_eval, evals_per_lane, d_pts);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost));
  long long mm = 0, me = 0;
  for (int w = 0; w < 12; ++w) mm = h[w] > mm ? h[w] : mm;
  for (int w = 12; w < 12 + n_eval; ++w) me = h[w] > me ? h[w] : me;
  const double dmma_per_sp = (double)iters * 4 * 6 * 4 * 2 * 3;
  printf("eval warps=%d ilp=%2d exp=%s mma=%d : ", n_eval, ILP, EXPKIND ? "estrin" : "cuda  ", (int)mma);
  if (mma) printf("dmma %.2f cyc/DMMA/subpart (%.1f%%)  ", mm / dmma_per_sp, 1600.0 / (mm / dmma_per_sp));
  if (n_eval) printf("evals: %.0f per 1000 cycles per SM (need ~250)", 16384.0 * 10 / me * 1000.0);
  printf("\n");
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  double* d_out; CK(cudaMalloc(&d_out, 1024));
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 128));
  double hp[64]; for (int i = 0; i < 64; ++i) hp[i] = 0.01 * i;
  double* d_pts; CK(cudaMalloc(&d_pts, sizeof hp)); CK(cudaMemcpy(d_pts, hp, sizeof hp, cudaMemcpyHostToDevice));
  run<8, 0>(0, d_out, d_cyc, d_pts, nsm, true);
  run<4, 0>(4, d_out, d_cyc, d_pts, nsm, false);
  run<8, 0>(4, d_out, d_cyc, d_pts, nsm, false);
  run<8, 1>(4, d_out, d_cyc, d_pts, nsm, false);
  run<4, 0>(4, d_out, d_cyc, d_pts, nsm, true);
  run<8, 0>(4, d_out, d_cyc, d_pts, nsm, true);
  run<4, 1>(4, d_out, d_cyc, d_pts, nsm, true);
  run<8, 1>(4, d_out, d_cyc, d_pts, nsm, true);
  run<12, 1>(4, d_out, d_cyc, d_pts, nsm, true);
  run<8, 0>(3, d_out, d_cyc, d_pts, nsm, true);
  run<8, 1>(3, d_out, d_cyc, d_pts, nsm, true);
  run<12, 1>(3, d_out, d_cyc, d_pts, nsm, true);
  run<16, 1>(3, d_out, d_cyc, d_pts, nsm, true);
  return 0;
}
