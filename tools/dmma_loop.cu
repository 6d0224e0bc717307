// How close to 1 DMMA / 16 cycles / sub-partition can the K3 inner loop get?  Variants of the operand feed
// (shared-memory fragment loads, software pipelining, DMMA ordering) with 3 or 4 warps per sub-partition.
// Prints cycles per DMMA per sub-partition (16.0 = pipe peak).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/dmma_loop.cu -o tools/dmma_loop
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ int opaque_int(int x) {
  int y;
  asm volatile("mov.b32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
#define UNIFORM_IF(cond) _Pragma("unroll 1") for (int r_ = opaque_int((cond) ? 1 : 0); r_ > 0; --r_)

// MODE 0: bf (8 LDS.64) per ct, per tile: LDS.128 af then 8 DMMA (4 independent, then the 4 dependent k-step-2 ones)
// MODE 1: same, af of the next tile loaded before the current tile's DMMAs (software pipelined)
// MODE 2: no shared-memory loads at all (register operands), same DMMA order
// MODE 3: as 0, tiles guarded by real branches (the opaque-loop trick of the product kernel)
// MODE 4: as 1 with guards
// MODE 5: two tiles interleaved: 8 independent DMMAs, then their 8 dependent ones
template <int MODE, int NT, int NQT>
__global__ void __launch_bounds__(512, 1) k_loop(double* out, long long* cyc, int iters, unsigned active) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, l = lane & 3;
  constexpr int LDQ = 8 * NQT + 4;
  double* Wb = smem;                                   // 32 x LDQ
  double2* ring = reinterpret_cast<double2*>(smem + 32 * LDQ) + (size_t)warp * (NT * 32) + lane;
  for (int k = tid; k < 32 * LDQ + 16 * NT * 64; k += blockDim.x) smem[k] = 1e-3 * (k % 17);
  __syncthreads();
  double acc[NT][NQT][2];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int nt = 0; nt < NQT; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) {
      double bf[2][NQT];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < NQT; ++nt)
          bf[ks][nt] = (MODE == 2) ? 1e-3 * (ks + nt + ct) : Wb[(8 * ct + 4 * ks + l) * LDQ + nt * 8 + g];
      if (MODE == 5) {
#pragma unroll
        for (int i = 0; i < NT; i += 2) {
          const double2 a0 = ring[i * 32], a1 = ring[(i + 1) * 32];
#pragma unroll
          for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], a0.x, bf[0][nt]);
#pragma unroll
          for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i + 1][nt][0], acc[i + 1][nt][1], a1.x, bf[0][nt]);
#pragma unroll
          for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], a0.y, bf[1][nt]);
#pragma unroll
          for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i + 1][nt][0], acc[i + 1][nt][1], a1.y, bf[1][nt]);
        }
      } else if (MODE == 1 || MODE == 4) {
        double2 af = ring[0];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const double2 cur = af;
          if (i + 1 < NT) af = ring[(i + 1) * 32];
          if (MODE == 4) {
            UNIFORM_IF(active & (1u << i)) {
#pragma unroll
              for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], cur.x, bf[0][nt]);
#pragma unroll
              for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], cur.y, bf[1][nt]);
            }
          } else {
#pragma unroll
            for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], cur.x, bf[0][nt]);
#pragma unroll
            for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], cur.y, bf[1][nt]);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          if (MODE == 3) {
            UNIFORM_IF(active & (1u << i)) {
              const double2 af = ring[i * 32];
#pragma unroll
              for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.x, bf[0][nt]);
#pragma unroll
              for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.y, bf[1][nt]);
            }
          } else {
            const double2 af = (MODE == 2) ? make_double2(1e-3 * i, 2e-3) : ring[i * 32];
#pragma unroll
            for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.x, bf[0][nt]);
#pragma unroll
            for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.y, bf[1][nt]);
          }
        }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int nt = 0; nt < NQT; ++nt) s += acc[i][nt][0] + acc[i][nt][1];
  if (s == 123.456) out[0] = s;
}

template <int MODE, int NT, int NQT>
static void run(int warps, const char* tag, double* d_out, long long* d_cyc, int nsm) {
  const int iters = 2000;
  const size_t smem = (32 * (8 * NQT + 4) + 16 * NT * 64) * sizeof(double);
  CK(cudaFuncSetAttribute(k_loop<MODE, NT, NQT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long h = 0;
  for (int rep = 0; rep < 2; ++rep) {
    k_loop<MODE, NT, NQT><<<nsm, warps * 32, smem>>>(d_out, d_cyc, iters, 0xffffffffu);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(&h, d_cyc, sizeof h, cudaMemcpyDeviceToHost));
  const double dmma_per_sp = (double)iters * 4 * NT * NQT * 2 * (warps / 4.0);
  printf("%-46s warps=%2d NT=%d NQT=%d : %.2f cyc/DMMA/subpart (%.1f %% of peak)\n", tag, warps, NT, NQT, h / dmma_per_sp,
         1600.0 / (h / dmma_per_sp));
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  double* d_out; CK(cudaMalloc(&d_out, 1024));
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 64));
  run<2, 4, 4>(16, "regs only", d_out, d_cyc, nsm);
  run<0, 4, 4>(16, "LDS bf + LDS.128 af per tile", d_out, d_cyc, nsm);
  run<1, 4, 4>(16, "  + af software-pipelined", d_out, d_cyc, nsm);
  run<3, 4, 4>(16, "LDS, guarded tiles (opaque loop)", d_out, d_cyc, nsm);
  run<4, 4, 4>(16, "  + af software-pipelined, guarded", d_out, d_cyc, nsm);
  run<5, 4, 4>(16, "two tiles interleaved", d_out, d_cyc, nsm);
  run<2, 6, 4>(12, "regs only", d_out, d_cyc, nsm);
  run<0, 6, 4>(12, "LDS bf + LDS.128 af per tile", d_out, d_cyc, nsm);
  run<1, 6, 4>(12, "  + af software-pipelined", d_out, d_cyc, nsm);
  run<3, 6, 4>(12, "LDS, guarded tiles (opaque loop)", d_out, d_cyc, nsm);
  run<5, 6, 4>(12, "two tiles interleaved", d_out, d_cyc, nsm);
  run<0, 1, 4>(16, "one tile per ct (late J)", d_out, d_cyc, nsm);
  run<3, 1, 4>(16, "one tile per ct, guarded", d_out, d_cyc, nsm);
  run<0, 2, 4>(16, "two tiles per ct", d_out, d_cyc, nsm);
  run<3, 2, 4>(16, "two tiles per ct, guarded", d_out, d_cyc, nsm);
  run<0, 6, 3>(16, "class 1 shape", d_out, d_cyc, nsm);
  run<3, 6, 3>(16, "class 1 shape, guarded", d_out, d_cyc, nsm);
  run<0, 8, 2>(16, "class 2 shape", d_out, d_cyc, nsm);
  return 0;
}
