// The control structure of the pair kernel without its data movement: triangular active set, one barrier per 32-column
// step, B fragments reloaded per column tile, guarded row tiles -- operands from static shared memory, no TMA, no
// kernel evaluations.  Prints the DMMA pipe utilisation per tile (n_pad = 512, 32 queries), to separate what the
// STRUCTURE costs from what data movement / evaluations cost.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/trmm_sim.cu -o tools/trmm_sim
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ int opaque_int(int x) { int y; asm volatile("mov.b32 %0, %1;" : "=r"(y) : "r"(x)); return y; }
#define UNIFORM_IF(cond) _Pragma("unroll 1") for (int r_ = opaque_int((cond) ? 1 : 0); r_ > 0; --r_)

// MODE bit0: barrier per step; bit1: diagonal block handled (t >= 4J+ct) vs (t >= 4J+4); bit2: STEP64 (two 32-blocks per barrier)
template <int NW, int NT, int NQT, int MODE>
__global__ void __launch_bounds__(NW * 32, 1) k_sim(double* out, long long* cyc, int tiles, int ntl) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, l = lane & 3;
  constexpr int LDQ = 8 * NQT + 4;
  double* Wb = smem;
  double2* ring = reinterpret_cast<double2*>(smem + 3 * 32 * LDQ) + (size_t)warp * (NT * 4 * 32) + lane;
  for (int k = tid; k < 3 * 32 * LDQ + NW * NT * 4 * 64; k += blockDim.x) smem[k] = 1e-3 * (k % 17);
  __syncthreads();
  const int nblk = ntl >> 2;
  const long long t0 = clock64();
  double s = 0;
  for (int tile = 0; tile < tiles; ++tile) {
    double acc[NT][NQT][2];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int nt = 0; nt < NQT; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;
    for (int J = 0; J < nblk; ++J) {
      if ((MODE & 1) && (!(MODE & 4) || !(J & 1))) __syncthreads();
      const double* Kj = Wb + (J % 3) * (32 * LDQ);
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) {
        double bf[2][NQT];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int nt = 0; nt < NQT; ++nt) bf[ks][nt] = Kj[(8 * ct + 4 * ks + l) * LDQ + nt * 8 + g];
        const int thr = (MODE & 2) ? 4 * J + ct : 4 * J + 4;
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          UNIFORM_IF(warp + NW * i >= thr && warp + NW * i < ntl) {
            const double2 af = ring[(i * 4 + ct) * 32];
#pragma unroll
            for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.x, bf[0][nt]);
#pragma unroll
            for (int nt = 0; nt < NQT; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], af.y, bf[1][nt]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int nt = 0; nt < NQT; ++nt) s += acc[i][nt][0] + acc[i][nt][1];
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
  if (s == 123.456) out[0] = s;
}

template <int NW, int NT, int NQT, int MODE>
static void run(const char* tag, double* d_out, long long* d_cyc, int nsm, int ntl) {
  const int tiles = 20;
  const size_t smem = (3 * 32 * (8 * NQT + 4) + NW * NT * 4 * 64) * sizeof(double);
  CK(cudaFuncSetAttribute(k_sim<NW, NT, NQT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long h = 0;
  for (int rep = 0; rep < 2; ++rep) { k_sim<NW, NT, NQT, MODE><<<nsm, NW * 32, smem>>>(d_out, d_cyc, tiles, ntl); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(&h, d_cyc, sizeof h, cudaMemcpyDeviceToHost));
  // DMMAs per tile: tiles (t, c) with c <= t (MODE&2) or strictly-lower 32-blocks
  long long tt = 0;
  for (int t = 0; t < ntl; ++t) for (int c = 0; c < ntl; ++c) { const int J = c >> 2, ct = c & 3; const int thr = (MODE & 2) ? 4 * J + ct : 4 * J + 4; if (t >= thr) ++tt; }
  const double ideal = (double)tt * 2 * NQT * 16 / 4;
  printf("%-40s NW=%2d NT=%d NQT=%d ntl=%d : %.0f cycles/tile, ideal %.0f -> %.1f %% of DMMA peak\n", tag, NW, NT, NQT, ntl, (double)h / tiles, ideal,
         100.0 * ideal / ((double)h / tiles));
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  double* d_out; CK(cudaMalloc(&d_out, 1024));
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 64));
  run<12, 6, 4, 3>("12 warps, barrier/step, diag", d_out, d_cyc, nsm, 64);
  run<12, 6, 4, 2>("12 warps, no barrier, diag", d_out, d_cyc, nsm, 64);
  run<12, 6, 4, 7>("12 warps, barrier/2 steps, diag", d_out, d_cyc, nsm, 64);
  run<16, 4, 4, 3>("16 warps, barrier/step, diag", d_out, d_cyc, nsm, 64);
  run<16, 4, 4, 2>("16 warps, no barrier, diag", d_out, d_cyc, nsm, 64);
  run<16, 4, 4, 1>("16 warps, barrier/step, TRSM shape", d_out, d_cyc, nsm, 64);
  run<12, 6, 4, 3>("12 warps, barrier/step, diag, n_pad 448", d_out, d_cyc, nsm, 56);
  run<12, 8, 3, 3>("class 1 shape, barrier/step", d_out, d_cyc, nsm, 80);
  run<12, 8, 3, 2>("class 1 shape, no barrier", d_out, d_cyc, nsm, 80);
  return 0;
}
