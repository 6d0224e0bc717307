// Do DMMA.8x8x4 and scalar DFMA share an execution pipe on B200 (sm_100a)?  And what is the dependent-issue latency
// of a DMMA?  Both answers bound what the fused query kernel (K3) can reach: it has to evaluate exp() kernels (DFMA
// chains) next to its DMMA triangular solve.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/fp64_mix.cu -o tools/fp64_mix
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// warps [0, n_dmma_warps) run DMMA loops (ILP accumulators), the rest run DFMA loops (8 chains).
// clock64 per role is written out so each role's own rate can be computed.
template <int ILP>
__global__ void __launch_bounds__(1024) k_mix(double* out, long long* cyc, int iters_dmma, int iters_dfma, int n_dmma_warps,
                                              double a, double b) {
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  double s = 0;
  if (warp < n_dmma_warps) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters_dmma; ++it) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  } else {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters_dfma; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
  if (s == 123.456) out[0] = s;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  double* d_out; CK(cudaMalloc(&d_out, 1024));
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 64 * sizeof(long long)));
  long long h[64];
  printf("device %s, %d SMs\n", prop.name, nsm);
  auto run = [&](auto kern, int ilp, int warps, int n_dmma, int it_dmma, int it_dfma, const char* tag) {
    CK(cudaMemset(d_cyc, 0, 64 * sizeof(long long)));
    kern<<<nsm, warps * 32>>>(d_out, d_cyc, it_dmma, it_dfma, n_dmma, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    kern<<<nsm, warps * 32>>>(d_out, d_cyc, it_dmma, it_dfma, n_dmma, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost));
    // per sub-partition: warps w with w%4 == sp.  cycles per DMMA on a sub-partition = cyc / (dmma warps on it * ilp * iters)
    double cd = 0, cf = 0;
    int nd = 0, nf = 0;
    for (int w = 0; w < warps; ++w) {
      if (w < n_dmma) { cd += (double)h[w]; ++nd; } else { cf += (double)h[w]; ++nf; }
    }
    double dmma_per_sp = nd ? (double)nd / 4.0 * ilp * it_dmma : 0;   // DMMAs issued per sub-partition
    double dfma_per_sp = nf ? (double)nf / 4.0 * 8 * it_dfma : 0;     // warp-DFMAs issued per sub-partition
    printf("%-28s warps=%2d dmma_warps=%2d ilp=%d : ", tag, warps, n_dmma, ilp);
    if (nd) printf("dmma %.0f cyc -> %.2f cyc/DMMA/subpart  ", cd / nd, (cd / nd) / dmma_per_sp);
    if (nf) printf("dfma %.0f cyc -> %.2f cyc/warpDFMA/subpart", cf / nf, (cf / nf) / dfma_per_sp);
    printf("\n");
  };
  // 1. latency: one warp per sub-partition, ILP 1, 2, 4
  run(k_mix<1>, 1, 4, 4, 20000, 0, "dmma latency ilp1");
  run(k_mix<2>, 2, 4, 4, 20000, 0, "dmma ilp2");
  run(k_mix<4>, 4, 4, 4, 20000, 0, "dmma ilp4");
  run(k_mix<8>, 8, 4, 4, 20000, 0, "dmma ilp8");
  run(k_mix<1>, 1, 16, 16, 20000, 0, "dmma 4 warps/sp ilp1");
  run(k_mix<2>, 2, 16, 16, 20000, 0, "dmma 4 warps/sp ilp2");
  // 2. alone
  run(k_mix<4>, 4, 8, 8, 20000, 0, "dmma alone 2w/sp");
  run(k_mix<4>, 4, 8, 0, 0, 20000, "dfma alone 2w/sp");
  run(k_mix<4>, 4, 16, 0, 0, 20000, "dfma alone 4w/sp");
  // 3. together: 8 DMMA warps + 8 DFMA warps, sized to take the same time alone (DMMA 16 cyc x 4 x it ; DFMA 2 cyc x 8 x it)
  run(k_mix<4>, 4, 16, 8, 20000, 80000, "mixed 2w dmma + 2w dfma /sp");
  run(k_mix<4>, 4, 16, 8, 20000, 20000, "mixed, dfma light (1/4)");
  run(k_mix<4>, 4, 16, 12, 20000, 60000, "mixed 3w dmma + 1w dfma /sp");
  return 0;
}
