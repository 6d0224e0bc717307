// FP64 roofline denominators for B200 (sm_100a): register-resident DFMA loop,
// DMMA.8x8x4 (mma.sync.m8n8k4.f64) loop, and cuBLAS DGEMM burst + sustained.
// MEASURED_PEAKS.json holds HBM and bf16 only; every FP64 fraction in this repo
// is quoted against the numbers this tool prints (profiles/fp64_peaks_r01.json).
//
// Also self-checks the DMMA fragment layout assumed by the fit / query kernels:
//   A (8x4): lane holds A[lane>>2][lane&3]
//   B (4x8): lane holds B[lane&3][lane>>2]
//   C (8x8): lane holds C[lane>>2][2*(lane&3)+{0,1}]
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/fp64_peak.cu -lcublas -o tools/fp64_peak
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(1024) k_dmma(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

// mixed: per DMMA, some shared-memory operand loads (closer to a real kernel)
__global__ void k_layout(const double* A, const double* B, double* C) {
  int lane = threadIdx.x;
  double a = A[(lane >> 2) * 4 + (lane & 3)];       // A row-major 8x4
  double b = B[(lane & 3) * 8 + (lane >> 2)];       // B row-major 4x8
  double c0 = 0, c1 = 0;
  dmma884(c0, c1, a, b);
  C[(lane >> 2) * 8 + 2 * (lane & 3) + 0] = c0;     // C row-major 8x8
  C[(lane >> 2) * 8 + 2 * (lane & 3) + 1] = c1;
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  const char* json_path = argc > 1 ? argv[1] : nullptr;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  printf("device: %s, %d SMs, cc %d.%d\n", prop.name, nsm, prop.major, prop.minor);
  double* d_out; CK(cudaMalloc(&d_out, 1024));

  // ---- layout self-test
  {
    std::vector<double> A(32), B(32), C(64), R(64, 0.0);
    for (int i = 0; i < 32; ++i) { A[i] = 0.25 * i - 3; B[i] = 1.0 / (i + 1); }
    for (int m = 0; m < 8; ++m) for (int n = 0; n < 8; ++n) for (int k = 0; k < 4; ++k) R[m * 8 + n] += A[m * 4 + k] * B[k * 8 + n];
    double *dA, *dB, *dC;
    CK(cudaMalloc(&dA, 256)); CK(cudaMalloc(&dB, 256)); CK(cudaMalloc(&dC, 512));
    CK(cudaMemcpy(dA, A.data(), 256, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), 256, cudaMemcpyHostToDevice));
    k_layout<<<1, 32>>>(dA, dB, dC);
    CK(cudaMemcpy(C.data(), dC, 512, cudaMemcpyDeviceToHost));
    double err = 0; for (int i = 0; i < 64; ++i) err = fmax(err, fabs(C[i] - R[i]));
    printf("dmma layout self-test: max abs err %.3e (%s)\n", err, err < 1e-12 ? "OK" : "MISMATCH");
  }

  double best_dfma = 0, best_dmma = 0;
  const int iters = 20000;
  // ---- DFMA
  for (int wpc = 4; wpc <= 32; wpc *= 2) {   // warps per CTA, 2 CTAs/SM
    int threads = wpc * 32 > 1024 ? 1024 : wpc * 32;
    auto run = [&](auto kern, int ilp, const char* nm, double flop_per_thread_iter, double& best) {
      for (int cps = 1; cps <= 2; ++cps) {
        if (threads * cps > 2048) continue;
        float ms = time_ms([&] { kern<<<nsm * cps, threads>>>(d_out, iters, 1.0000001, 1e-9); }, 3);
        double tf = flop_per_thread_iter * ilp * (double)iters * threads * nsm * cps / (ms * 1e-3) / 1e12;
        printf("%s ilp=%d threads=%d cta/sm=%d : %.3f ms  %.2f TFLOP/s\n", nm, ilp, threads, cps, ms, tf);
        if (tf > best) best = tf;
      }
    };
    run(k_dfma<8>, 8, "dfma", 2.0, best_dfma);
    run(k_dfma<16>, 16, "dfma", 2.0, best_dfma);
    // DMMA: 8x8x4 MACs per warp instruction = 512 flop / 32 lanes = 16 flop per thread
    run(k_dmma<4>, 4, "dmma", 16.0, best_dmma);
    run(k_dmma<8>, 8, "dmma", 16.0, best_dmma);
    run(k_dmma<16>, 16, "dmma", 16.0, best_dmma);
  }
  printf("PEAK dfma %.2f TFLOP/s, dmma %.2f TFLOP/s\n", best_dfma, best_dmma);

  // ---- cuBLAS DGEMM burst and sustained
  double dgemm_burst = 0, dgemm_sust = 0;
  {
    cublasHandle_t h; cublasCreate(&h);
    for (int n : {4096, 8192}) {
      double *A, *B, *C;
      size_t bytes = (size_t)n * n * 8;
      CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
      CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes));
      double one = 1.0, zero = 0.0;
      float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 5);
      double tf = 2.0 * n * (double)n * n / (ms * 1e-3) / 1e12;
      printf("cublasDgemm n=%d burst: %.3f ms  %.2f TFLOP/s\n", n, ms, tf);
      if (tf > dgemm_burst) dgemm_burst = tf;
      if (n == 8192) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        int reps = (int)(3000.0 / ms) + 1;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float tot; cudaEventElapsedTime(&tot, e0, e1);
        dgemm_sust = 2.0 * n * (double)n * n * reps / (tot * 1e-3) / 1e12;
        printf("cublasDgemm n=%d sustained (%d reps, %.1f s): %.2f TFLOP/s\n", n, reps, tot * 1e-3, dgemm_sust);
      }
      cudaFree(A); cudaFree(B); cudaFree(C);
    }
    cublasDestroy(h);
  }
  // sustained DMMA/DFMA (3 s)
  double dmma_sust = 0, dfma_sust = 0;
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int reps = 0; float tot = 0;
    cudaEventRecord(e0);
    for (reps = 0; reps < 600; ++reps) k_dmma<8><<<nsm * 2, 512>>>(d_out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&tot, e0, e1);
    dmma_sust = 16.0 * 8 * (double)iters * 512 * nsm * 2 * reps / (tot * 1e-3) / 1e12;
    printf("dmma sustained (%.1f s): %.2f TFLOP/s\n", tot * 1e-3, dmma_sust);
    cudaEventRecord(e0);
    for (reps = 0; reps < 400; ++reps) k_dfma<8><<<nsm * 2, 512>>>(d_out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&tot, e0, e1);
    dfma_sust = 2.0 * 8 * (double)iters * 512 * nsm * 2 * reps / (tot * 1e-3) / 1e12;
    printf("dfma sustained (%.1f s): %.2f TFLOP/s\n", tot * 1e-3, dfma_sust);
  }
  if (json_path) {
    FILE* f = fopen(json_path, "w");
    if (f) {
      fprintf(f, "{\"gpu_name\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f, "
                 "\"dfma_tflops_sustained\": %.3f, \"dmma_tflops_sustained\": %.3f, "
                 "\"dgemm_tflops\": %.3f, \"dgemm_tflops_sustained\": %.3f, "
                 "\"how\": \"tools/fp64_peak.cu: register-resident fma / mma.sync.m8n8k4.f64 loops (best of 3, CUDA events) and cublasDgemm 4096^3, 8192^3 burst + 3 s sustained\"}\n",
              prop.name, nsm, best_dfma, best_dmma, dfma_sust, dmma_sust, dgemm_burst, dgemm_sust);
      fclose(f);
    }
  }
  return 0;
}
