// Micro-benchmark of the serial step of the batched Cholesky: Cholesky + inverse of one 32x32 block by one warp
// (factor_block32 of csrc/pmk_fit.cu), alone on an SM sub-partition and with several such warps per SM.
// Variants: U = fully unrolled left-looking (round 1), R = rolled left-looking with four partial sums,
//           G = right-looking with the block's rows in registers (lane = row), inverse by a second register sweep.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/factor_bench.cu -o tools/factor_bench
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
static constexpr unsigned kFull = 0xffffffffu;
static constexpr int LD = 36, LDD = 33;

template <bool UNROLL>
__device__ __noinline__ int factor_ll(double* Dbuf, double* Ibuf, int lane, long long* t_loop, long long* t_tail) {
  const double* rowp = Dbuf + lane * LDD;
  const double* xcol = Ibuf + lane;
  int info = 0;
  long long tl = 0, tt = 0;
#pragma unroll(UNROLL ? 32 : 1)
  for (int j = 0; j < 32; ++j) {
    long long c0 = clock64();
    const double* lj = Dbuf + j * LDD;
    double s0 = rowp[j], s1 = 0.0, s2 = 0.0, s3 = 0.0;
    double t0 = (j == lane) ? 1.0 : 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    int k = 0;
#pragma unroll(UNROLL ? 8 : 1)
    for (; k + 4 <= j; k += 4) {
      const double l0 = lj[k], l1 = lj[k + 1], l2 = lj[k + 2], l3 = lj[k + 3];
      s0 = fma(-rowp[k], l0, s0); s1 = fma(-rowp[k + 1], l1, s1); s2 = fma(-rowp[k + 2], l2, s2); s3 = fma(-rowp[k + 3], l3, s3);
      t0 = fma(-l0, xcol[k * LD], t0); t1 = fma(-l1, xcol[(k + 1) * LD], t1); t2 = fma(-l2, xcol[(k + 2) * LD], t2); t3 = fma(-l3, xcol[(k + 3) * LD], t3);
    }
#pragma unroll(UNROLL ? 3 : 1)
    for (; k < j; ++k) {
      const double l0 = lj[k];
      s0 = fma(-rowp[k], l0, s0);
      t0 = fma(-l0, xcol[k * LD], t0);
    }
    const double s = (s0 + s1) + (s2 + s3);
    long long c1 = clock64();
    const double d = __shfl_sync(kFull, s, j);
    if (!(d > 0.0) && info == 0) info = j + 1;
    const double inv = rsqrt(d);
    __syncwarp();
    Dbuf[lane * LDD + j] = lane > j ? s * inv : (lane == j ? d * inv : 0.0);
    Ibuf[j * LD + lane] = (j >= lane) ? ((t0 + t1) + (t2 + t3)) * inv : 0.0;
    __syncwarp();
    long long c2 = clock64();
    tl += c1 - c0; tt += c2 - c1;
  }
  *t_loop = tl; *t_tail = tt;
  return info;
}

// right-looking, rows in registers: lane i owns row i (a[0..31]).  Step j: pivot from lane j, column scaled, rank-1 update
// of the trailing columns with the column broadcast by shuffles.  Then X = inv(L): lane c owns column c of X in registers.
__device__ __noinline__ int factor_rl(double* Dbuf, double* Ibuf, int lane, long long* t_loop, long long* t_tail) {
  double a[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) a[k] = Dbuf[lane * LDD + k];
  int info = 0;
  long long c0 = clock64();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double d = __shfl_sync(kFull, a[j], j);
    if (!(d > 0.0) && info == 0) info = j + 1;
    const double inv = rsqrt(d);
    const double lij = lane > j ? a[j] * inv : (lane == j ? d * inv : 0.0);
    a[j] = lij;
#pragma unroll
    for (int k = j + 1; k < 32; ++k) {
      const double lkj = __shfl_sync(kFull, lij, k);
      a[k] = fma(-lij, lkj, a[k]);
    }
  }
  long long c1 = clock64();
#pragma unroll
  for (int k = 0; k < 32; ++k) Dbuf[lane * LDD + k] = k <= lane ? a[k] : 0.0;
  __syncwarp();
  // inverse: lane c owns column c: x_cc = 1/l_cc, x_ic = -(sum_{k=c}^{i-1} l_ik x_kc) / l_ii   (l_ik broadcast from smem)
  double x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) {
      const double lik = Dbuf[i * LDD + k];
      if (k & 1) s1 = fma(-lik, x[k], s1); else s0 = fma(-lik, x[k], s0);
    }
    x[i] = (i >= lane) ? (s0 + s1) / Dbuf[i * LDD + i] : 0.0;
    Ibuf[i * LD + lane] = x[i];
  }
  long long c2 = clock64();
  *t_loop = c1 - c0; *t_tail = c2 - c1;
  return info;
}

// 1/sqrt(d) for a positive normal d: hardware seed (MUFU.RSQ64H, ~2^-20) + one third-order step, four dependent FP64 operations
// instead of the library routine's two Newton steps and special-case handling (the caller has checked d > 0).
__device__ __forceinline__ double rsqrt_fast(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  const double q = y * e;
  return fma(q, p, y);
}

// G3 = G2 with rsqrt_fast
__device__ __noinline__ int factor_rl3(double* Dbuf, double* Ibuf, int lane, long long* t_loop, long long* t_tail) {
  double a[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) a[k] = Dbuf[lane * LDD + k];
  int info = 0;
  long long c0 = clock64();
  double myinv = 0.0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double d = __shfl_sync(kFull, a[j], j);
    if (!(d > 0.0) && info == 0) info = j + 1;
    const double inv = rsqrt_fast(d);
    if (lane == j) myinv = inv;
    const double lij = lane > j ? a[j] * inv : (lane == j ? d * inv : 0.0);
    a[j] = lij;
#pragma unroll
    for (int k = j + 1; k < 32; ++k) {
      const double lkj = __shfl_sync(kFull, lij, k);
      a[k] = fma(-lij, lkj, a[k]);
    }
  }
  long long c1 = clock64();
#pragma unroll
  for (int k = 0; k < 32; ++k) Dbuf[lane * LDD + k] = k <= lane ? a[k] : 0.0;
  __syncwarp();
  double r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double ij = __shfl_sync(kFull, myinv, j);
    const double x = r[j] * ij;
    Ibuf[j * LD + lane] = x;
#pragma unroll
    for (int i = j + 1; i < 32; ++i) r[i] = fma(-Dbuf[i * LDD + j], x, r[i]);
  }
  long long c2 = clock64();
  *t_loop = c1 - c0; *t_tail = c2 - c1;
  return info;
}

// dependent-issue latency of DFMA and of the two rsqrt forms (cycles per operation in a chain of 256)
__global__ void k_lat(double* out, long long* cyc, double a, double b) {
  double x = a;
  long long c0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 256; ++i) x = fma(x, b, a);
  long long c1 = clock64();
  double y = fabs(x) + 1.0;
#pragma unroll 16
  for (int i = 0; i < 256; ++i) y = rsqrt(y) + 1.5;
  long long c2 = clock64();
  double z = fabs(y) + 1.0;
#pragma unroll 16
  for (int i = 0; i < 256; ++i) z = rsqrt_fast(z) + 1.5;
  long long c3 = clock64();
  double w = z;
#pragma unroll 16
  for (int i = 0; i < 256; ++i) w = __shfl_sync(kFull, w, (i * 7) & 31);
  long long c4 = clock64();
  if (threadIdx.x == 0) { cyc[0] = (c1 - c0) / 256; cyc[1] = (c2 - c1) / 256; cyc[2] = (c3 - c2) / 256; cyc[3] = (c4 - c3) / 256; }
  out[threadIdx.x] = x + y + z + w;
}

// G2: right-looking Cholesky in registers (as G) + RIGHT-LOOKING inverse in registers: lane c owns column c of X as a residual
// r (column c of I); step j: x_jc = r_j / l_jj, then r_i -= l_ij x_jc for i > j (independent FMAs, l_ij broadcast from smem).
__device__ __noinline__ int factor_rl2(double* Dbuf, double* Ibuf, int lane, long long* t_loop, long long* t_tail) {
  double a[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) a[k] = Dbuf[lane * LDD + k];
  int info = 0;
  long long c0 = clock64();
  double myinv = 0.0;                 // 1 / l_jj of row j == lane
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double d = __shfl_sync(kFull, a[j], j);
    if (!(d > 0.0) && info == 0) info = j + 1;
    const double inv = rsqrt(d);
    if (lane == j) myinv = inv;
    const double lij = lane > j ? a[j] * inv : (lane == j ? d * inv : 0.0);
    a[j] = lij;
#pragma unroll
    for (int k = j + 1; k < 32; ++k) {
      const double lkj = __shfl_sync(kFull, lij, k);
      a[k] = fma(-lij, lkj, a[k]);
    }
  }
  long long c1 = clock64();
#pragma unroll
  for (int k = 0; k < 32; ++k) Dbuf[lane * LDD + k] = k <= lane ? a[k] : 0.0;
  __syncwarp();
  double r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double ij = __shfl_sync(kFull, myinv, j);
    const double x = r[j] * ij;                       // zero for lane > j (r_j stays 0 there)
    Ibuf[j * LD + lane] = x;
#pragma unroll
    for (int i = j + 1; i < 32; ++i) r[i] = fma(-Dbuf[i * LDD + j], x, r[i]);
  }
  long long c2 = clock64();
  *t_loop = c1 - c0; *t_tail = c2 - c1;
  return info;
}

// S: right-looking Cholesky with the block in SHARED memory (no register array: fits any register cap), inverse as in G2 but
// with the residual column in shared memory too (Ibuf column = lane).
__device__ __noinline__ int factor_sm(double* Dbuf, double* Ibuf, int lane, long long* t_loop, long long* t_tail) {
  int info = 0;
  long long c0 = clock64();
  double* row = Dbuf + lane * LDD;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double d = Dbuf[j * LDD + j];
    if (!(d > 0.0) && info == 0) info = j + 1;
    const double inv = rsqrt(d);
    const double lij = lane > j ? row[j] * inv : (lane == j ? d * inv : 0.0);
    row[j] = lij;
    __syncwarp();
#pragma unroll
    for (int k = j + 1; k < 32; ++k) row[k] = fma(-lij, Dbuf[k * LDD + j], row[k]);
    __syncwarp();
  }
  long long c1 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) Ibuf[i * LD + lane] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double x = Ibuf[j * LD + lane] / Dbuf[j * LDD + j];
    Ibuf[j * LD + lane] = x;
#pragma unroll
    for (int i = j + 1; i < 32; ++i) Ibuf[i * LD + lane] = fma(-Dbuf[i * LDD + j], x, Ibuf[i * LD + lane]);
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 32; ++k) if (k > lane) row[k] = 0.0;
  long long c2 = clock64();
  *t_loop = c1 - c0; *t_tail = c2 - c1;
  return info;
}

template <int V>
__global__ void __launch_bounds__(32, 24) k_bench(const double* __restrict__ A, double* __restrict__ out, long long* cyc, int reps) {
  __shared__ double Dbuf[32 * LDD];
  __shared__ double Ibuf[32 * LD];
  const int lane = threadIdx.x;
  long long tot = 0, tl = 0, tt = 0;
  int info = 0;
  for (int r = 0; r < reps; ++r) {
    for (int k = 0; k < 32; ++k) Dbuf[lane * LDD + k] = A[(size_t)blockIdx.x * 1024 + lane * 32 + k];
    __syncwarp();
    long long a, b, c0 = clock64();
    if (V == 0) info |= factor_ll<true>(Dbuf, Ibuf, lane, &a, &b);
    else if (V == 1) info |= factor_ll<false>(Dbuf, Ibuf, lane, &a, &b);
    else if (V == 2) info |= factor_rl(Dbuf, Ibuf, lane, &a, &b);
    else if (V == 3) info |= factor_rl2(Dbuf, Ibuf, lane, &a, &b);
    else if (V == 5) info |= factor_rl3(Dbuf, Ibuf, lane, &a, &b);
    else info |= factor_sm(Dbuf, Ibuf, lane, &a, &b);
    tot += clock64() - c0; tl += a; tt += b;
    __syncwarp();
  }
  for (int k = 0; k < 32; ++k) {
    out[(size_t)blockIdx.x * 2048 + lane * 32 + k] = Dbuf[lane * LDD + k];
    out[(size_t)blockIdx.x * 2048 + 1024 + lane * 32 + k] = Ibuf[lane * LD + k];
  }
  if (lane == 0) { cyc[blockIdx.x * 4 + 0] = tot / reps; cyc[blockIdx.x * 4 + 1] = tl / reps; cyc[blockIdx.x * 4 + 2] = tt / reps; cyc[blockIdx.x * 4 + 3] = info; }
}

int main() {
  const int reps = 4;
  {
    double* dO; long long* dC; long long c[4];
    CK(cudaMalloc(&dO, 1024)); CK(cudaMalloc(&dC, 64));
    k_lat<<<1, 32>>>(dO, dC, 0.3, 0.999);
    CK(cudaMemcpy(c, dC, 32, cudaMemcpyDeviceToHost));
    printf("dependent latency, cycles: DFMA %lld, rsqrt(double)+add %lld, rsqrt_fast+add %lld, SHFL.64 %lld\n", c[0], c[1], c[2], c[3]);
  }
  for (int nb : {148, 148 * 9}) {
    std::vector<double> A((size_t)nb * 1024);
    for (int b = 0; b < nb; ++b)
      for (int i = 0; i < 32; ++i)
        for (int k = 0; k < 32; ++k) A[(size_t)b * 1024 + i * 32 + k] = exp(-0.05 * (i - k) * (i - k)) + (i == k ? 0.01 : 0.0);
    double *dA, *dO; long long* dC;
    CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dO, (size_t)nb * 2048 * 8)); CK(cudaMalloc(&dC, (size_t)nb * 32));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    std::vector<double> ref;
    for (int v = 0; v < 6; ++v) {
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      for (int w = 0; w < 2; ++w) {
        CK(cudaEventRecord(e0));
        if (v == 0) k_bench<0><<<nb, 32>>>(dA, dO, dC, reps);
        else if (v == 1) k_bench<1><<<nb, 32>>>(dA, dO, dC, reps);
        else if (v == 2) k_bench<2><<<nb, 32>>>(dA, dO, dC, reps);
        else if (v == 3) k_bench<3><<<nb, 32>>>(dA, dO, dC, reps);
        else if (v == 4) k_bench<4><<<nb, 32>>>(dA, dO, dC, reps);
        else k_bench<5><<<nb, 32>>>(dA, dO, dC, reps);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      }
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      std::vector<long long> c((size_t)nb * 4); std::vector<double> o((size_t)nb * 2048);
      CK(cudaMemcpy(c.data(), dC, c.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o.data(), dO, o.size() * 8, cudaMemcpyDeviceToHost));
      double err = 0;
      if (v == 0) ref = o; else for (size_t i = 0; i < o.size(); ++i) err = fmax(err, fabs(o[i] - ref[i]));
      // residual L X = I of block 0
      double res = 0;
      for (int i = 0; i < 32; ++i) for (int k = 0; k < 32; ++k) { double s = 0; for (int m = 0; m < 32; ++m) s += o[i * 32 + m] * o[1024 + m * 32 + k]; res = fmax(res, fabs(s - (i == k))); }
      printf("blocks %5d variant %c: %7lld cycles per factor (part1 %7lld part2 %7lld) info %lld kernel %.1f us  max|diff vs U| %.2e  |L X - I| %.2e\n", nb, "URG2S3"[v], c[0], c[1], c[2], c[3], ms * 1e3, err, res);
    }
    cudaFree(dA); cudaFree(dO); cudaFree(dC);
  }
  return 0;
}
