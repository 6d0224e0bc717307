// Fit path: fused Gram + batched blocked Cholesky (K1+K2) and the triangular solves for alpha.
//
// Replaces the body of fitmixtureGP! (reference src/RKHS/mixtureGP.jl:92-115):
//     U = constructkernelmatrix(X, θ); U[i,i] += σ²; c = U\y; L = cholesky(U).L
// One CTA per BSP leaf.  Left-looking blocked Cholesky with 32-column panels:
//   panel J:  C = K[J:, J] - L[J:, 0:J] * L[J, 0:J]^T
//             (K evaluated on the fly straight into the accumulators -- the Gram matrix is never
//              written to HBM; the trailing update runs on DMMA.8x8x4 with both operands read as
//              packed fragment tiles, 512 B coalesced per warp load)
//             diagonal block: Cholesky + explicit 32x32 inverse by warp 0 (registers + shuffles)
//             L[J+1:, J] = C * inv(L_JJ)^T on DMMA, stored once in packed-tile form.
// Several CTAs are resident per SM so that one leaf's serial diagonal-block phase overlaps the
// other leaves' DMMA phases.
#include "pmk_internal.cuh"

namespace pmk {

static constexpr unsigned kFull = 0xffffffffu;
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
__device__ __noinline__ int factor_block32(double* Dbuf, double* Ibuf, int lane) {
  // left-looking (dot-product form) Cholesky, column j: lane i >= j owns L[i][j]
  for (int j = 0; j < 32; ++j) {
    double s0 = Dbuf[lane * LDD + j], s1 = 0.0;
    int k = 0;
    for (; k + 1 < j; k += 2) {
      s0 = fma(-Dbuf[lane * LDD + k], Dbuf[j * LDD + k], s0);
      s1 = fma(-Dbuf[lane * LDD + k + 1], Dbuf[j * LDD + k + 1], s1);
    }
    if (k < j) s0 = fma(-Dbuf[lane * LDD + k], Dbuf[j * LDD + k], s0);
    const double s = s0 + s1;
    const double d = __shfl_sync(kFull, s, j);
    if (!(d > 0.0)) return j + 1;                 // uniform: d is a broadcast
    const double ljj = sqrt(d);
    const double inv = 1.0 / ljj;
    __syncwarp();
    if (lane == j) Dbuf[lane * LDD + j] = ljj;
    else if (lane > j) Dbuf[lane * LDD + j] = s * inv;
    else Dbuf[lane * LDD + j] = 0.0;              // zero the upper triangle
    __syncwarp();
  }
  // column `lane` of X = inv(L): x_i = (delta_i,lane - sum_{k<i} L[i][k] x_k) / L[i][i]
  for (int i = 0; i < 32; ++i) {
    double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
    int k = 0;
    for (; k + 1 < i; k += 2) {
      s0 = fma(-Dbuf[i * LDD + k], Ibuf[k * LD + lane], s0);
      s1 = fma(-Dbuf[i * LDD + k + 1], Ibuf[(k + 1) * LD + lane], s1);
    }
    if (k < i) s0 = fma(-Dbuf[i * LDD + k], Ibuf[k * LD + lane], s0);
    Ibuf[i * LD + lane] = (i >= lane) ? (s0 + s1) / Dbuf[i * LDD + i] : 0.0;
  }
  __syncwarp();
  return 0;
}

// trailing-update inner loop of k_chol for NV valid row tiles (NV is warp-uniform)
template <int R, int NV>
__device__ __forceinline__ void chol_kloop(double (&acc)[R][4][2], const double2* __restrict__ Lp, const int (&bo)[4],
                                           const int (&ao)[R], int nct) {
#pragma unroll 4
  for (int ct = 0; ct < nct; ++ct) {
    double2 bf[4], af[NV];
#pragma unroll
    for (int b = 0; b < 4; ++b) bf[b] = Lp[bo[b] + ct * 32];
#pragma unroll
    for (int r = 0; r < NV; ++r) af[r] = Lp[ao[r] + ct * 32];
#pragma unroll
    for (int r = 0; r < NV; ++r) {
#pragma unroll
      for (int b = 0; b < 4; ++b) dmma884(acc[r][b][0], acc[r][b][1], af[r].x, bf[b].x);
    }
#pragma unroll
    for (int r = 0; r < NV; ++r) {
#pragma unroll
      for (int b = 0; b < 4; ++b) dmma884(acc[r][b][0], acc[r][b][1], af[r].y, bf[b].y);
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <int D, int NW, int R>
__global__ void __launch_bounds__(NW * 32, 3)
k_chol(LeafTable lt, const int* __restrict__ order, KParams kp, double sigma2) {
  __shared__ double Dbuf[32 * LDD];
  __shared__ double Ibuf[32 * LD];
  __shared__ int s_fail;
  const int p = order[blockIdx.x];
  const int n = lt.n[p], npad = lt.npad[p];
  const int nblk = npad >> 5, ntl = npad >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, l = lane & 3;
  const double* __restrict__ xs = lt.xs + lt.xoff[p];
  const int64_t xstride = lt.xstride;
  double2* Lp = reinterpret_cast<double2*>(lt.L + lt.loff[p]);
  double2* Ip = reinterpret_cast<double2*>(lt.Linv + lt.ioff[p]);
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();

  const int src_lo = (lane & ~3) | (l >> 1);
  const int src_hi = (lane & ~3) | (2 + (l >> 1));

  for (int J = 0; J < nblk; ++J) {
    const int t0 = 4 * J;
    const int nchunks = (ntl - t0 + NW * R - 1) / (NW * R);
    for (int c = 0; c < nchunks; ++c) {
      int t[R];
      bool tv[R];
      double acc[R][4][2];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        t[r] = t0 + c * NW * R + warp + NW * r;
        tv[r] = t[r] < ntl;
      }
      // ---- Gram entries straight into the (negated) accumulators: acc = -(K + sigma2*I) ----------
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = 8 * (tv[r] ? t[r] : t0) + g;
        double xr[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xr[d] = xs[d * xstride + row];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = 32 * J + 8 * b + 2 * l + e;
            double kv = 0.0;
            if (tv[r] && col <= row) {
              if (row < n) {   // col <= row < n
                double xc[D];
#pragma unroll
                for (int d = 0; d < D; ++d) xc[d] = xs[d * xstride + col];
                kv = eval_kernel<D>(kp, xr, xc);      // evalkernel(X[i], X[j]), i >= j  (RKHS.jl:21-25)
                if (row == col) kv = __dadd_rn(kv, sigma2);   // mixtureGP.jl:102-104
              } else {
                kv = (row == col) ? 1.0 : 0.0;        // identity padding
              }
            }
            acc[r][b][e] = -kv;
          }
        }
      }
      // ---- acc += L[t, 0:J] * L[J, 0:J]^T on the FP64 tensor cores -------------------------------
      {
        int bo[4], ao[R];
#pragma unroll
        for (int b = 0; b < 4; ++b) bo[b] = (int)tri(t0 + b) * 32 + lane;
#pragma unroll
        for (int r = 0; r < R; ++r) ao[r] = (int)tri(tv[r] ? t[r] : t0) * 32 + lane;
        // the number of valid row tiles of this warp is uniform: pick a loop specialised for it, so
        // that no DMMA is predicated (a predicated-off DMMA still occupies the pipe) and the loads of
        // the unrolled iterations can be hoisted freely
        if (tv[R - 1]) chol_kloop<R, R>(acc, Lp, bo, ao, 4 * J);
        else if (R > 1 && tv[0]) chol_kloop<R, 1>(acc, Lp, bo, ao, 4 * J);
      }
      // ---- diagonal block: factor + invert (chunk 0 carries row tiles t0..t0+3 on warps 0..3) ----
      if (c == 0) {
        if (warp < 4) {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            Dbuf[(8 * warp + g) * LDD + 8 * b + 2 * l + 0] = -acc[0][b][0];
            Dbuf[(8 * warp + g) * LDD + 8 * b + 2 * l + 1] = -acc[0][b][1];
          }
        }
        __syncthreads();
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
        }
        __syncthreads();
        if (s_fail) return;
      }
      // ---- panel rows below the diagonal block: L[t, J] = C * inv(L_JJ)^T ------------------------
#pragma unroll
      for (int r = 0; r < R; ++r) {
        PMK_UNIFORM_IF(tv[r] && t[r] >= t0 + 4) {
          double alo[4], ahi[4];
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const double v0 = __shfl_sync(kFull, acc[r][kb][0], src_lo);
            const double v1 = __shfl_sync(kFull, acc[r][kb][1], src_lo);
            const double w0 = __shfl_sync(kFull, acc[r][kb][0], src_hi);
            const double w1 = __shfl_sync(kFull, acc[r][kb][1], src_hi);
            alo[kb] = -((l & 1) ? v1 : v0);
            ahi[kb] = -((l & 1) ? w1 : w0);
          }
          double* tile_row = reinterpret_cast<double*>(Lp + (tri(t[r]) + t0) * 32);
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            double o0 = 0.0, o1 = 0.0;
#pragma unroll
            for (int kb = 0; kb <= cb; ++kb) {
              const double blo = Ibuf[(8 * cb + g) * LD + 8 * kb + l];
              const double bhi = Ibuf[(8 * cb + g) * LD + 8 * kb + 4 + l];
              dmma884(o0, o1, alo[kb], blo);
              dmma884(o0, o1, ahi[kb], bhi);
            }
            // C-fragment (row g, cols 2l, 2l+1) -> packed fragment-major tile
            double* tile = tile_row + cb * 64;
            const int c0 = 2 * l, c1 = 2 * l + 1;
            tile[(g * 4 + (c0 & 3)) * 2 + (c0 >> 2)] = o0;
            tile[(g * 4 + (c1 & 3)) * 2 + (c1 >> 2)] = o1;
          }
        }
      }
    }
    __syncthreads();   // panel J complete and visible before panel J+1 reads it
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

template <int NW>
__global__ void __launch_bounds__(NW * 32)
k_solve_alpha(LeafTable lt, const int* __restrict__ order) {
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
  for (int i = threadIdx.x; i < npad; i += NW * 32) z[i] = lt.y[xo + i];
  __syncthreads();
  // ---- forward: z <- L^-1 y
  for (int J = 0; J < nblk; ++J) {
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
  for (int i = threadIdx.x; i < npad; i += NW * 32) lt.alpha[xo + i] = i < n ? z[i] : 0.0;
}

// dense column-major n x n lower-triangular copy of leaf p's factor (pmk_get_L)
__global__ void k_unpack_L(LeafTable lt, int p, double* __restrict__ out) {
  const int n = lt.n[p];
  const double* __restrict__ Lp = lt.L + lt.loff[p];
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

void launch_chol(int D, const LeafTable& lt, const int* d_order, int n_order, KParams kp, double sigma2, cudaStream_t s) {
  constexpr int NW = 8, R = 2;
  if (n_order <= 0) return;
  switch (D) {
    case 1: k_chol<1, NW, R><<<n_order, NW * 32, 0, s>>>(lt, d_order, kp, sigma2); break;
    case 2: k_chol<2, NW, R><<<n_order, NW * 32, 0, s>>>(lt, d_order, kp, sigma2); break;
    case 3: k_chol<3, NW, R><<<n_order, NW * 32, 0, s>>>(lt, d_order, kp, sigma2); break;
    default: break;
  }
}

void launch_solve(const LeafTable& lt, const int* d_order, int n_order, int max_npad, cudaStream_t s) {
  constexpr int NW = 8;
  if (n_order <= 0) return;
  const size_t smem = (size_t)(max_npad + NW * 32 + 32) * sizeof(double);
  k_solve_alpha<NW><<<n_order, NW * 32, smem, s>>>(lt, d_order);
}

void launch_unpack_L(const LeafTable& lt, int p, int n, double* d_out, cudaStream_t s) {
  int64_t total = (int64_t)n * n;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  if (blocks < 1) blocks = 1;
  k_unpack_L<<<blocks, 256, 0, s>>>(lt, p, d_out);
}

}  // namespace pmk
