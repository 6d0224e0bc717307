// Query-side BSP work on the device: home leaf (findpartition), neighbour leaves
// (findneighbourpartitions), mixture weights, pair-list construction and the final convex combine.
//
// Replaces reference src/patchwork/partition.jl:248-262, src/RKHS/mixtureGP.jl:339-405 and the
// weight / normalise / combine lines of querymixtureGP! (mixtureGP.jl:224-272).
// Every floating-point value that feeds a comparison is computed with the oracle's operation order
// (explicit __dmul_rn/__dadd_rn, no FMA contraction), so leaf ids and neighbour lists are bit-exact.
#include "pmk_internal.cuh"

namespace pmk {

template <int D>
__device__ __forceinline__ int descend(const TreeDev& tr, const double* x) {
  const int Lv = tr.levels - 1;
  int node = 0, leaf = 0;
  for (int d = 0; d < Lv; ++d) {
    double v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = tr.hv[k * tr.n_hp + node];
    const double h = dot_seq<D>(v, x);
    const int right = !(h < tr.hc[node]);                 // partition.jl:254: dot(v,x) < c ? left : right
    leaf = leaf * 2 + right;
    node += right ? (1 << (Lv - 1 - d)) : 1;
  }
  return leaf + 1;
}

template <int D>
__global__ void k_home(TreeDev tr, int64_t Nq, const double* __restrict__ Xq, int32_t* __restrict__ home) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= Nq) return;
  double x[D];
#pragma unroll
  for (int d = 0; d < D; ++d) x[d] = Xq[j * D + d];
  home[j] = descend<D>(tr, x);
}

// Pass FILL=false counts the slots of every query (kept neighbours + 1); FILL=true writes them.
// Slot order = reference order: kept hyperplanes in increasing index, home leaf last.
template <int D, bool FILL>
__global__ void k_neighbours(TreeDev tr, QueryPlan q, double radius, double delta, int wkind, double wparam,
                             int32_t* __restrict__ leaf_count /* global leaf ids, 0-based slot */) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= q.Nq) return;
  double p[D];
#pragma unroll
  for (int d = 0; d < D; ++d) p[d] = q.Xq[j * D + d];
  const int home = q.home[j];
  int64_t slot = FILL ? q.pair_off[j] : 0;
  int kept = 0;
  for (int i = 0; i < tr.n_hp; ++i) {
    double u[D];
#pragma unroll
    for (int d = 0; d < D; ++d) u[d] = tr.hv[d * tr.n_hp + i];
    const double c = tr.hc[i];
    const double t = __dadd_rn(-dot_seq<D>(u, p), c);            // mixtureGP.jl:361  t = -dot(u,p) + c
    // z = p + t.*u ; norm(z - p)                                   mixtureGP.jl:362,367
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double z = __dadd_rn(p[d], __dmul_rn(t, u[d]));
      const double dd = __dsub_rn(z, p[d]);
      s = (d == 0) ? __dmul_rn(dd, dd) : __dadd_rn(s, __dmul_rn(dd, dd));
    }
    if (__dsqrt_rn(s) < radius) {
      double z1[D], z2[D];
      const double tp = __dadd_rn(t, delta), tm = __dsub_rn(t, delta);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        z1[d] = __dadd_rn(p[d], __dmul_rn(tp, u[d]));              // mixtureGP.jl:370-371
        z2[d] = __dadd_rn(p[d], __dmul_rn(tm, u[d]));
      }
      const int r1 = descend<D>(tr, z1);
      const int r2 = descend<D>(tr, z2);
      if ((r2 == home) != (r1 == home)) {                          // mixtureGP.jl:387 xor
        if (FILL) {
          const int nb = (r1 == home) ? r2 : r1;                   // mixtureGP.jl:392-395
          q.pair_leaf[slot] = nb;
          q.pair_q[slot] = (int32_t)j;
          q.pair_hp[slot] = i + 1;
          q.pair_t[slot] = t;
          q.pair_w[slot] = k_tau(wkind, wparam, fabs(t));          // mixtureGP.jl:231
          atomicAdd(&leaf_count[nb - 1], 1);
          ++slot;
        }
        ++kept;
      }
    }
  }
  if (FILL) {
    q.pair_leaf[slot] = home;                                       // mixtureGP.jl:237-239, w[end] = 1
    q.pair_q[slot] = (int32_t)j;
    q.pair_hp[slot] = 0;
    q.pair_t[slot] = 0.0;
    q.pair_w[slot] = 1.0;
    atomicAdd(&leaf_count[home - 1], 1);
  } else {
    q.npairs[j] = kept + 1;
  }
}

// Yq = dot(w,u), Vq = dot(w, v.*w) with w = w_tilde / sum(w_tilde)   (mixtureGP.jl:263-272)
__global__ void k_combine(int64_t Nq, const int64_t* __restrict__ pair_off, const double* __restrict__ pw,
                          const double* __restrict__ pu, const double* __restrict__ pv, double* __restrict__ Yq,
                          double* __restrict__ Vq, int mean_only) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= Nq) return;
  const int64_t a = pair_off[j], b = pair_off[j + 1];
  double sw = 0.0;
  for (int64_t s = a; s < b; ++s) sw = __dadd_rn(sw, pw[s]);
  double y = 0.0, v = 0.0;
  for (int64_t s = a; s < b; ++s) {
    const double w = __ddiv_rn(pw[s], sw);
    y = __dadd_rn(y, __dmul_rn(w, pu[s]));
    if (!mean_only) v = __dadd_rn(v, __dmul_rn(w, __dmul_rn(pv[s], w)));
  }
  Yq[j] = y;
  if (!mean_only) Vq[j] = v;
}

// single-block exclusive scan of leaf pair counts -> leaf pair starts (n <= a few 10^4)
__global__ void k_scan_small(const int32_t* __restrict__ in, int64_t* __restrict__ out, int n) {
  __shared__ int64_t carry;
  __shared__ int64_t buf[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    int64_t v = i < n ? in[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      int64_t add = threadIdx.x >= off ? buf[threadIdx.x - off] : 0;
      __syncthreads();
      buf[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < n) out[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

// ---------------------------------------------------------------------------------------------
void launch_home(int D, const TreeDev& tr, int64_t Nq, const double* dXq, int32_t* d_home, cudaStream_t s) {
  const int T = 256;
  const unsigned B = (unsigned)((Nq + T - 1) / T);
  if (B == 0) return;
  switch (D) {
    case 1: k_home<1><<<B, T, 0, s>>>(tr, Nq, dXq, d_home); break;
    case 2: k_home<2><<<B, T, 0, s>>>(tr, Nq, dXq, d_home); break;
    case 3: k_home<3><<<B, T, 0, s>>>(tr, Nq, dXq, d_home); break;
    default: break;
  }
}

template <bool FILL>
static void launch_nb_t(int D, const TreeDev& tr, const QueryPlan& q, double radius, double delta, int wkind,
                        double wparam, int32_t* d_leaf_count, cudaStream_t s) {
  const int T = 128;
  const unsigned B = (unsigned)((q.Nq + T - 1) / T);
  if (B == 0) return;
  switch (D) {
    case 1: k_neighbours<1, FILL><<<B, T, 0, s>>>(tr, q, radius, delta, wkind, wparam, d_leaf_count); break;
    case 2: k_neighbours<2, FILL><<<B, T, 0, s>>>(tr, q, radius, delta, wkind, wparam, d_leaf_count); break;
    case 3: k_neighbours<3, FILL><<<B, T, 0, s>>>(tr, q, radius, delta, wkind, wparam, d_leaf_count); break;
    default: break;
  }
}

void launch_neighbours(int D, bool fill, const TreeDev& tr, const QueryPlan& q, double radius, double delta, int wkind,
                       double wparam, int32_t* d_leaf_count, cudaStream_t s) {
  if (fill) launch_nb_t<true>(D, tr, q, radius, delta, wkind, wparam, d_leaf_count, s);
  else launch_nb_t<false>(D, tr, q, radius, delta, wkind, wparam, d_leaf_count, s);
}

void launch_combine(int64_t Nq, const int64_t* pair_off, const double* pw, const double* pu, const double* pv,
                    double* dYq, double* dVq, int mean_only, cudaStream_t s) {
  const int T = 256;
  const unsigned B = (unsigned)((Nq + T - 1) / T);
  if (B == 0) return;
  k_combine<<<B, T, 0, s>>>(Nq, pair_off, pw, pu, pv, dYq, dVq, mean_only);
}

void launch_scan_small(const int32_t* in, int64_t* out, int n, cudaStream_t s) {
  k_scan_small<<<1, 1024, 0, s>>>(in, out, n);
}

}  // namespace pmk
