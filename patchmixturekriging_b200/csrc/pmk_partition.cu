// setuppartition on the device (reference src/patchwork/partition.jl:64-217), one tree level per step.
//
// What a node of the reference does (gethyperplane :86-100, splitpoints :64-83):
//   mu = mean(X)                     Base pairwise sum (blocks of <= 1024 summed left to right, longer ranges halved), / n
//   z  = X[1] - mu                   `size(X,2)` of a Vector is 1, so only the node's FIRST point enters
//   v  = V[:,1] of svd(z')           LAPACK dgesdd -- stays with the caller's LinearAlgebra (the bits of v are the host
//                                    LAPACK's; no device restatement can promise them), see pmk.h
//   f_n = dot(v, X[n]); c = median(f); left iff f_n < c; the children keep the global order of their points
// The device keeps the training points in place and a permutation `perm` of their ids in which every node of the current
// level is a contiguous, ascending segment (left-first DFS order of the nodes = position order, so after the last level
// the segments ARE the leaves in AbstractTrees.Leaves order).  Per level:
//   k_part_block_sums + k_part_node_z : mu and z of every node, in the reference's summation order
//   k_part_project                    : f for every point (un-fused multiply/add, the order of the oracle's dot_seq)
//   two radix sorts (by f, then stably by node id) and k_part_median : c of every node
//   k_part_flags, exclusive scan, k_part_scatter, k_part_child_offsets : stable split of every segment
// HBM-bound integer/FP64 streaming work, N points per level; nothing here is GEMM-shaped.
#include <cstdint>
#include <vector>

#include "pmk_common.cuh"

namespace pmk {

// ---- host: the summation plan of Base.mapreduce_impl(+, A, ifirst, ilast, 1024) for one range of n elements -----------
// Emits the sequential blocks in order, each with its depth in the halving tree; the pairwise combination is then
// "push the block sums in order, and while the two topmost entries have the same depth, replace them by their sum one
// level up" -- exactly the recursion's order of additions.
void partition_sum_plan(int64_t n, int64_t base, std::vector<int64_t>& start, std::vector<int32_t>& len, std::vector<int32_t>& depth) {
  struct Fr { int64_t lo, hi; int d; };
  std::vector<Fr> st;
  st.push_back({0, n - 1, 0});
  while (!st.empty()) {
    Fr f = st.back();
    st.pop_back();
    if (f.hi - f.lo < 1024) {
      start.push_back(base + f.lo);
      len.push_back((int32_t)(f.hi - f.lo + 1));
      depth.push_back(f.d);
    } else {
      const int64_t mid = f.lo + ((f.hi - f.lo) >> 1);
      st.push_back({mid + 1, f.hi, f.d + 1});   // popped second: blocks come out left to right
      st.push_back({f.lo, mid, f.d + 1});
    }
  }
}

// ---- device ------------------------------------------------------------------------------------------------------------
// One thread per sequential block: s = x[0]; s += x[i] in order (partition.jl:89 through Base's mapreduce).
template <int D>
__global__ void k_part_block_sums(const double* __restrict__ X, const int32_t* __restrict__ perm, const int64_t* __restrict__ blk_start,
                                  const int32_t* __restrict__ blk_len, int n_blk, double* __restrict__ blk_sum) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blk) return;
  const int32_t* p = perm + blk_start[b];
  const int n = blk_len[b];
  double s[D];
  {
    const double* x = X + (int64_t)p[0] * D;
#pragma unroll
    for (int d = 0; d < D; ++d) s[d] = x[d];
  }
  int i = 1;
  for (; i + 4 <= n; i += 4) {          // four independent gathers in flight, added in order
    double t[4][D];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double* x = X + (int64_t)p[i + u] * D;
#pragma unroll
      for (int d = 0; d < D; ++d) t[u][d] = x[d];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int d = 0; d < D; ++d) s[d] = __dadd_rn(s[d], t[u][d]);
  }
  for (; i < n; ++i) {
    const double* x = X + (int64_t)p[i] * D;
#pragma unroll
    for (int d = 0; d < D; ++d) s[d] = __dadd_rn(s[d], x[d]);
  }
#pragma unroll
  for (int d = 0; d < D; ++d) blk_sum[(int64_t)b * D + d] = s[d];
}

// One thread per node: pairwise combination of its block sums, mu = sum / n, z = X[first point] - mu.
template <int D>
__global__ void k_part_node_z(const double* __restrict__ X, const int32_t* __restrict__ perm, const int64_t* __restrict__ seg_off,
                              const int32_t* __restrict__ node_blk_off, const int32_t* __restrict__ blk_depth,
                              const double* __restrict__ blk_sum, int n_nodes, double* __restrict__ z_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_nodes) return;
  double val[40][D];
  int dep[40];
  int top = 0;
  for (int b = node_blk_off[j]; b < node_blk_off[j + 1]; ++b) {
#pragma unroll
    for (int d = 0; d < D; ++d) val[top][d] = blk_sum[(int64_t)b * D + d];
    dep[top] = blk_depth[b];
    ++top;
    while (top >= 2 && dep[top - 1] == dep[top - 2]) {
#pragma unroll
      for (int d = 0; d < D; ++d) val[top - 2][d] = __dadd_rn(val[top - 2][d], val[top - 1][d]);
      --dep[top - 2];
      --top;
    }
  }
  const int64_t s = seg_off[j];
  const double n = (double)(seg_off[j + 1] - s);
  const double* x = X + (int64_t)perm[s] * D;
#pragma unroll
  for (int d = 0; d < D; ++d) z_out[(int64_t)j * D + d] = __dsub_rn(x[d], __ddiv_rn(val[0][d], n));
}

__device__ __forceinline__ int node_of_position(const int64_t* __restrict__ seg_off, int n_nodes, int64_t i) {
  int lo = 0, hi = n_nodes;      // largest j with seg_off[j] <= i  (empty segments cannot occur: the API rejects them)
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (seg_off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// f_i = dot(v_node, X[perm[i]]): products rounded, added left to right (oracle dot_seq; partition.jl:69).
template <int D>
__global__ void k_part_project(const double* __restrict__ X, const int32_t* __restrict__ perm, const int64_t* __restrict__ seg_off,
                               int n_nodes, const double* __restrict__ v, int64_t N, double* __restrict__ f, int32_t* __restrict__ node_id) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int j = node_of_position(seg_off, n_nodes, i);
  const double* x = X + (int64_t)perm[i] * D;
  const double* vj = v + (int64_t)j * D;
  double s = __dmul_rn(vj[0], x[0]);
#pragma unroll
  for (int d = 1; d < D; ++d) s = __dadd_rn(s, __dmul_rn(vj[d], x[d]));
  f[i] = s;
  node_id[i] = j;
}

// Statistics.median (partition.jl:70): odd n -> the middle element, even n -> a/2 + b/2 of the two middle ones.
__global__ void k_part_median(const double* __restrict__ f_sorted, const int64_t* __restrict__ seg_off, int n_nodes,
                              double* __restrict__ c_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_nodes) return;
  const int64_t s = seg_off[j], n = seg_off[j + 1] - s;
  if (n & 1) {
    c_out[j] = f_sorted[s + (n - 1) / 2];
  } else {
    const double a = f_sorted[s + n / 2 - 1], b = f_sorted[s + n / 2];
    c_out[j] = __dadd_rn(__ddiv_rn(a, 2.0), __ddiv_rn(b, 2.0));
  }
}

__global__ void k_part_flags(const double* __restrict__ f, const int32_t* __restrict__ node_id, const double* __restrict__ c, int64_t N,
                             int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > N) return;
  flag[i] = (i < N && f[i] < c[node_id[i]]) ? 1 : 0;     // flag[N] = 0 closes the exclusive scan
}

// Stable split: the points with f < c keep their order at the front of the node's segment, the others behind them.
__global__ void k_part_scatter(const int32_t* __restrict__ perm, const int32_t* __restrict__ node_id, const int32_t* __restrict__ flag,
                               const int32_t* __restrict__ scan, const int64_t* __restrict__ seg_off, int64_t N, int32_t* __restrict__ perm_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int j = node_id[i];
  const int64_t s = seg_off[j], e = seg_off[j + 1];
  const int64_t left_total = scan[e] - scan[s];
  const int64_t left_before = scan[i] - scan[s];
  const int64_t dst = flag[i] ? s + left_before : s + left_total + (i - s) - left_before;
  perm_out[dst] = perm[i];
}

__global__ void k_part_child_offsets(const int32_t* __restrict__ scan, const int64_t* __restrict__ seg_off, int n_nodes, int64_t N,
                                     int64_t* __restrict__ child_off) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n_nodes) return;
  if (j == n_nodes) {
    child_off[2 * (int64_t)n_nodes] = N;
    return;
  }
  const int64_t s = seg_off[j], e = seg_off[j + 1];
  child_off[2 * (int64_t)j] = s;
  child_off[2 * (int64_t)j + 1] = s + (scan[e] - scan[s]);
}

__global__ void k_part_iota(int32_t* __restrict__ perm, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) perm[i] = (int32_t)i;
}

__global__ void k_part_one_based(const int32_t* __restrict__ perm, int64_t N, int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = perm[i] + 1;
}

// ---- launchers ---------------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

void launch_part_iota(int32_t* perm, int64_t N, cudaStream_t s) { k_part_iota<<<blocks_for(N, 256), 256, 0, s>>>(perm, N); }

void launch_part_one_based(const int32_t* perm, int64_t N, int32_t* out, cudaStream_t s) {
  k_part_one_based<<<blocks_for(N, 256), 256, 0, s>>>(perm, N, out);
}

void launch_part_block_sums(int D, const double* X, const int32_t* perm, const int64_t* blk_start, const int32_t* blk_len, int n_blk,
                            double* blk_sum, cudaStream_t s) {
  const unsigned g = blocks_for(n_blk, 64);
  if (D == 1) k_part_block_sums<1><<<g, 64, 0, s>>>(X, perm, blk_start, blk_len, n_blk, blk_sum);
  else if (D == 2) k_part_block_sums<2><<<g, 64, 0, s>>>(X, perm, blk_start, blk_len, n_blk, blk_sum);
  else k_part_block_sums<3><<<g, 64, 0, s>>>(X, perm, blk_start, blk_len, n_blk, blk_sum);
}

void launch_part_node_z(int D, const double* X, const int32_t* perm, const int64_t* seg_off, const int32_t* node_blk_off,
                        const int32_t* blk_depth, const double* blk_sum, int n_nodes, double* z_out, cudaStream_t s) {
  const unsigned g = blocks_for(n_nodes, 64);
  if (D == 1) k_part_node_z<1><<<g, 64, 0, s>>>(X, perm, seg_off, node_blk_off, blk_depth, blk_sum, n_nodes, z_out);
  else if (D == 2) k_part_node_z<2><<<g, 64, 0, s>>>(X, perm, seg_off, node_blk_off, blk_depth, blk_sum, n_nodes, z_out);
  else k_part_node_z<3><<<g, 64, 0, s>>>(X, perm, seg_off, node_blk_off, blk_depth, blk_sum, n_nodes, z_out);
}

void launch_part_project(int D, const double* X, const int32_t* perm, const int64_t* seg_off, int n_nodes, const double* v, int64_t N,
                         double* f, int32_t* node_id, cudaStream_t s) {
  const unsigned g = blocks_for(N, 256);
  if (D == 1) k_part_project<1><<<g, 256, 0, s>>>(X, perm, seg_off, n_nodes, v, N, f, node_id);
  else if (D == 2) k_part_project<2><<<g, 256, 0, s>>>(X, perm, seg_off, n_nodes, v, N, f, node_id);
  else k_part_project<3><<<g, 256, 0, s>>>(X, perm, seg_off, n_nodes, v, N, f, node_id);
}

void launch_part_median(const double* f_sorted, const int64_t* seg_off, int n_nodes, double* c_out, cudaStream_t s) {
  k_part_median<<<blocks_for(n_nodes, 128), 128, 0, s>>>(f_sorted, seg_off, n_nodes, c_out);
}

void launch_part_flags(const double* f, const int32_t* node_id, const double* c, int64_t N, int32_t* flag, cudaStream_t s) {
  k_part_flags<<<blocks_for(N + 1, 256), 256, 0, s>>>(f, node_id, c, N, flag);
}

void launch_part_scatter(const int32_t* perm, const int32_t* node_id, const int32_t* flag, const int32_t* scan, const int64_t* seg_off,
                         int64_t N, int32_t* perm_out, cudaStream_t s) {
  k_part_scatter<<<blocks_for(N, 256), 256, 0, s>>>(perm, node_id, flag, scan, seg_off, N, perm_out);
}

void launch_part_child_offsets(const int32_t* scan, const int64_t* seg_off, int n_nodes, int64_t N, int64_t* child_off, cudaStream_t s) {
  k_part_child_offsets<<<blocks_for(n_nodes + 1, 128), 128, 0, s>>>(scan, seg_off, n_nodes, N, child_off);
}

}  // namespace pmk
