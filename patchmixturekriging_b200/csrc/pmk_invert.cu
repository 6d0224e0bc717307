// P = inv(L) for every leaf by RECURSIVE DOUBLING on the packed tiles -- the operand of the explicit-inverse pair kernels
// (pmk_query_rowp.cuh), formed once per fit.
//
//   [ A  0 ]^-1   [  A^-1            0   ]
//   [ C  B ]    = [ -B^-1 C A^-1    B^-1 ]
//
// The 32x32 diagonal blocks of inv(L) are already there (Linv, written by the factorisation).  A leaf's block range is split
// recursively (left part = the largest power of two below the size); a node of the recursion tree only needs its two
// children, so all nodes of one HEIGHT -- of all leaves -- are independent and go into one launch pair:
//     T = C * A^-1     (k_inv_T: dense x lower-triangular; T parked in a scratch buffer, B-fragment-major)
//     P21 = -B^-1 * T  (k_inv_R: lower-triangular x dense; written into P's packed tiles)
// Every flop is a DMMA inside a 32x32 output block owned by one warp; there is no substitution chain anywhere, which is
// what held the former builder (the substitution pair kernel on identity right-hand sides) at a third of the FP64 peak.
// n^3/3 flops per leaf, the same as that builder.
#include <algorithm>
#include <vector>
#include "pmk_internal.cuh"

namespace pmk {

static constexpr int kInvWarps = 8;

// per shape (= number of 32-row blocks nb) and height: where that shape's nodes of that height sit in the node array
struct InvPlanDev {
  const InvNode* nodes;
  const int* off;               // [kInvMaxBlocks + 1][kInvMaxHeight + 1]
  const int* cnt;
};

// the inverse's diagonal 32x32 blocks: the 10 lower tiles of block J of Linv -> P tiles (4J + a, 4J + b)
__global__ void __launch_bounds__(256)
k_inv_diag(LeafTable lt, int first_leaf) {
  const int p = first_leaf + blockIdx.y;
  const int nb = lt.npad[p] >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int J = blockIdx.x * 8 + warp;
  if (J >= nb) return;
  const double2* __restrict__ src = reinterpret_cast<const double2*>(lt.Linv + lt.ioff[p]) + (size_t)J * (kInvTilesPerBlock * 32) + lane;
  double2* dst = reinterpret_cast<double2*>(lt.P + lt.loff[p]) + lane;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) dst[(tri(4 * J + a) + 4 * J + b) * 32] = src[(a * (a + 1) / 2 + b) * 32];
}

// Work of one height of one leaf = the 32x32 output blocks of all its nodes of that height, numbered node-major (node j owns
// items [j * bpn, j * bpn + its block count), bpn = the largest block count of a node of this height over all shapes).
// One warp = one item; a CTA takes kInvWarps consecutive items, so that at the low heights (one or four blocks per node) the
// warps of a CTA are spread over several nodes instead of idling, and at the top heights the CTAs working on one leaf at the
// same time share its tiles in L2 / L1.
__device__ __forceinline__ bool inv_item(const LeafTable& lt, const InvPlanDev& pl, int first_leaf, int height, int bpn, int warp, int& p,
                                         InvNode& nd, int& b) {
  p = first_leaf + blockIdx.y;
  const int nb = lt.npad[p] >> 5;
  const int idx = nb * (kInvMaxHeight + 1) + height;
  const int item = blockIdx.x * kInvWarps + warp;
  const int node = item / bpn;
  b = item % bpn;
  if (node >= pl.cnt[idx]) return false;
  nd = pl.nodes[pl.off[idx] + node];
  return b < (nd.hi - nd.mid) * (nd.mid - nd.lo);
}

// T = C * A^-1 for the node: T tile (ti, tj) = sum_{tk >= tj} L(ti, tk) * P(tk, tj), tk inside A's range.
// L tiles are the left operand as stored (A-fragment-major); the P tiles are needed as B fragments, i.e. transposed:
// two 8-byte loads per lane instead of one 16-byte load.  T goes to the scratch buffer at the tile's own packed position,
// stored B-fragment-major so that k_inv_R loads it with one LDG.128.
__global__ void __launch_bounds__(kInvWarps * 32, 2)
k_inv_T(LeafTable lt, InvPlanDev pl, double* __restrict__ scratch, int first_leaf, int height, int bpn) {
  int p, b;
  InvNode nd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!inv_item(lt, pl, first_leaf, height, bpn, warp, p, nd, b)) return;
  const int g = lane >> 2, l = lane & 3;
  const double2* __restrict__ Lp = reinterpret_cast<const double2*>(lt.L + lt.loff[p]) + lane;
  const double* __restrict__ Pd = lt.P + lt.loff[p];
  double* Td = scratch + lt.loff[p];
  // element (row l, column g) and (row l + 4, column g) of a packed tile = this lane's B fragments of two k-steps
  const int e0 = (l * 4 + (g & 3)) * 2 + (g >> 2), e1 = e0 + 32;
  const int nrb = nd.hi - nd.mid;
  {
    // column-major deal: the warps of a CTA work on consecutive ROW blocks of one column block, so they share the right operand
    // P(tk, 4bj ..) -- and its k range -- through L1 (k_inv_R shares its left operand the same way)
    const int bi = nd.mid + b % nrb, bj = nd.lo + b / nrb;
    double acc[4][4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
    size_t arow[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) arow[r] = tri(4 * bi + r) * 32;
    // the column block's own (triangular) diagonal block of A^-1: P(tk, tj) exists only for tk >= tj
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int tk = 4 * bj + kk;
      double2 af[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) af[r] = Lp[arow[r] + (size_t)tk * 32];
      const double* prow = Pd + (tri(tk) + 4 * bj) * 64;
#pragma unroll
      for (int c = 0; c <= kk; ++c) {
        const double b0 = prow[c * 64 + e0], b1 = prow[c * 64 + e1];
#pragma unroll
        for (int r = 0; r < 4; ++r) dmma884(acc[r][c][0], acc[r][c][1], af[r].x, b0);
#pragma unroll
        for (int r = 0; r < 4; ++r) dmma884(acc[r][c][0], acc[r][c][1], af[r].y, b1);
      }
    }
    // the full blocks below it
    for (int tk = 4 * bj + 4; tk < 4 * nd.mid; ++tk) {
      double2 af[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) af[r] = Lp[arow[r] + (size_t)tk * 32];
      const double* prow = Pd + (tri(tk) + 4 * bj) * 64;
      double b0[4], b1[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        b0[c] = prow[c * 64 + e0];
        b1[c] = prow[c * 64 + e1];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int r = 0; r < 4; ++r) dmma884(acc[r][c][0], acc[r][c][1], af[r].x, b0[c]);
#pragma unroll
        for (int r = 0; r < 4; ++r) dmma884(acc[r][c][0], acc[r][c][1], af[r].y, b1[c]);
      }
    }
    // C fragment (row g, columns 2l, 2l+1) -> B-fragment-major tile: element (r, c) at ((c*4 + (r&3))*2 + (r>>2))
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double* trow = Td + (tri(4 * bi + r) + 4 * bj) * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        trow[c * 64 + ((2 * l) * 4 + (g & 3)) * 2 + (g >> 2)] = acc[r][c][0];
        trow[c * 64 + ((2 * l + 1) * 4 + (g & 3)) * 2 + (g >> 2)] = acc[r][c][1];
      }
    }
  }
}

// P21 = -B^-1 * T for the node: P tile (ti, tj) = -sum_{tk <= ti} P(ti, tk) * T(tk, tj), tk inside B's range.
__global__ void __launch_bounds__(kInvWarps * 32, 2)
k_inv_R(LeafTable lt, InvPlanDev pl, const double* __restrict__ scratch, int first_leaf, int height, int bpn) {
  int p, b;
  InvNode nd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!inv_item(lt, pl, first_leaf, height, bpn, warp, p, nd, b)) return;
  const int g = lane >> 2, l = lane & 3;
  double* Pd = lt.P + lt.loff[p];
  const double2* Pp = reinterpret_cast<const double2*>(Pd) + lane;
  const double2* __restrict__ Tp = reinterpret_cast<const double2*>(scratch + lt.loff[p]) + lane;
  const int ncb = nd.mid - nd.lo;
  {
    const int bi = nd.mid + b / ncb, bj = nd.lo + b % ncb;
    double acc[4][4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
    size_t arow[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) arow[r] = tri(4 * bi + r) * 32;
    // full blocks of B^-1 left of the row block's own diagonal block
    for (int tk = 4 * nd.mid; tk < 4 * bi; ++tk) {
      double2 af[4], bf[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) af[r] = Pp[arow[r] + (size_t)tk * 32];
      const double2* trow = Tp + (tri(tk) + 4 * bj) * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) bf[c] = trow[c * 32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int r = 0; r < 4; ++r) dmma884(acc[r][c][0], acc[r][c][1], af[r].x, bf[c].x);
#pragma unroll
        for (int r = 0; r < 4; ++r) dmma884(acc[r][c][0], acc[r][c][1], af[r].y, bf[c].y);
      }
    }
    // the (triangular) diagonal block: P(4bi + r, tk) exists only for tk <= 4bi + r
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int tk = 4 * bi + kk;
      double2 af[4], bf[4];
#pragma unroll
      for (int r = kk; r < 4; ++r) af[r] = Pp[arow[r] + (size_t)tk * 32];
      const double2* trow = Tp + (tri(tk) + 4 * bj) * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) bf[c] = trow[c * 32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int r = kk; r < 4; ++r) dmma884(acc[r][c][0], acc[r][c][1], af[r].x, bf[c].x);
#pragma unroll
        for (int r = kk; r < 4; ++r) dmma884(acc[r][c][0], acc[r][c][1], af[r].y, bf[c].y);
      }
    }
    // C fragment (row g, columns 2l, 2l+1) -> packed A-fragment-major tile, negated
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double* prow = Pd + (tri(4 * bi + r) + 4 * bj) * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int q0 = 2 * l, q1 = 2 * l + 1;
        prow[c * 64 + (g * 4 + (q0 & 3)) * 2 + (q0 >> 2)] = -acc[r][c][0];
        prow[c * 64 + (g * 4 + (q1 & 3)) * 2 + (q1 >> 2)] = -acc[r][c][1];
      }
    }
  }
}

// ---- host side: the recursion plan of every shape present, grouped by height (InvPlanHost: pmk_internal.cuh) --------------
static int inv_split(int lo, int hi, std::vector<std::vector<InvNode>>& by_height) {
  const int size = hi - lo;
  if (size <= 1) return 0;
  int left = 1;
  while (left * 2 < size) left *= 2;          // largest power of two below the size
  const int mid = lo + left;
  const int h = 1 + std::max(inv_split(lo, mid, by_height), inv_split(mid, hi, by_height));
  if ((int)by_height.size() <= h) by_height.resize(h + 1);
  by_height[h].push_back(InvNode{(short)lo, (short)mid, (short)hi, 0});
  return h;
}

// shapes: number of 32-row blocks of every leaf in the range; returns the plan for the shapes present
InvPlanHost make_inverse_plan(const std::vector<int>& shapes_present) {
  InvPlanHost pl;
  pl.off.assign((kInvMaxBlocks + 1) * (kInvMaxHeight + 1), 0);
  pl.cnt.assign((kInvMaxBlocks + 1) * (kInvMaxHeight + 1), 0);
  for (int nb : shapes_present) {
    if (nb < 1 || nb > kInvMaxBlocks) continue;
    std::vector<std::vector<InvNode>> by_height;
    const int H = inv_split(0, nb, by_height);
    pl.max_height = std::max(pl.max_height, H);
    for (int h = 1; h <= H && h <= kInvMaxHeight; ++h) {
      const int idx = nb * (kInvMaxHeight + 1) + h;
      pl.off[idx] = (int)pl.nodes.size();
      pl.cnt[idx] = (int)by_height[h].size();
      pl.max_cnt[h] = std::max(pl.max_cnt[h], pl.cnt[idx]);
      for (const InvNode& nd : by_height[h]) pl.max_blocks[h] = std::max(pl.max_blocks[h], (nd.hi - nd.mid) * (nd.mid - nd.lo));
      pl.nodes.insert(pl.nodes.end(), by_height[h].begin(), by_height[h].end());
    }
  }
  return pl;
}

int inverse_plan_table_ints() { return (kInvMaxBlocks + 1) * (kInvMaxHeight + 1); }

// d_nodes / d_off / d_cnt: the plan uploaded by the caller.  Launches 1 + 2 * max_height kernels on the stream.
void launch_inverse(const LeafTable& lt, const InvPlanHost& plh, const void* d_nodes, const int* d_off, const int* d_cnt,
                    double* scratch, int first_leaf, int n_leaves, int max_npad, cudaStream_t s, int64_t* launches) {
  if (n_leaves <= 0) return;
  InvPlanDev pl{reinterpret_cast<const InvNode*>(d_nodes), d_off, d_cnt};
  k_inv_diag<<<dim3((max_npad / 32 + 7) / 8, n_leaves), 256, 0, s>>>(lt, first_leaf);
  ++*launches;
  for (int h = 1; h <= plh.max_height; ++h) {
    if (plh.max_cnt[h] == 0) continue;
    const int bpn = plh.max_blocks[h];
    const dim3 grid((plh.max_cnt[h] * bpn + kInvWarps - 1) / kInvWarps, n_leaves);
    k_inv_T<<<grid, kInvWarps * 32, 0, s>>>(lt, pl, scratch, first_leaf, h, bpn);
    k_inv_R<<<grid, kInvWarps * 32, 0, s>>>(lt, pl, scratch, first_leaf, h, bpn);
    *launches += 2;
  }
}

}  // namespace pmk
