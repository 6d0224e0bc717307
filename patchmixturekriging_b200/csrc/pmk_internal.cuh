// Internal device-side descriptors shared by the kernels and the C-ABI layer.
#pragma once
#include <vector>
#include "pmk_common.cuh"

namespace pmk {

// All leaves of one fitted model, resident in HBM.
//  xs    : training inputs, SoA, each leaf padded to npad (multiple of 32): xs[d*xstride + xoff[p] + i]
//  y     : targets, same padding;  alpha: GP weights (c_set), same padding, pad entries 0
//  L     : packed lower Cholesky factors (pmk_common.cuh layout), leaf p at L + loff[p];
//          rows/cols >= n[p] are identity padding
//  Linv  : inverses of the 32x32 diagonal blocks of L, leaf p block J at Linv + ioff[p] + J*640
//  M     : L with every strictly-lower 32x32 block right-multiplied by its column block's inverse (k_make_M)
struct LeafTable {
  int n_leaves;
  const int* n;
  const int* npad;
  const int64_t* xoff;
  const int64_t* loff;
  const int64_t* ioff;
  double* xs;
  int64_t xstride;
  double* y;
  double* alpha;
  double* L;
  double* M;          // same tile layout as L: M_IJ = L_IJ inv(L_JJ) for the strictly-lower 32x32 blocks (pair kernel operand)
  double* Linv;
  double* P;          // same tile layout as L: the full inverse inv(L) (operand of the explicit-inverse pair kernel)
  int* info;          // per leaf: 0 ok, >0 = order of the first non-positive leading minor
};

// Flattened complete BSP tree: hyperplanes of internal nodes in pre-order.
//  node k at depth d: left child k+1, right child k + 2^(L-1-d), L = levels-1; leaf id = 1 + path bits.
struct TreeDev {
  int levels;
  int n_hp;
  const double* hv;   // SoA: hv[d*n_hp + k]
  const double* hc;
};

struct QueryPlan {
  int64_t Nq;
  const double* Xq;       // D x Nq point-major (device)
  int32_t* home;          // Nq, 1-based
  int32_t* npairs;        // Nq
  int64_t* pair_off;      // Nq+1
  int64_t n_pairs;
  int32_t* pair_leaf;     // 1-based global leaf id
  int32_t* pair_q;        // query index
  int32_t* pair_hp;       // 1-based hyperplane index, 0 for the home slot
  double* pair_t;
  double* pair_w;
};

// Work description of one size class for the fused pair kernel (pmk_query.cu).
struct PairWork {
  const int* class_leaves;          // local (handle) leaf indices of this size class
  int n_class_leaves;
  const int64_t* tile_off;          // n_class_leaves + 1, exclusive scan of tiles per class leaf
  const int64_t* leaf_pair_start;   // total_leaves + 1, start of each GLOBAL leaf's run in sorted_pair
  const int32_t* sorted_pair;       // pair ids sorted (stably) by leaf
  int64_t leaf_base;                // global id (0-based) of local leaf 0
};

// The fit's batched Cholesky: leaves of at least this many padded rows are factored by the level-synchronous kernels
// (k_chol_factor64 / k_chol_panel64), smaller ones by the one-CTA-per-leaf kernel (k_chol).  Measured on a B200 (chol phase, ms):
// 512-point leaves (C3) 10.05 vs 10.33, one GPU's eighth of them 1.62 vs 1.56; 1027-point leaves (C4) 118.0 vs 144.6.
static constexpr int kCholLevelsMinNpad = 768;

// Recursion plan of the explicit inverse P = inv(L) by recursive doubling (pmk_invert.cu): for every shape (number of 32-row
// blocks) present and every height of the recursion tree, that shape's nodes of that height.
static constexpr int kInvMaxBlocks = PMK_MAX_LEAF_POINTS / 32;      // 64
static constexpr int kInvMaxHeight = 8;
struct InvNode {
  short lo, mid, hi, pad;       // block ranges: A = [lo, mid), B = [mid, hi), C = rows [mid, hi) x columns [lo, mid)
};
struct InvPlanHost {
  std::vector<InvNode> nodes;
  std::vector<int> off, cnt;    // [kInvMaxBlocks + 1][kInvMaxHeight + 1]
  int max_height = 0;
  int max_cnt[kInvMaxHeight + 1] = {};      // nodes of that height, largest over the shapes
  int max_blocks[kInvMaxHeight + 1] = {};   // 32x32 output blocks of a node of that height, largest over the shapes
};

}  // namespace pmk
