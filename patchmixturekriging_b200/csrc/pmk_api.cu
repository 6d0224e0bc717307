// C-ABI layer of libpmk_b200.so (see include/pmk.h).  Owns the device state of one fitted
// mixture-GP model, validates arguments, orders the kernels on the handle's stream and maps
// failures to status codes.  No CPU compute path exists here: every numerical result comes
// from the sm_100a kernels in pmk_fit.cu / pmk_tree.cu / pmk_query.cu / pmk_gram.cu.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include "pmk_internal.cuh"

namespace pmk {
// launchers defined in the kernel translation units
void launch_pack(int D, const LeafTable& lt, const int64_t* d_leaf_off, const double* dX, const double* dy, cudaStream_t s);
void launch_gram_tiles(int D, const LeafTable& lt, const int* d_order, int n_order, int max_npad, KParams kp, double sigma2,
                       cudaStream_t s);
void launch_chol(const LeafTable& lt, const int* d_order, int n_order, cudaStream_t s);
void launch_solve(const LeafTable& lt, const int* d_order, int n_order, int max_npad, cudaStream_t s, const double* rhs = nullptr,
                  double* out = nullptr, int backward_only = 0);
int launch_chol_levels(const LeafTable& lt, const int* d_order, const std::vector<int>& leaves_per_panel, int max_npad, int with_z,
                       cudaStream_t s);
void launch_make_M(const LeafTable& lt, int first_leaf, int n_leaves, int max_npad, cudaStream_t s);
InvPlanHost make_inverse_plan(const std::vector<int>& shapes_present);
void launch_inverse(const LeafTable& lt, const InvPlanHost& plh, const void* d_nodes, const int* d_off, const int* d_cnt,
                    double* scratch, int first_leaf, int n_leaves, int max_npad, cudaStream_t s, int64_t* launches);
void launch_unpack_L(const LeafTable& lt, int p, int n, double* d_out, cudaStream_t s, int which = 0);
void read_chol_cycles(unsigned long long* out, bool reset);
void read_query_cycles(int D, unsigned long long* out, bool reset);
void launch_home(int D, const TreeDev& tr, int64_t Nq, const double* dXq, int32_t* d_home, int32_t* d_leaf_qcount,
                 cudaStream_t s);
void launch_leaf_bbox(int D, int n_leaves, int64_t Nq, const double* dXq, const int32_t* qperm, const int64_t* leaf_qstart,
                      double* bbox, cudaStream_t s);
void launch_leaf_candidates(int D, bool fill, int n_leaves, const TreeDev& tr, const double* bbox, double radius,
                            int32_t* cand_count, const int64_t* cand_start, int32_t* cand, cudaStream_t s);
void launch_neighbours(int D, bool fill, bool pruned, int n_leaves, const TreeDev& tr, const QueryPlan& q, double radius,
                       double delta, int wkind, double wparam, int32_t* d_leaf_count, const int32_t* qperm,
                       const int64_t* leaf_qstart, const int64_t* cand_start, const int32_t* cand, uint16_t* kept_rec,
                       cudaStream_t s);
void launch_combine(int64_t Nq, const int64_t* pair_off, const double* pw, const double* pu, const double* pv, double* dYq,
                    double* dVq, int mean_only, cudaStream_t s);
void launch_scan_small(const int32_t* in, int64_t* out, int n, cudaStream_t s);
void launch_eps_partitions(int D, bool fill, const TreeDev& tr, int64_t N, const double* dX, double eps, int32_t* counts,
                           const int64_t* off, int32_t* pair_leaf, int32_t* pair_pt, int32_t* leaf_count, cudaStream_t s);
void launch_aos_to_soa(int D, const double* dX, int64_t n, int64_t stride, double* xs, cudaStream_t s);
void launch_gram(int D, const double* xr, int64_t xr_stride, int n, const double* xc, int64_t xc_stride, int m, KParams kp,
                 double sigma2, int symmetric, double* dK, cudaStream_t s, int fast_exp = 0);
void partition_sum_plan(int64_t n, int64_t base, std::vector<int64_t>& start, std::vector<int32_t>& len, std::vector<int32_t>& depth);
void launch_part_iota(int32_t* perm, int64_t N, cudaStream_t s);
void launch_part_one_based(const int32_t* perm, int64_t N, int32_t* out, cudaStream_t s);
void launch_part_block_sums(int D, const double* X, const int32_t* perm, const int64_t* blk_start, const int32_t* blk_len, int n_blk,
                            double* blk_sum, cudaStream_t s);
void launch_part_node_z(int D, const double* X, const int32_t* perm, const int64_t* seg_off, const int32_t* node_blk_off,
                        const int32_t* blk_depth, const double* blk_sum, int n_nodes, double* z_out, cudaStream_t s);
void launch_part_project(int D, const double* X, const int32_t* perm, const int64_t* seg_off, int n_nodes, const double* v, int64_t N,
                         double* f, int32_t* node_id, cudaStream_t s);
void launch_part_median(const double* f_sorted, const int64_t* seg_off, int n_nodes, double* c_out, cudaStream_t s);
void launch_part_flags(const double* f, const int32_t* node_id, const double* c, int64_t N, int32_t* flag, cudaStream_t s);
void launch_part_scatter(const int32_t* perm, const int32_t* node_id, const int32_t* flag, const int32_t* scan, const int64_t* seg_off,
                         int64_t N, int32_t* perm_out, cudaStream_t s);
void launch_part_child_offsets(const int32_t* scan, const int64_t* seg_off, int n_nodes, int64_t N, int64_t* child_off, cudaStream_t s);
void launch_pack_sorted_pairs(int D, const QueryPlan& q, const int32_t* sorted_pair, double* X_sorted, int32_t* leaf_sorted, cudaStream_t s);
void launch_unpack_sorted_pairs(int64_t n, const int32_t* sorted_pair, const double* us, const double* vs, double* pu, double* pv, cudaStream_t s);
void launch_run_starts(const int32_t* keys, int64_t R, int64_t TL, int64_t leaf_base, int64_t n_own, int64_t* start, int* foreign, cudaStream_t s);
void launch_diag_range(const LeafTable& lt, double* out2, cudaStream_t s);
void launch_alpha_residual(int D, const LeafTable& lt, const int* d_order, int n_order, KParams kp, double sigma2, double* r, cudaStream_t s);
void launch_alpha_add(const LeafTable& lt, const double* d, int64_t n, cudaStream_t s);
void launch_dense_debug(int D, const TreeDev& tr, const QueryPlan& q, double radius, double delta, uint8_t* keep, double* ts, double* zs,
                        cudaStream_t s);
double launch_dmma_peak(double* d_out, int iters, cudaStream_t s);
int query_class_of(int npad);
int query_class_mq(int cls);
void launch_query_pairs(int D, int cls, unsigned grid, const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp,
                        int mean_only, double* pu, double* pv, cudaStream_t s);
void launch_class_tiles(const PairWork& w, int mq, int32_t* tiles, cudaStream_t s);
bool launch_query_rowp(int D, int cls, const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp, int flags, int npmax,
                       double* pu, double* pv, cudaStream_t s);
}  // namespace pmk

using namespace pmk;

namespace {

struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

thread_local std::string g_create_error;

}  // namespace

static constexpr int kNumClasses = 5;   // query size classes (pmk_query.cu)

struct pmk_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;

  // model
  bool fitted = false;
  bool m_ready = false;    // M (operand of the substitution pair kernel) built for the current factors
  bool p_ready = false;    // P = inv(L) (operand of the explicit-inverse pair kernel) built for the current factors
  int solver = -1;         // PMK_OPT_QUERY_SOLVER: -1 = by conditioning (default); 0 = explicit inverse, row-panel product;
                           // 1 = blocked substitution (TRSM)
  int class_max_npad[5] = {};
  int inverse_builder = 0; // PMK_OPT_INVERSE_BUILDER: 0 = recursive doubling (pmk_invert.cu), 1 = substitution kernel on identity columns
  InvPlanHost inv_plan;    // recursion plan of the shapes of the current model
  DBuf d_inv_nodes, d_inv_off, d_inv_cnt;
  int D = 0;
  int64_t n_leaves = 0, total_leaves = 0;
  int64_t leaf_base = 0;                   // global 0-based id of this handle's leaf 0 (pmk_set_leaf_base: sub-tree ownership)
  int64_t total_opt = 0;                   // leaves of the whole model when this handle owns a part of it (0 = all of them)
  KParams kp{0, 1.0};
  double sigma2 = 0.0;
  int max_npad = 0;
  int64_t L_doubles = 0, Linv_doubles = 0, x_points = 0;
  std::vector<int> h_n, h_npad;
  std::vector<int64_t> h_xoff, h_loff, h_ioff;
  int64_t xstride = 0;
  DBuf d_n, d_npad, d_xoff, d_loff, d_ioff, d_xs, d_y, d_alpha, d_L, d_M, d_Linv, d_info, d_order, d_leafoff, d_Xin, d_yin;
  DBuf d_class_leaves[kNumClasses], d_class_tiles[kNumClasses], d_tile_off[kNumClasses];
  int n_class[kNumClasses] = {};
  // inversion work lists (leaves of the fit range, per size class): tile = MQ columns of inv(L)
  DBuf d_P, d_T, d_inv_leaves[kNumClasses], d_inv_tile_off[kNumClasses];
  int n_inv_class[kNumClasses] = {};
  int64_t inv_tiles[kNumClasses] = {};
  LeafTable lt{};

  // tree
  bool tree_set = false;
  int tree_D = 0, levels = 1, n_hp = 0;
  DBuf d_hv, d_hc;
  TreeDev tree{};

  // query
  DBuf d_Xq, d_home, d_npairs, d_pair_off, d_Yq, d_Vq;
  DBuf d_pair_leaf, d_pair_q, d_pair_hp, d_pair_t, d_pair_w, d_pair_u, d_pair_v, d_sorted_pair, d_keys_out, d_iota;
  DBuf d_leaf_count, d_leaf_pair_start, d_cub;
  DBuf d_leaf_qcount, d_leaf_qstart, d_qperm, d_qkeys, d_bbox, d_cand_count, d_cand_start, d_cand, d_kept;
  bool full_scan = false;   // PMK_OPT_FULL_HYPERPLANE_SCAN
  bool force_full_scan = false;   // the tree's normals are not unit vectors, or it has more hyperplanes than the pruned search's 16-bit slots
  // routed pairs (this handle as the OWNER of leaves other handles' queries touch): sort scratch, separate from its own plan
  DBuf r_keys, r_sorted, r_leaf_start, d_info2;
  // conditioning: (max diag(L) / min diag(L))^2 over the leaves of the last fit, a lower bound of cond(K + sigma2 I)
  DBuf d_diag_range;
  double cond_est = 0.0;
  int alpha_refine = -1;    // PMK_OPT_ALPHA_REFINE: -1 auto (flagged models), 0 never, 1 always
  int gram_fast_exp = 0;    // PMK_OPT_GRAM_FAST_EXP: table-driven exp in the standalone Gram kernel (squared exponential; <= 2 ulp)
  int chol_variant = -1;    // PMK_OPT_CHOL_VARIANT: -1 = by leaf size (default), 0 = level-synchronous kernels, 1 = one CTA per leaf
  // organizetrainingsets on the device (results of the last call)
  DBuf o_X, o_counts, o_off, o_pl, o_pp, o_sl, o_sp, o_lcount, o_lstart;
  int64_t o_N = 0, o_total = 0, o_leaves = 0;
  // setuppartition on the device, level by level (pmk_partition.cu): state between the calls
  DBuf pt_X, pt_perm[2], pt_seg[2], pt_f, pt_f_tmp, pt_f_sorted, pt_node, pt_node_tmp, pt_node_sorted, pt_flag, pt_scan;
  DBuf pt_blk_start, pt_blk_len, pt_blk_depth, pt_node_blk_off, pt_blk_sum, pt_z, pt_v, pt_c;
  int pt_D = 0, pt_levels = 0, pt_depth = -1, pt_cur = 0;
  bool pt_z_done = false;
  int64_t pt_N = 0;
  std::vector<int64_t> pt_hseg;   // host copy of the current level's segment offsets (nodes + 1)
  DBuf d_scratch;   // Gram scratch
  QueryPlan plan{};
  bool plan_valid = false;
  int last_flags = 0;
  double last_radius = 0.0, last_delta = 0.0;

  // timings
  cudaEvent_t ev[2 * PMK_T_COUNT] = {};
  bool ev_used[PMK_T_COUNT] = {};
  double ms[PMK_T_COUNT] = {};
};

namespace {

int fail(pmk_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CU(h, expr)                                                                                   \
  do {                                                                                                \
    cudaError_t e_ = (expr);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return fail(h, PMK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define KCHECK(h, what)                                                                               \
  do {                                                                                                \
    cudaError_t e_ = cudaGetLastError();                                                              \
    if (e_ != cudaSuccess) return fail(h, PMK_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e_)); \
    ++(h)->launches;                                                                                  \
  } while (0)

struct Timer {
  pmk_handle* h;
  int slot;
  Timer(pmk_handle* h_, int slot_) : h(h_), slot(slot_) {
    cudaEventRecord(h->ev[2 * slot], h->stream);
  }
  ~Timer() {
    cudaEventRecord(h->ev[2 * slot + 1], h->stream);
    h->ev_used[slot] = true;
  }
};

int parse_kernel(pmk_handle* h, int kernel_id, const double* kparams, int nparams, KParams* out) {
  if (kernel_id < PMK_KERNEL_SQEXP || kernel_id > PMK_KERNEL_RQ) return fail(h, PMK_ERR_UNSUPPORTED, "unknown kernel id %d", kernel_id);
  out->kind = kernel_id;
  out->p = (nparams >= 1 && kparams) ? kparams[0] : 1.0;
  return PMK_OK;
}

bool is_stationary_host(int kind) {
  return kind == PMK_KERNEL_SQEXP || kind == PMK_KERNEL_SPLINE34 || kind == PMK_KERNEL_SPLINE12 ||
         kind == PMK_KERNEL_SPLINE32 || kind == PMK_KERNEL_RQ;
}

int set_device(pmk_handle* h) {
  CU(h, cudaSetDevice(h->device));
  return PMK_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int pmk_version(void) { return 100; }

// host only (no CUDA call): the recursion plan of the explicit inverse for a leaf of n_blocks 32-row blocks
int pmk_inverse_plan(int n_blocks, int max_nodes, int16_t* nodes4, int* n_nodes) {
  if (n_blocks < 1 || n_blocks > kInvMaxBlocks || !n_nodes) return PMK_ERR_ARG;
  const InvPlanHost pl = make_inverse_plan(std::vector<int>{n_blocks});
  int k = 0;
  for (int h = 1; h <= pl.max_height; ++h) {
    const int idx = n_blocks * (kInvMaxHeight + 1) + h;
    for (int i = 0; i < pl.cnt[idx]; ++i, ++k) {
      if (nodes4 && k < max_nodes) {
        const InvNode& nd = pl.nodes[pl.off[idx] + i];
        nodes4[4 * k + 0] = nd.lo;
        nodes4[4 * k + 1] = nd.mid;
        nodes4[4 * k + 2] = nd.hi;
        nodes4[4 * k + 3] = (int16_t)h;
      }
    }
  }
  *n_nodes = k;
  return PMK_OK;
}

const char* pmk_last_error(const pmk_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int pmk_create(pmk_handle** out, int device) {
  if (!out) return fail(nullptr, PMK_ERR_ARG, "pmk_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, PMK_ERR_CUDA, "pmk_create: no CUDA device (%s); libpmk_b200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return fail(nullptr, PMK_ERR_ARG, "pmk_create: device %d out of range [0,%d)", device, ndev);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, PMK_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, PMK_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major < 10)
    return fail(nullptr, PMK_ERR_CUDA, "pmk_create: device %s is sm_%d%d; this library is built for sm_100a only", prop.name,
                prop.major, prop.minor);
  pmk_handle* h = new (std::nothrow) pmk_handle();
  if (!h) return fail(nullptr, PMK_ERR_CUDA, "out of host memory");
  h->device = device;
  if (const char* ev = getenv("PMK_CHOL_VARIANT")) h->chol_variant = std::max(-1, std::min(1, atoi(ev)));      // A/B timing of the two factorisations
  e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete h;
    return fail(nullptr, PMK_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  for (int i = 0; i < 2 * PMK_T_COUNT; ++i) cudaEventCreate(&h->ev[i]);
  if (h->d_info2.ensure(sizeof(int)) != cudaSuccess) {
    pmk_destroy(h);
    return fail(nullptr, PMK_ERR_CUDA, "cudaMalloc failed");
  }
  *out = h;
  return PMK_OK;
}

void pmk_destroy(pmk_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  DBuf* bufs[] = {&h->d_n, &h->d_npad, &h->d_xoff, &h->d_loff, &h->d_ioff, &h->d_xs, &h->d_y, &h->d_alpha, &h->d_L, &h->d_M, &h->d_Linv,
                  &h->d_info, &h->d_order, &h->d_leafoff, &h->d_Xin, &h->d_yin, &h->d_hv, &h->d_hc, &h->d_Xq, &h->d_home,
                  &h->d_npairs, &h->d_pair_off, &h->d_Yq, &h->d_Vq, &h->d_pair_leaf, &h->d_pair_q, &h->d_pair_hp, &h->d_pair_t,
                  &h->d_pair_w, &h->d_pair_u, &h->d_pair_v, &h->d_sorted_pair, &h->d_keys_out, &h->d_iota, &h->d_leaf_count,
                  &h->d_leaf_pair_start, &h->d_cub, &h->d_scratch, &h->d_leaf_qcount, &h->d_leaf_qstart, &h->d_qperm,
                  &h->d_qkeys, &h->d_bbox, &h->d_cand_count, &h->d_cand_start, &h->d_cand, &h->d_kept, &h->o_X, &h->o_counts, &h->o_off, &h->o_pl, &h->o_pp, &h->o_sl,
                  &h->o_sp, &h->o_lcount, &h->o_lstart};
  for (DBuf* b : bufs) b->release();
  for (int c = 0; c < kNumClasses; ++c) {
    h->d_class_leaves[c].release();
    h->d_class_tiles[c].release();
    h->d_tile_off[c].release();
    h->d_inv_leaves[c].release();
    h->d_inv_tile_off[c].release();
  }
  DBuf* pbufs[] = {&h->pt_X, &h->pt_perm[0], &h->pt_perm[1], &h->pt_seg[0], &h->pt_seg[1], &h->pt_f, &h->pt_f_tmp, &h->pt_f_sorted,
                   &h->pt_node, &h->pt_node_tmp, &h->pt_node_sorted, &h->pt_flag, &h->pt_scan, &h->pt_blk_start, &h->pt_blk_len,
                   &h->pt_blk_depth, &h->pt_node_blk_off, &h->pt_blk_sum, &h->pt_z, &h->pt_v, &h->pt_c};
  for (DBuf* b : pbufs) b->release();
  h->d_P.release();
  h->d_T.release();
  h->r_keys.release();
  h->r_sorted.release();
  h->r_leaf_start.release();
  h->d_info2.release();
  h->d_diag_range.release();
  h->d_inv_nodes.release();
  h->d_inv_off.release();
  h->d_inv_cnt.release();
  for (int i = 0; i < 2 * PMK_T_COUNT; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  cudaStreamDestroy(h->stream);
  delete h;
}

void* pmk_stream(pmk_handle* h) { return h ? (void*)h->stream : nullptr; }

int pmk_synchronize(pmk_handle* h) {
  if (!h) return PMK_ERR_ARG;
  CU(h, cudaStreamSynchronize(h->stream));
  return PMK_OK;
}

int64_t pmk_launch_count(const pmk_handle* h) { return h ? h->launches : 0; }

int pmk_set_option(pmk_handle* h, int option, int64_t value) {
  if (!h) return PMK_ERR_ARG;
  switch (option) {
    case PMK_OPT_FULL_HYPERPLANE_SCAN: h->full_scan = value != 0; h->plan_valid = false; return PMK_OK;
    case PMK_OPT_QUERY_SOLVER:
      if (value < -1 || value > 1)
        return fail(h, PMK_ERR_ARG, "PMK_OPT_QUERY_SOLVER: -1 (by conditioning), 0 (explicit inverse) or 1 (substitution)");
      h->solver = (int)value;
      return PMK_OK;
    case PMK_OPT_GRAM_FAST_EXP:
      h->gram_fast_exp = value != 0;
      return PMK_OK;
    case PMK_OPT_CHOL_VARIANT:
      if (value < -1 || value > 1) return fail(h, PMK_ERR_ARG, "PMK_OPT_CHOL_VARIANT: -1 (by leaf size), 0 (level-synchronous) or 1 (one CTA per leaf)");
      h->chol_variant = (int)value;
      return PMK_OK;
    case PMK_OPT_ALPHA_REFINE:
      if (value < -1 || value > 1) return fail(h, PMK_ERR_ARG, "PMK_OPT_ALPHA_REFINE: -1 (by conditioning), 0 (never) or 1 (always)");
      h->alpha_refine = (int)value;
      return PMK_OK;
    case PMK_OPT_INVERSE_BUILDER:
      if (value != 0 && value != 1) return fail(h, PMK_ERR_ARG, "PMK_OPT_INVERSE_BUILDER: 0 (recursive doubling) or 1 (substitution)");
      h->inverse_builder = (int)value;
      h->p_ready = false;
      return PMK_OK;
    default: return fail(h, PMK_ERR_ARG, "unknown option %d", option);
  }
}

int pmk_debug_counters(pmk_handle* h, uint64_t* out8, int reset) {
  if (!h || !out8) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  CU(h, cudaStreamSynchronize(h->stream));
  if (reset & 2) read_query_cycles(h->D, reinterpret_cast<unsigned long long*>(out8), (reset & 1) != 0);
  else read_chol_cycles(reinterpret_cast<unsigned long long*>(out8), (reset & 1) != 0);
  return PMK_OK;
}

int pmk_get_timings(pmk_handle* h, double* ms) {
  if (!h || !ms) return PMK_ERR_ARG;
  CU(h, cudaStreamSynchronize(h->stream));
  for (int s = 0; s < PMK_T_COUNT; ++s) {
    if (h->ev_used[s]) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, h->ev[2 * s], h->ev[2 * s + 1]) == cudaSuccess) h->ms[s] = t;
    }
    ms[s] = h->ms[s];
  }
  return PMK_OK;
}

// ---------------------------------------------------------------------------------------------
// Gram
static int gram_impl(pmk_handle* h, int D, int64_t n, const double* X, int64_t m, const double* Z, int kernel_id,
                     const double* kparams, int nparams, double sigma2, double* K_out) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (D < 1 || D > PMK_MAX_DIM) return fail(h, PMK_ERR_UNSUPPORTED, "D=%d unsupported (1..%d)", D, PMK_MAX_DIM);
  if (n < 0 || m < 0 || n > INT32_MAX || m > INT32_MAX) return fail(h, PMK_ERR_ARG, "bad sizes n=%lld m=%lld", (long long)n, (long long)m);
  if (n == 0 || m == 0) return PMK_OK;
  if (!X || !K_out) return fail(h, PMK_ERR_ARG, "NULL pointer");
  KParams kp;
  if (int rc = parse_kernel(h, kernel_id, kparams, nparams, &kp)) return rc;
  const bool sym = (Z == nullptr);
  const int64_t sx = (n + 127) / 128 * 128, sz = (m + 127) / 128 * 128;
  auto up16 = [](int64_t v) { return (v + 15) / 16 * 16; };   // keep every sub-buffer 128-byte aligned (TMA, 16-B stores)
  const int64_t nXa = up16((int64_t)D * n), nZa = sym ? 0 : up16((int64_t)D * m);
  const int64_t nxs = (int64_t)D * sx, nzs = sym ? 0 : (int64_t)D * sz;
  const size_t need = (size_t)(nXa + nZa + nxs + nzs + n * m) * sizeof(double);
  CU(h, h->d_scratch.ensure(need));
  double* dXa = h->d_scratch.as<double>();
  double* dZa = dXa + nXa;
  double* xs = dZa + nZa;
  double* zs = xs + nxs;
  double* dK = zs + nzs;
  CU(h, cudaMemcpyAsync(dXa, X, sizeof(double) * D * n, cudaMemcpyHostToDevice, h->stream));
  launch_aos_to_soa(D, dXa, n, sx, xs, h->stream);
  KCHECK(h, "k_aos_to_soa");
  if (!sym) {
    CU(h, cudaMemcpyAsync(dZa, Z, sizeof(double) * D * m, cudaMemcpyHostToDevice, h->stream));
    launch_aos_to_soa(D, dZa, m, sz, zs, h->stream);
    KCHECK(h, "k_aos_to_soa");
  }
  {
    Timer t(h, PMK_T_GRAM);
    launch_gram(D, xs, sx, (int)n, sym ? xs : zs, sym ? sx : sz, (int)m, kp, sigma2, sym ? 1 : 0, dK, h->stream, h->gram_fast_exp);
  }
  KCHECK(h, "k_gram");
  CU(h, cudaMemcpyAsync(K_out, dK, sizeof(double) * n * m, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return PMK_OK;
}

int pmk_gram(pmk_handle* h, int D, int64_t n, const double* X, int kernel_id, const double* kparams, int nparams,
             double sigma2, double* K_out) {
  return gram_impl(h, D, n, X, n, nullptr, kernel_id, kparams, nparams, sigma2, K_out);
}

int pmk_cross_gram(pmk_handle* h, int D, int64_t n, const double* X, int64_t m, const double* Z, int kernel_id,
                   const double* kparams, int nparams, double* K_out) {
  if (!Z && m > 0) return fail(h, PMK_ERR_ARG, "Z is NULL");
  return gram_impl(h, D, n, X, m, Z, kernel_id, kparams, nparams, 0.0, K_out);
}

// ---------------------------------------------------------------------------------------------
// fit
// PMK_OPT_QUERY_SOLVER = -1 (default): by the conditioning of the fitted leaves.  s = inv(L) kq through the explicit inverse
// carries an error of order cond(L) eps |inv(L)||kq| -- measured against the reference's dtrsv: 2e-11 of the variance at
// cond(K + sigma2 I) ~ 3e4, 3e-9 at 3e6 (profiles/parity_floor_r02.json) -- while blocked substitution stays at the level at
// which dtrsv and dtrsm differ from each other.  A model whose worst leaf has (max diag L / min diag L)^2 above the
// threshold is therefore queried by substitution; the estimate is a lower bound of cond(K + sigma2 I), a by-product of the fit.
static constexpr double kCondFlag = 1e4;
static constexpr int64_t kInvScratchDoubles = (int64_t)1 << 30;     // 8 GB of scratch for the recursive inverse (per chunk of leaves)
static int effective_solver(const pmk_handle* h) {
  if (h->solver >= 0) return h->solver;
  return h->cond_est >= kCondFlag ? 1 : 0;
}
// Does the variance query stream P = inv(L)?  Solver 0 (row-panel kernel), every kernel function.
static bool uses_inverse(const pmk_handle* h) { return effective_solver(h) == 0; }

// 0, 1, 2, ... on the device (pair / query ids for the radix sorts), grown on demand
static int ensure_iota(pmk_handle* h, int64_t n) {
  if (h->d_iota.cap < sizeof(int32_t) * (size_t)n) {
    CU(h, h->d_iota.ensure(sizeof(int32_t) * n));
    launch_part_iota(h->d_iota.as<int32_t>(), (int64_t)(h->d_iota.cap / sizeof(int32_t)), h->stream);
    KCHECK(h, "k_part_iota");
  }
  return PMK_OK;
}

// The pair kernels' operands: P = inv(L) (explicit-inverse solvers) and/or M_IJ = L_IJ inv(L_JJ) (substitution solver).
// P comes from recursive doubling on the packed tiles (pmk_invert.cu, with a bounded scratch buffer) or, with
// PMK_OPT_INVERSE_BUILDER = 1, from the substitution pair kernel run on identity right-hand sides (tile = MQ columns of the inverse).
static int build_operands(pmk_handle* h, bool want_M, bool want_P) {
  const int64_t f0 = 0, f1 = h->n_leaves;
  const bool all = true;
  auto make_M = [&]() -> int {
    Timer tm(h, PMK_T_Q_MAKE_M);
    CU(h, h->d_M.ensure(sizeof(double) * (size_t)h->L_doubles));
    h->lt.M = h->d_M.as<double>();
    launch_make_M(h->lt, (int)f0, (int)(f1 - f0), h->max_npad, h->stream);
    KCHECK(h, "k_make_M");
    h->m_ready = all;
    return PMK_OK;
  };
  if (want_P && !h->p_ready) {
    CU(h, h->d_P.ensure(sizeof(double) * (size_t)h->L_doubles));
    h->lt.P = h->d_P.as<double>();
    if (h->inverse_builder == 0) {
      Timer tm(h, PMK_T_Q_INVERT);
      // scratch for T = C inv(A), addressed like the factor (scratch + loff[p]): leaves go through in chunks of at most
      // kInvScratchDoubles so that a large model (C5: 68 GB of factors) does not need a third buffer of its size
      std::vector<std::pair<int64_t, int64_t>> chunks;       // [first leaf, count)
      int64_t need = 0;
      for (int64_t a = f0; a < f1;) {
        int64_t b = a;
        const int64_t base = h->h_loff[a];
        auto end_of = [&](int64_t p) { return p + 1 < h->n_leaves ? h->h_loff[p + 1] : h->L_doubles; };
        while (b < f1 && (b == a || end_of(b) - base <= kInvScratchDoubles)) ++b;
        need = std::max(need, end_of(b - 1) - base);
        chunks.emplace_back(a, b - a);
        a = b;
      }
      CU(h, h->d_T.ensure(sizeof(double) * (size_t)need));
      for (const auto& ck : chunks) {
        launch_inverse(h->lt, h->inv_plan, h->d_inv_nodes.p, h->d_inv_off.as<int>(), h->d_inv_cnt.as<int>(),
                       h->d_T.as<double>() - h->h_loff[ck.first], (int)ck.first, (int)ck.second, h->max_npad, h->stream, &h->launches);
      }
      --h->launches;
      KCHECK(h, "k_inv_* (recursive inverse)");
    } else {
      if (!h->m_ready)
        if (int rc = make_M()) return rc;
      Timer tm(h, PMK_T_Q_INVERT);
      for (int c = 0; c < kNumClasses; ++c) {
        if (h->n_inv_class[c] == 0) continue;
        PairWork w{};
        w.class_leaves = h->d_inv_leaves[c].as<int>();
        w.n_class_leaves = h->n_inv_class[c];
        w.tile_off = h->d_inv_tile_off[c].as<int64_t>();
        w.leaf_base = 0;
        QueryPlan q{};
        launch_query_pairs(h->D, c, (unsigned)h->inv_tiles[c], h->lt, w, q, h->kp, 4, nullptr, nullptr, h->stream);
        KCHECK(h, "k_query_pairs (inversion)");
      }
    }
    h->p_ready = all;
  }
  if (want_M && !h->m_ready)
    if (int rc = make_M()) return rc;
  return PMK_OK;
}

int pmk_build_M(pmk_handle* h) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (h->n_leaves == 0) return fail(h, PMK_ERR_STATE, "no model laid out (call pmk_fit first)");
  h->m_ready = false;
  h->p_ready = false;
  const bool use_P = uses_inverse(h);
  return build_operands(h, !use_P, use_P);
}

// Lays the model out in HBM (padded leaves, packed factor slots, size classes, inversion plan) and, with `compute`, packs the
// training sets and factorises the leaves of the fit range.  Without `compute` (pmk_load_model) the buffers are left for the
// caller to fill; dX / dy are not read.
static int fit_impl(pmk_handle* h, int D, int64_t n_leaves, const int64_t* leaf_off, const double* dX, const double* dy,
                    int kernel_id, const double* kparams, int nparams, double sigma2, int64_t* bad_leaf, int* info, bool compute) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  h->fitted = false;
  if (bad_leaf) *bad_leaf = 0;
  if (info) *info = 0;
  if (D < 1 || D > PMK_MAX_DIM) return fail(h, PMK_ERR_UNSUPPORTED, "D=%d unsupported (1..%d)", D, PMK_MAX_DIM);
  if (n_leaves < 1 || n_leaves > (1 << 24)) return fail(h, PMK_ERR_ARG, "n_leaves=%lld out of range", (long long)n_leaves);
  if (!leaf_off || (compute && (!dX || !dy))) return fail(h, PMK_ERR_ARG, "NULL pointer");
  KParams kp;
  if (int rc = parse_kernel(h, kernel_id, kparams, nparams, &kp)) return rc;
  if (leaf_off[0] != 0) return fail(h, PMK_ERR_ARG, "leaf_off[0] must be 0");
  // host-side layout of the padded leaves
  h->h_n.assign(n_leaves, 0);
  h->h_npad.assign(n_leaves, 0);
  h->h_xoff.assign(n_leaves, 0);
  h->h_loff.assign(n_leaves, 0);
  h->h_ioff.assign(n_leaves, 0);
  int64_t xo = 0, lo = 0, io = 0;
  int max_npad = 0;
  for (int64_t p = 0; p < n_leaves; ++p) {
    const int64_t np = leaf_off[p + 1] - leaf_off[p];
    if (np < 1) return fail(h, PMK_ERR_ARG, "leaf %lld is empty (the reference asserts !isempty(X), RKHS.jl:199)", (long long)(p + 1));
    if (np > PMK_MAX_LEAF_POINTS)
      return fail(h, PMK_ERR_UNSUPPORTED, "leaf %lld has %lld points (> %d)", (long long)(p + 1), (long long)np, PMK_MAX_LEAF_POINTS);
    const int npad = (int)((np + 31) / 32 * 32);
    h->h_n[p] = (int)np;
    h->h_npad[p] = npad;
    h->h_xoff[p] = xo;
    h->h_loff[p] = lo;
    h->h_ioff[p] = io;
    xo += npad;
    lo += (int64_t)ltile_doubles(npad);
    io += (int64_t)(npad / 32) * kInvDoublesPerBlock;
    max_npad = std::max(max_npad, npad);
  }
  const int64_t total_pts = leaf_off[n_leaves];
  h->xstride = xo + 256;   // slack: the Gram kernel's TMA tiles may read past the last leaf
  h->L_doubles = lo;
  h->Linv_doubles = io;
  h->x_points = xo;
  h->max_npad = max_npad;
  {   // recursion plan of the explicit inverse for the shapes (numbers of 32-row blocks) of this model
    std::vector<int> shapes;
    std::vector<char> seen(kInvMaxBlocks + 1, 0);
    for (int64_t p = 0; p < n_leaves; ++p) {
      const int nb = h->h_npad[p] / 32;
      if (nb >= 1 && nb <= kInvMaxBlocks && !seen[nb]) {
        seen[nb] = 1;
        shapes.push_back(nb);
      }
    }
    h->inv_plan = make_inverse_plan(shapes);
    CU(h, h->d_inv_nodes.ensure(sizeof(InvNode) * std::max<size_t>(1, h->inv_plan.nodes.size())));
    CU(h, h->d_inv_off.ensure(sizeof(int) * h->inv_plan.off.size()));
    CU(h, h->d_inv_cnt.ensure(sizeof(int) * h->inv_plan.cnt.size()));
    if (!h->inv_plan.nodes.empty())
      CU(h, cudaMemcpyAsync(h->d_inv_nodes.p, h->inv_plan.nodes.data(), sizeof(InvNode) * h->inv_plan.nodes.size(), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->d_inv_off.p, h->inv_plan.off.data(), sizeof(int) * h->inv_plan.off.size(), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->d_inv_cnt.p, h->inv_plan.cnt.data(), sizeof(int) * h->inv_plan.cnt.size(), cudaMemcpyHostToDevice, h->stream));
  }
  h->D = D;
  h->n_leaves = n_leaves;
  h->total_leaves = h->total_opt > 0 ? h->total_opt : n_leaves;
  if (h->total_opt > 0 && h->leaf_base + n_leaves > h->total_opt)
    return fail(h, PMK_ERR_ARG, "leaves [%lld, %lld) do not fit a model of %lld leaves (pmk_set_leaf_base)", (long long)h->leaf_base,
                (long long)(h->leaf_base + n_leaves), (long long)h->total_opt);
  h->kp = kp;
  h->sigma2 = sigma2;

  CU(h, h->d_n.ensure(sizeof(int) * n_leaves));
  CU(h, h->d_npad.ensure(sizeof(int) * n_leaves));
  CU(h, h->d_xoff.ensure(sizeof(int64_t) * n_leaves));
  CU(h, h->d_loff.ensure(sizeof(int64_t) * n_leaves));
  CU(h, h->d_ioff.ensure(sizeof(int64_t) * n_leaves));
  CU(h, h->d_info.ensure(sizeof(int) * n_leaves));
  CU(h, h->d_order.ensure(sizeof(int) * n_leaves));
  CU(h, h->d_leafoff.ensure(sizeof(int64_t) * (n_leaves + 1)));
  CU(h, h->d_xs.ensure(sizeof(double) * h->xstride * D));
  CU(h, h->d_y.ensure(sizeof(double) * h->xstride));
  CU(h, h->d_alpha.ensure(sizeof(double) * h->xstride));
  CU(h, h->d_L.ensure(sizeof(double) * (size_t)lo));
  CU(h, h->d_Linv.ensure(sizeof(double) * (size_t)io));
  (void)total_pts;

  // every leaf is factorised, largest first (LPT order)
  const int64_t f0 = 0, f1 = n_leaves;
  std::vector<int> order;
  for (int64_t p = f0; p < f1; ++p) order.push_back((int)p);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h->h_npad[a] > h->h_npad[b]; });
  const int n_order = (int)order.size();
  // query size classes
  std::vector<int> cls[kNumClasses];
  for (int c = 0; c < kNumClasses; ++c) h->class_max_npad[c] = 0;
  for (int64_t p = 0; p < n_leaves; ++p) {
    const int c = query_class_of(h->h_npad[p]);
    cls[c].push_back((int)p);
    h->class_max_npad[c] = std::max(h->class_max_npad[c], h->h_npad[p]);
  }
  for (int c = 0; c < kNumClasses; ++c) {
    h->n_class[c] = (int)cls[c].size();
    if (!cls[c].empty()) {
      CU(h, h->d_class_leaves[c].ensure(sizeof(int) * cls[c].size()));
      CU(h, h->d_class_tiles[c].ensure(sizeof(int32_t) * cls[c].size()));
      CU(h, h->d_tile_off[c].ensure(sizeof(int64_t) * (cls[c].size() + 1)));
      CU(h, cudaMemcpyAsync(h->d_class_leaves[c].p, cls[c].data(), sizeof(int) * cls[c].size(), cudaMemcpyHostToDevice, h->stream));
    }
  }
  std::vector<int> icls[kNumClasses];
  std::vector<int64_t> ioff_t[kNumClasses];
  for (int64_t p = f0; p < f1; ++p) icls[query_class_of(h->h_npad[p])].push_back((int)p);
  for (int c = 0; c < kNumClasses; ++c) {
    h->n_inv_class[c] = (int)icls[c].size();
    h->inv_tiles[c] = 0;
    if (icls[c].empty()) continue;
    const int mq = query_class_mq(c);
    ioff_t[c].push_back(0);
    for (int p : icls[c]) ioff_t[c].push_back(ioff_t[c].back() + (h->h_npad[p] + mq - 1) / mq);
    h->inv_tiles[c] = ioff_t[c].back();
    CU(h, h->d_inv_leaves[c].ensure(sizeof(int) * icls[c].size()));
    CU(h, h->d_inv_tile_off[c].ensure(sizeof(int64_t) * ioff_t[c].size()));
    CU(h, cudaMemcpyAsync(h->d_inv_leaves[c].p, icls[c].data(), sizeof(int) * icls[c].size(), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->d_inv_tile_off[c].p, ioff_t[c].data(), sizeof(int64_t) * ioff_t[c].size(), cudaMemcpyHostToDevice, h->stream));
  }
  CU(h, cudaMemcpyAsync(h->d_n.p, h->h_n.data(), sizeof(int) * n_leaves, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_npad.p, h->h_npad.data(), sizeof(int) * n_leaves, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_xoff.p, h->h_xoff.data(), sizeof(int64_t) * n_leaves, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_loff.p, h->h_loff.data(), sizeof(int64_t) * n_leaves, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_ioff.p, h->h_ioff.data(), sizeof(int64_t) * n_leaves, cudaMemcpyHostToDevice, h->stream));
  if (n_order) CU(h, cudaMemcpyAsync(h->d_order.p, order.data(), sizeof(int) * n_order, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_leafoff.p, leaf_off, sizeof(int64_t) * (n_leaves + 1), cudaMemcpyHostToDevice, h->stream));
  // the host vectors above are pageable: the async copies have completed staging on return, but
  // `order`/`cls` die at scope exit, so drain the stream before that.
  CU(h, cudaStreamSynchronize(h->stream));

  LeafTable& lt = h->lt;
  lt.n_leaves = (int)n_leaves;
  lt.n = h->d_n.as<int>();
  lt.npad = h->d_npad.as<int>();
  lt.xoff = h->d_xoff.as<int64_t>();
  lt.loff = h->d_loff.as<int64_t>();
  lt.ioff = h->d_ioff.as<int64_t>();
  lt.xs = h->d_xs.as<double>();
  lt.xstride = h->xstride;
  lt.y = h->d_y.as<double>();
  lt.alpha = h->d_alpha.as<double>();
  lt.L = h->d_L.as<double>();
  lt.M = nullptr;          // allocated and built on the first variance query after a fit (or by pmk_build_M)
  lt.P = nullptr;
  lt.Linv = h->d_Linv.as<double>();
  lt.info = h->d_info.as<int>();
  h->m_ready = false;
  h->p_ready = false;
  h->plan_valid = false;
  if (!compute) return PMK_OK;

  {
    Timer t(h, PMK_T_FIT_PACK);
    launch_pack(D, lt, h->d_leafoff.as<int64_t>(), dX, dy, h->stream);
  }
  KCHECK(h, "k_pack_leaves");
  {
    Timer t(h, PMK_T_FIT_GRAM);
    launch_gram_tiles(D, lt, h->d_order.as<int>(), n_order, max_npad, kp, sigma2, h->stream);
  }
  KCHECK(h, "k_gram_tiles");
  {
    Timer t(h, PMK_T_FIT_CHOL);
    // `order` is sorted by size: the leaves of n_pad >= kCholLevelsMinNpad are a prefix.  They advance panel by panel through the
    // level-synchronous kernels; the smaller ones take the one-CTA-per-leaf kernel.  The choice is a property of the leaf, so a
    // leaf's factor has the same bits whatever else is in the batch (pmk_multi: any rank count).
    int n_levels = 0;
    if (h->chol_variant == 0) n_levels = n_order;
    else if (h->chol_variant < 0)
      while (n_levels < n_order && h->h_npad[order[n_levels]] >= kCholLevelsMinNpad) ++n_levels;
    if (n_levels > 0) {
      // leaves of the prefix that still have columns 64 Jp ..: again a prefix
      std::vector<int> per_panel((max_npad + 63) / 64, 0);
      for (int k = 0; k < n_levels; ++k)
        for (int Jp = 0; Jp < (h->h_npad[order[k]] + 63) / 64; ++Jp) ++per_panel[Jp];
      h->launches += launch_chol_levels(lt, h->d_order.as<int>(), per_panel, max_npad, 1, h->stream) - 1;
    }
    if (n_levels < n_order) {
      launch_chol(lt, h->d_order.as<int>() + n_levels, n_order - n_levels, h->stream);
      if (n_levels > 0) ++h->launches;
    }
  }
  KCHECK(h, "k_chol");
  {
    Timer t(h, PMK_T_FIT_SOLVE);
    // both factorisations leave z = L^-1 y in lt.alpha (formed panel by panel on the tiles they stream anyway): only the backward sweep is left
    launch_solve(lt, h->d_order.as<int>(), n_order, max_npad, h->stream, lt.alpha, nullptr, 1);
  }
  KCHECK(h, "k_solve_alpha");
  // conditioning estimate (by-product of the factor) and status: first failing leaf
  CU(h, h->d_diag_range.ensure(sizeof(double)));
  CU(h, cudaMemsetAsync(h->d_diag_range.p, 0, sizeof(double), h->stream));
  launch_diag_range(lt, h->d_diag_range.as<double>(), h->stream);
  KCHECK(h, "k_diag_range");
  std::vector<int> h_info(n_leaves);
  double cond_est = 0.0;
  CU(h, cudaMemcpyAsync(h_info.data(), h->d_info.p, sizeof(int) * n_leaves, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemcpyAsync(&cond_est, h->d_diag_range.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  h->cond_est = cond_est;
  for (int64_t p = 0; p < n_leaves; ++p) {
    if (h_info[p] != 0) {
      if (bad_leaf) *bad_leaf = h->leaf_base + p + 1;
      if (info) *info = h_info[p];
      return fail(h, PMK_ERR_NOT_POSDEF, "leaf %lld: K + sigma2*I is not positive definite (info=%d)",
                  (long long)(h->leaf_base + p + 1), h_info[p]);
    }
  }
  if (h->alpha_refine == 1 || (h->alpha_refine < 0 && cond_est >= kCondFlag)) {
    // alpha <- alpha + (L L^T)^-1 (y - (K + sigma2 I) alpha): the reference's c = U\y is an LU solve (mixtureGP.jl:106); one step
    // of refinement in working precision makes the Cholesky solution componentwise backward-stable like it
    Timer t(h, PMK_T_FIT_REFINE);
    CU(h, h->d_scratch.ensure(sizeof(double) * 2 * (size_t)h->xstride));
    double* r = h->d_scratch.as<double>();
    double* d = r + h->xstride;
    launch_alpha_residual(D, lt, h->d_order.as<int>(), n_order, kp, sigma2, r, h->stream);
    KCHECK(h, "k_alpha_residual");
    launch_solve(lt, h->d_order.as<int>(), n_order, max_npad, h->stream, r, d);
    KCHECK(h, "k_solve_alpha");
    launch_alpha_add(lt, d, h->x_points, h->stream);
    KCHECK(h, "k_alpha_add");
  }
  h->m_ready = false;
  h->p_ready = false;
  h->fitted = true;
  h->plan_valid = false;
  return PMK_OK;
}

int pmk_fit_dev(pmk_handle* h, int D, int64_t n_leaves, const int64_t* leaf_off, const double* dX, const double* dy,
                int kernel_id, const double* kparams, int nparams, double sigma2, int64_t* bad_leaf, int* info) {
  return fit_impl(h, D, n_leaves, leaf_off, dX, dy, kernel_id, kparams, nparams, sigma2, bad_leaf, info, true);
}

int pmk_fit(pmk_handle* h, int D, int64_t n_leaves, const int64_t* leaf_off, const double* X, const double* y, int kernel_id,
            const double* kparams, int nparams, double sigma2, int64_t* bad_leaf, int* info) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!leaf_off || !X || !y || n_leaves < 1) return fail(h, PMK_ERR_ARG, "NULL pointer or n_leaves < 1");
  if (D < 1 || D > PMK_MAX_DIM) return fail(h, PMK_ERR_UNSUPPORTED, "D=%d unsupported (1..%d)", D, PMK_MAX_DIM);
  const int64_t total = leaf_off[n_leaves];
  if (total < 1) return fail(h, PMK_ERR_ARG, "no training points");
  CU(h, h->d_Xin.ensure(sizeof(double) * total * D));
  CU(h, h->d_yin.ensure(sizeof(double) * total));
  CU(h, cudaMemcpyAsync(h->d_Xin.p, X, sizeof(double) * total * D, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_yin.p, y, sizeof(double) * total, cudaMemcpyHostToDevice, h->stream));
  return pmk_fit_dev(h, D, n_leaves, leaf_off, h->d_Xin.as<double>(), h->d_yin.as<double>(), kernel_id, kparams, nparams, sigma2,
                     bad_leaf, info);
}

// ---------------------------------------------------------------------------------------------
// setuppartition on the device (partition.jl:106-217), one level per call pair; see pmk.h and pmk_partition.cu
int pmk_partition_sum_plan(int64_t n, int64_t max_blocks, int64_t* blk_start, int32_t* blk_len, int32_t* blk_depth, int64_t* n_blocks) {
  if (n < 1 || !n_blocks) return PMK_ERR_ARG;
  std::vector<int64_t> st;
  std::vector<int32_t> ln, dp;
  partition_sum_plan(n, 0, st, ln, dp);
  *n_blocks = (int64_t)st.size();
  if ((int64_t)st.size() > max_blocks) return PMK_ERR_ARG;
  for (size_t i = 0; i < st.size(); ++i) {
    if (blk_start) blk_start[i] = st[i];
    if (blk_len) blk_len[i] = ln[i];
    if (blk_depth) blk_depth[i] = dp[i];
  }
  return PMK_OK;
}

int pmk_partition_begin(pmk_handle* h, int D, int64_t N, const double* X, int levels) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  h->pt_depth = -1;
  if (D < 1 || D > PMK_MAX_DIM) return fail(h, PMK_ERR_UNSUPPORTED, "D=%d unsupported (1..%d)", D, PMK_MAX_DIM);
  if (levels < 2 || levels > 25) return fail(h, PMK_ERR_ARG, "levels=%d out of range [2,25] (examples/mixGP.jl:108)", levels);
  if (!X || N < 1 || N > INT32_MAX - 1) return fail(h, PMK_ERR_ARG, "bad N or NULL X");
  if (N < ((int64_t)1 << (levels - 1))) return fail(h, PMK_ERR_ARG, "fewer points (%lld) than leaves (2^%d)", (long long)N, levels - 1);
  const int64_t max_nodes = (int64_t)1 << (levels - 1);
  CU(h, h->pt_X.ensure(sizeof(double) * N * D));
  for (int k = 0; k < 2; ++k) {
    CU(h, h->pt_perm[k].ensure(sizeof(int32_t) * N));
    CU(h, h->pt_seg[k].ensure(sizeof(int64_t) * (max_nodes + 1)));
  }
  CU(h, h->pt_f.ensure(sizeof(double) * N));
  CU(h, h->pt_f_tmp.ensure(sizeof(double) * N));
  CU(h, h->pt_f_sorted.ensure(sizeof(double) * N));
  CU(h, h->pt_node.ensure(sizeof(int32_t) * N));
  CU(h, h->pt_node_tmp.ensure(sizeof(int32_t) * N));
  CU(h, h->pt_node_sorted.ensure(sizeof(int32_t) * N));
  CU(h, h->pt_flag.ensure(sizeof(int32_t) * (N + 1)));
  CU(h, h->pt_scan.ensure(sizeof(int32_t) * (N + 1)));
  CU(h, h->pt_z.ensure(sizeof(double) * max_nodes * D));
  CU(h, h->pt_v.ensure(sizeof(double) * max_nodes * D));
  CU(h, h->pt_c.ensure(sizeof(double) * max_nodes));
  CU(h, cudaMemcpyAsync(h->pt_X.p, X, sizeof(double) * N * D, cudaMemcpyHostToDevice, h->stream));
  launch_part_iota(h->pt_perm[0].as<int32_t>(), N, h->stream);
  KCHECK(h, "k_part_iota");
  h->pt_hseg.assign({0, N});
  CU(h, cudaMemcpyAsync(h->pt_seg[0].p, h->pt_hseg.data(), sizeof(int64_t) * 2, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  h->pt_D = D;
  h->pt_N = N;
  h->pt_levels = levels;
  h->pt_cur = 0;
  h->pt_depth = 0;
  h->pt_z_done = false;
  return PMK_OK;
}

int pmk_partition_level_z(pmk_handle* h, int depth, double* z_out) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (h->pt_depth < 0) return fail(h, PMK_ERR_STATE, "pmk_partition_begin has not been called");
  if (depth != h->pt_depth || depth >= h->pt_levels - 1)
    return fail(h, PMK_ERR_STATE, "pmk_partition_level_z: depth %d requested, next depth is %d of %d", depth, h->pt_depth, h->pt_levels - 1);
  if (!z_out) return fail(h, PMK_ERR_ARG, "NULL pointer");
  const int nodes = 1 << depth;
  const int D = h->pt_D;
  std::vector<int64_t> st;
  std::vector<int32_t> ln, dp, nbo(nodes + 1, 0);
  for (int j = 0; j < nodes; ++j) {
    partition_sum_plan(h->pt_hseg[j + 1] - h->pt_hseg[j], h->pt_hseg[j], st, ln, dp);
    nbo[j + 1] = (int32_t)st.size();
  }
  const int n_blk = (int)st.size();
  CU(h, h->pt_blk_start.ensure(sizeof(int64_t) * n_blk));
  CU(h, h->pt_blk_len.ensure(sizeof(int32_t) * n_blk));
  CU(h, h->pt_blk_depth.ensure(sizeof(int32_t) * n_blk));
  CU(h, h->pt_node_blk_off.ensure(sizeof(int32_t) * (nodes + 1)));
  CU(h, h->pt_blk_sum.ensure(sizeof(double) * (size_t)n_blk * D));
  CU(h, cudaMemcpyAsync(h->pt_blk_start.p, st.data(), sizeof(int64_t) * n_blk, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->pt_blk_len.p, ln.data(), sizeof(int32_t) * n_blk, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->pt_blk_depth.p, dp.data(), sizeof(int32_t) * n_blk, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->pt_node_blk_off.p, nbo.data(), sizeof(int32_t) * (nodes + 1), cudaMemcpyHostToDevice, h->stream));
  const int32_t* perm = h->pt_perm[h->pt_cur].as<int32_t>();
  const int64_t* seg = h->pt_seg[h->pt_cur].as<int64_t>();
  launch_part_block_sums(D, h->pt_X.as<double>(), perm, h->pt_blk_start.as<int64_t>(), h->pt_blk_len.as<int32_t>(), n_blk,
                         h->pt_blk_sum.as<double>(), h->stream);
  KCHECK(h, "k_part_block_sums");
  launch_part_node_z(D, h->pt_X.as<double>(), perm, seg, h->pt_node_blk_off.as<int32_t>(), h->pt_blk_depth.as<int32_t>(),
                     h->pt_blk_sum.as<double>(), nodes, h->pt_z.as<double>(), h->stream);
  KCHECK(h, "k_part_node_z");
  CU(h, cudaMemcpyAsync(z_out, h->pt_z.p, sizeof(double) * (size_t)nodes * D, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));     // also keeps the pageable plan vectors alive until the copies are staged
  h->pt_z_done = true;
  return PMK_OK;
}

int pmk_partition_level_split(pmk_handle* h, int depth, const double* v, double* c_out) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (h->pt_depth < 0) return fail(h, PMK_ERR_STATE, "pmk_partition_begin has not been called");
  if (depth != h->pt_depth || !h->pt_z_done)
    return fail(h, PMK_ERR_STATE, "pmk_partition_level_split: call pmk_partition_level_z for depth %d first", h->pt_depth);
  if (!v || !c_out) return fail(h, PMK_ERR_ARG, "NULL pointer");
  const int nodes = 1 << depth;
  const int D = h->pt_D;
  const int64_t N = h->pt_N;
  cudaStream_t s = h->stream;
  const int cur = h->pt_cur, nxt = cur ^ 1;
  const int32_t* perm = h->pt_perm[cur].as<int32_t>();
  const int64_t* seg = h->pt_seg[cur].as<int64_t>();
  CU(h, cudaMemcpyAsync(h->pt_v.p, v, sizeof(double) * (size_t)nodes * D, cudaMemcpyHostToDevice, s));
  launch_part_project(D, h->pt_X.as<double>(), perm, seg, nodes, h->pt_v.as<double>(), N, h->pt_f.as<double>(),
                      h->pt_node.as<int32_t>(), s);
  KCHECK(h, "k_part_project");
  // f ascending inside every node: sort by f, then stably by node id
  {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs((void*)nullptr, tb, h->pt_f.as<double>(), h->pt_f_tmp.as<double>(), h->pt_node.as<int32_t>(),
                                    h->pt_node_tmp.as<int32_t>(), (int)N, 0, 64, s);
    CU(h, h->d_cub.ensure(tb));
    CU(h, cub::DeviceRadixSort::SortPairs(h->d_cub.p, tb, h->pt_f.as<double>(), h->pt_f_tmp.as<double>(), h->pt_node.as<int32_t>(),
                                          h->pt_node_tmp.as<int32_t>(), (int)N, 0, 64, s));
    ++h->launches;
  }
  const double* f_sorted = h->pt_f_tmp.as<double>();
  if (depth > 0) {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs((void*)nullptr, tb, h->pt_node_tmp.as<int32_t>(), h->pt_node_sorted.as<int32_t>(),
                                    h->pt_f_tmp.as<double>(), h->pt_f_sorted.as<double>(), (int)N, 0, depth, s);
    CU(h, h->d_cub.ensure(tb));
    CU(h, cub::DeviceRadixSort::SortPairs(h->d_cub.p, tb, h->pt_node_tmp.as<int32_t>(), h->pt_node_sorted.as<int32_t>(),
                                          h->pt_f_tmp.as<double>(), h->pt_f_sorted.as<double>(), (int)N, 0, depth, s));
    ++h->launches;
    f_sorted = h->pt_f_sorted.as<double>();
  }
  launch_part_median(f_sorted, seg, nodes, h->pt_c.as<double>(), s);
  KCHECK(h, "k_part_median");
  launch_part_flags(h->pt_f.as<double>(), h->pt_node.as<int32_t>(), h->pt_c.as<double>(), N, h->pt_flag.as<int32_t>(), s);
  KCHECK(h, "k_part_flags");
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum((void*)nullptr, tb, h->pt_flag.as<int32_t>(), h->pt_scan.as<int32_t>(), (int)(N + 1), s);
    CU(h, h->d_cub.ensure(tb));
    CU(h, cub::DeviceScan::ExclusiveSum(h->d_cub.p, tb, h->pt_flag.as<int32_t>(), h->pt_scan.as<int32_t>(), (int)(N + 1), s));
    ++h->launches;
  }
  launch_part_scatter(perm, h->pt_node.as<int32_t>(), h->pt_flag.as<int32_t>(), h->pt_scan.as<int32_t>(), seg, N,
                      h->pt_perm[nxt].as<int32_t>(), s);
  KCHECK(h, "k_part_scatter");
  launch_part_child_offsets(h->pt_scan.as<int32_t>(), seg, nodes, N, h->pt_seg[nxt].as<int64_t>(), s);
  KCHECK(h, "k_part_child_offsets");
  std::vector<int64_t> child(2 * (size_t)nodes + 1);
  CU(h, cudaMemcpyAsync(child.data(), h->pt_seg[nxt].p, sizeof(int64_t) * child.size(), cudaMemcpyDeviceToHost, s));
  CU(h, cudaMemcpyAsync(c_out, h->pt_c.p, sizeof(double) * nodes, cudaMemcpyDeviceToHost, s));
  CU(h, cudaStreamSynchronize(s));
  for (size_t k = 0; k + 1 < child.size(); ++k)
    if (child[k + 1] <= child[k]) {
      h->pt_depth = -1;
      return fail(h, PMK_ERR_ARG, "the split at depth %d leaves child %lld without points (the reference fails in mean() of an empty set, "
                  "partition.jl:89)", depth, (long long)k);
    }
  h->pt_hseg.swap(child);
  h->pt_cur = nxt;
  h->pt_depth = depth + 1;
  h->pt_z_done = false;
  return PMK_OK;
}

int pmk_partition_fetch(pmk_handle* h, int64_t* leaf_off_out, int32_t* inds_out) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (h->pt_depth < 0 || h->pt_depth != h->pt_levels - 1)
    return fail(h, PMK_ERR_STATE, "pmk_partition_fetch: the tree is not complete (depth %d of %d)", h->pt_depth, h->pt_levels - 1);
  if (leaf_off_out) memcpy(leaf_off_out, h->pt_hseg.data(), sizeof(int64_t) * h->pt_hseg.size());
  if (inds_out) {
    int32_t* one = h->pt_perm[h->pt_cur ^ 1].as<int32_t>();
    launch_part_one_based(h->pt_perm[h->pt_cur].as<int32_t>(), h->pt_N, one, h->stream);
    KCHECK(h, "k_part_one_based");
    CU(h, cudaMemcpyAsync(inds_out, one, sizeof(int32_t) * h->pt_N, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  return PMK_OK;
}

// ---------------------------------------------------------------------------------------------
// checkpoint: the fitted model as one flat file (the reference keeps MixtureGPType in memory only)
namespace {
struct ModelFileHeader {
  char magic[8];            // "PMKB200\0"
  int32_t version, D, kernel_kind, levels, tree_D, n_hp, tree_set, reserved;
  int64_t n_leaves, xstride, x_points, L_doubles, Linv_doubles;
  double kernel_p, sigma2;
};
const char kModelMagic[8] = {'P', 'M', 'K', 'B', '2', '0', '0', 0};

// device span -> file through a bounded pinned-size staging vector (the factors are GBs)
int dump_dev(pmk_handle* h, FILE* f, const void* dptr, size_t bytes, std::vector<char>& stage) {
  const char* src = static_cast<const char*>(dptr);
  for (size_t done = 0; done < bytes;) {
    const size_t n = std::min(stage.size(), bytes - done);
    CU(h, cudaMemcpyAsync(stage.data(), src + done, n, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (fwrite(stage.data(), 1, n, f) != n) return fail(h, PMK_ERR_ARG, "pmk_save_model: short write");
    done += n;
  }
  return PMK_OK;
}
int fill_dev(pmk_handle* h, FILE* f, void* dptr, size_t bytes, std::vector<char>& stage) {
  char* dst = static_cast<char*>(dptr);
  for (size_t done = 0; done < bytes;) {
    const size_t n = std::min(stage.size(), bytes - done);
    if (fread(stage.data(), 1, n, f) != n) return fail(h, PMK_ERR_ARG, "pmk_load_model: file truncated");
    CU(h, cudaMemcpyAsync(dst + done, stage.data(), n, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    done += n;
  }
  return PMK_OK;
}
struct FileCloser {
  FILE* f;
  ~FileCloser() { if (f) fclose(f); }
};
}  // namespace

int pmk_save_model(pmk_handle* h, const char* path) {
  if (!h || !path) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->fitted) return fail(h, PMK_ERR_STATE, "pmk_save_model: model is not fitted");
  if (h->leaf_base != 0 || h->n_leaves != h->total_leaves)
    return fail(h, PMK_ERR_STATE, "pmk_save_model: this handle owns leaves [%lld, %lld) of a %lld-leaf model; a model file holds a whole model",
                (long long)h->leaf_base, (long long)(h->leaf_base + h->n_leaves), (long long)h->total_leaves);
  FileCloser fc{fopen(path, "wb")};
  if (!fc.f) return fail(h, PMK_ERR_ARG, "pmk_save_model: cannot open %s for writing", path);
  ModelFileHeader hd{};
  memcpy(hd.magic, kModelMagic, 8);
  hd.version = 1;
  hd.D = h->D;
  hd.kernel_kind = h->kp.kind;
  hd.kernel_p = h->kp.p;
  hd.sigma2 = h->sigma2;
  hd.tree_set = h->tree_set ? 1 : 0;
  hd.levels = h->tree_set ? h->levels : 1;
  hd.tree_D = h->tree_set ? h->tree_D : 0;
  hd.n_hp = h->tree_set ? h->n_hp : 0;
  hd.n_leaves = h->n_leaves;
  hd.xstride = h->xstride;
  hd.x_points = h->x_points;
  hd.L_doubles = h->L_doubles;
  hd.Linv_doubles = h->Linv_doubles;
  if (fwrite(&hd, sizeof hd, 1, fc.f) != 1) return fail(h, PMK_ERR_ARG, "pmk_save_model: short write");
  std::vector<int64_t> off(h->n_leaves + 1, 0);
  for (int64_t p = 0; p < h->n_leaves; ++p) off[p + 1] = off[p] + h->h_n[p];
  if (fwrite(off.data(), sizeof(int64_t), off.size(), fc.f) != off.size()) return fail(h, PMK_ERR_ARG, "pmk_save_model: short write");
  std::vector<char> stage((size_t)64 << 20);
  if (int rc = dump_dev(h, fc.f, h->d_xs.p, sizeof(double) * h->xstride * h->D, stage)) return rc;
  if (int rc = dump_dev(h, fc.f, h->d_y.p, sizeof(double) * h->xstride, stage)) return rc;
  if (int rc = dump_dev(h, fc.f, h->d_alpha.p, sizeof(double) * h->xstride, stage)) return rc;
  if (int rc = dump_dev(h, fc.f, h->d_L.p, sizeof(double) * (size_t)h->L_doubles, stage)) return rc;
  if (int rc = dump_dev(h, fc.f, h->d_Linv.p, sizeof(double) * (size_t)h->Linv_doubles, stage)) return rc;
  if (hd.n_hp > 0) {
    if (int rc = dump_dev(h, fc.f, h->d_hv.p, sizeof(double) * (size_t)hd.tree_D * hd.n_hp, stage)) return rc;
    if (int rc = dump_dev(h, fc.f, h->d_hc.p, sizeof(double) * (size_t)hd.n_hp, stage)) return rc;
  }
  FILE* f = fc.f;
  fc.f = nullptr;
  if (fclose(f) != 0) return fail(h, PMK_ERR_ARG, "pmk_save_model: close of %s failed", path);
  return PMK_OK;
}

int pmk_load_model(pmk_handle* h, const char* path) {
  if (!h || !path) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  FileCloser fc{fopen(path, "rb")};
  if (!fc.f) return fail(h, PMK_ERR_ARG, "pmk_load_model: cannot open %s", path);
  ModelFileHeader hd{};
  if (fread(&hd, sizeof hd, 1, fc.f) != 1 || memcmp(hd.magic, kModelMagic, 8) != 0)
    return fail(h, PMK_ERR_ARG, "pmk_load_model: %s is not a libpmk_b200 model file", path);
  if (hd.version != 1) return fail(h, PMK_ERR_UNSUPPORTED, "pmk_load_model: file version %d", hd.version);
  if (hd.n_leaves < 1 || hd.n_leaves > (1 << 24)) return fail(h, PMK_ERR_ARG, "pmk_load_model: corrupt header");
  std::vector<int64_t> off(hd.n_leaves + 1);
  if (fread(off.data(), sizeof(int64_t), off.size(), fc.f) != off.size()) return fail(h, PMK_ERR_ARG, "pmk_load_model: file truncated");
  h->leaf_base = 0;     // a model file holds a whole model
  h->total_opt = 0;
  const double kparam = hd.kernel_p;
  if (int rc = fit_impl(h, hd.D, hd.n_leaves, off.data(), nullptr, nullptr, hd.kernel_kind, &kparam, 1, hd.sigma2, nullptr, nullptr, false))
    return rc;
  if (h->xstride != hd.xstride || h->x_points != hd.x_points || h->L_doubles != hd.L_doubles || h->Linv_doubles != hd.Linv_doubles)
    return fail(h, PMK_ERR_UNSUPPORTED, "pmk_load_model: %s was written with a different memory layout", path);
  std::vector<char> stage((size_t)64 << 20);
  if (int rc = fill_dev(h, fc.f, h->d_xs.p, sizeof(double) * h->xstride * h->D, stage)) return rc;
  if (int rc = fill_dev(h, fc.f, h->d_y.p, sizeof(double) * h->xstride, stage)) return rc;
  if (int rc = fill_dev(h, fc.f, h->d_alpha.p, sizeof(double) * h->xstride, stage)) return rc;
  if (int rc = fill_dev(h, fc.f, h->d_L.p, sizeof(double) * (size_t)h->L_doubles, stage)) return rc;
  if (int rc = fill_dev(h, fc.f, h->d_Linv.p, sizeof(double) * (size_t)h->Linv_doubles, stage)) return rc;
  if (hd.tree_set) {
    if (hd.tree_D < 1 || hd.tree_D > PMK_MAX_DIM || hd.levels < 1 || hd.levels > 25 || hd.n_hp != (1 << (hd.levels - 1)) - 1)
      return fail(h, PMK_ERR_ARG, "pmk_load_model: corrupt tree header");
    h->tree_D = hd.tree_D;
    h->levels = hd.levels;
    h->n_hp = hd.n_hp;
    if (hd.n_hp > 0) {
      CU(h, h->d_hv.ensure(sizeof(double) * (size_t)hd.tree_D * hd.n_hp));
      CU(h, h->d_hc.ensure(sizeof(double) * (size_t)hd.n_hp));
      if (int rc = fill_dev(h, fc.f, h->d_hv.p, sizeof(double) * (size_t)hd.tree_D * hd.n_hp, stage)) return rc;
      if (int rc = fill_dev(h, fc.f, h->d_hc.p, sizeof(double) * (size_t)hd.n_hp, stage)) return rc;
    }
    h->tree.levels = hd.levels;
    h->tree.n_hp = hd.n_hp;
    h->tree.hv = h->d_hv.as<double>();
    h->tree.hc = h->d_hc.as<double>();
    h->tree_set = true;
    h->force_full_scan = hd.n_hp > 65535;
    if (hd.n_hp > 0 && !h->force_full_scan) {
      std::vector<double> soa((size_t)hd.tree_D * hd.n_hp);
      CU(h, cudaMemcpyAsync(soa.data(), h->d_hv.p, sizeof(double) * soa.size(), cudaMemcpyDeviceToHost, h->stream));
      CU(h, cudaStreamSynchronize(h->stream));
      for (int k = 0; k < hd.n_hp; ++k) {
        double nn = 0.0;
        for (int d = 0; d < hd.tree_D; ++d) nn += soa[(size_t)d * hd.n_hp + k] * soa[(size_t)d * hd.n_hp + k];
        if (!(std::fabs(std::sqrt(nn) - 1.0) <= 1e-9)) h->force_full_scan = true;
      }
    }
  }
  CU(h, cudaMemsetAsync(h->d_info.p, 0, sizeof(int) * h->n_leaves, h->stream));
  // the conditioning estimate is a function of the stored factors: recompute it, so that PMK_OPT_QUERY_SOLVER = -1 resolves
  // to the solver the saving handle used (bit-identical queries)
  CU(h, h->d_diag_range.ensure(sizeof(double)));
  CU(h, cudaMemsetAsync(h->d_diag_range.p, 0, sizeof(double), h->stream));
  launch_diag_range(h->lt, h->d_diag_range.as<double>(), h->stream);
  KCHECK(h, "k_diag_range");
  double cond_est = 0.0;
  CU(h, cudaMemcpyAsync(&cond_est, h->d_diag_range.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  h->cond_est = cond_est;
  h->fitted = true;
  h->plan_valid = false;
  return PMK_OK;
}

static int check_leaf(pmk_handle* h, int64_t leaf, int64_t* local) {
  if (!h) return PMK_ERR_ARG;
  if (h->n_leaves == 0) return fail(h, PMK_ERR_STATE, "model is not fitted");
  const int64_t p = leaf - 1 - h->leaf_base;
  if (p < 0 || p >= h->n_leaves) return fail(h, PMK_ERR_ARG, "leaf %lld not owned by this handle", (long long)leaf);
  *local = p;
  return PMK_OK;
}

int pmk_leaf_size(pmk_handle* h, int64_t leaf, int64_t* n_out) {
  int64_t p;
  if (int rc = check_leaf(h, leaf, &p)) return rc;
  if (n_out) *n_out = h->h_n[p];
  return PMK_OK;
}

int pmk_get_alpha(pmk_handle* h, int64_t leaf, double* out) {
  int64_t p;
  if (int rc = check_leaf(h, leaf, &p)) return rc;
  if (!out) return fail(h, PMK_ERR_ARG, "NULL pointer");
  if (int rc = set_device(h)) return rc;
  CU(h, cudaMemcpyAsync(out, h->d_alpha.as<double>() + h->h_xoff[p], sizeof(double) * h->h_n[p], cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return PMK_OK;
}

int pmk_get_X(pmk_handle* h, int64_t leaf, double* out) {
  int64_t p;
  if (int rc = check_leaf(h, leaf, &p)) return rc;
  if (!out) return fail(h, PMK_ERR_ARG, "NULL pointer");
  if (int rc = set_device(h)) return rc;
  const int n = h->h_n[p];
  std::vector<double> soa((size_t)n * h->D);
  for (int d = 0; d < h->D; ++d)
    CU(h, cudaMemcpyAsync(soa.data() + (size_t)d * n, h->d_xs.as<double>() + (size_t)d * h->xstride + h->h_xoff[p], sizeof(double) * n,
                          cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < h->D; ++d) out[(size_t)i * h->D + d] = soa[(size_t)d * n + i];
  return PMK_OK;
}

int pmk_model_info(pmk_handle* h, int* D, int64_t* n_leaves, int* kernel_id, double* kparam, double* sigma2, int* levels) {
  if (!h) return PMK_ERR_ARG;
  if (h->n_leaves == 0) return fail(h, PMK_ERR_STATE, "no model laid out (call pmk_fit or pmk_load_model first)");
  if (D) *D = h->D;
  if (n_leaves) *n_leaves = h->n_leaves;
  if (kernel_id) *kernel_id = h->kp.kind;
  if (kparam) *kparam = h->kp.p;
  if (sigma2) *sigma2 = h->sigma2;
  if (levels) *levels = h->tree_set ? h->levels : 0;
  return PMK_OK;
}

int pmk_get_tree(pmk_handle* h, double* hp_v, double* hp_c) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->tree_set) return fail(h, PMK_ERR_STATE, "no tree set (pmk_set_tree / pmk_load_model)");
  if (h->n_hp == 0) return PMK_OK;
  if (!hp_v || !hp_c) return fail(h, PMK_ERR_ARG, "NULL pointer");
  std::vector<double> soa((size_t)h->tree_D * h->n_hp);
  CU(h, cudaMemcpyAsync(soa.data(), h->d_hv.p, sizeof(double) * soa.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemcpyAsync(hp_c, h->d_hc.p, sizeof(double) * h->n_hp, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  for (int k = 0; k < h->n_hp; ++k)
    for (int d = 0; d < h->tree_D; ++d) hp_v[(size_t)k * h->tree_D + d] = soa[(size_t)d * h->n_hp + k];
  return PMK_OK;
}

int pmk_set_alpha(pmk_handle* h, int64_t leaf, const double* c) {
  int64_t p;
  if (int rc = check_leaf(h, leaf, &p)) return rc;
  if (!c) return fail(h, PMK_ERR_ARG, "NULL pointer");
  if (int rc = set_device(h)) return rc;
  CU(h, cudaMemcpyAsync(h->d_alpha.as<double>() + h->h_xoff[p], c, sizeof(double) * h->h_n[p], cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return PMK_OK;
}

int pmk_get_L(pmk_handle* h, int64_t leaf, double* out) {
  int64_t p;
  if (int rc = check_leaf(h, leaf, &p)) return rc;
  if (!out) return fail(h, PMK_ERR_ARG, "NULL pointer");
  if (int rc = set_device(h)) return rc;
  const int n = h->h_n[p];
  CU(h, h->d_scratch.ensure(sizeof(double) * (size_t)n * n));
  launch_unpack_L(h->lt, (int)p, n, h->d_scratch.as<double>(), h->stream);
  KCHECK(h, "k_unpack_L");
  CU(h, cudaMemcpyAsync(out, h->d_scratch.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return PMK_OK;
}

int pmk_get_Linv(pmk_handle* h, int64_t leaf, double* out) {
  int64_t p;
  if (int rc = check_leaf(h, leaf, &p)) return rc;
  if (!out) return fail(h, PMK_ERR_ARG, "NULL pointer");
  if (int rc = set_device(h)) return rc;
  if (int rc = build_operands(h, false, true)) return rc;
  const int n = h->h_n[p];
  CU(h, h->d_scratch.ensure(sizeof(double) * (size_t)n * n));
  launch_unpack_L(h->lt, (int)p, n, h->d_scratch.as<double>(), h->stream, 1);
  KCHECK(h, "k_unpack_L");
  CU(h, cudaMemcpyAsync(out, h->d_scratch.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return PMK_OK;
}

int pmk_get_K(pmk_handle* h, int64_t leaf, double* out) {
  int64_t p;
  if (int rc = check_leaf(h, leaf, &p)) return rc;
  if (!out) return fail(h, PMK_ERR_ARG, "NULL pointer");
  if (int rc = set_device(h)) return rc;
  const int n = h->h_n[p];
  CU(h, h->d_scratch.ensure(sizeof(double) * (size_t)n * n));
  const double* xs = h->d_xs.as<double>() + h->h_xoff[p];
  {
    Timer t(h, PMK_T_GRAM);
    launch_gram(h->D, xs, h->xstride, n, xs, h->xstride, n, h->kp, 0.0, 1, h->d_scratch.as<double>(), h->stream);
  }
  KCHECK(h, "k_gram");
  CU(h, cudaMemcpyAsync(out, h->d_scratch.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return PMK_OK;
}

// ---------------------------------------------------------------------------------------------
// tree
int pmk_set_tree(pmk_handle* h, int D, int levels, const double* hp_v, const double* hp_c) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (D < 1 || D > PMK_MAX_DIM) return fail(h, PMK_ERR_UNSUPPORTED, "D=%d unsupported", D);
  if (levels < 1 || levels > 25) return fail(h, PMK_ERR_ARG, "levels=%d out of range [1,25]", levels);
  const int n_hp = (1 << (levels - 1)) - 1;
  if (n_hp > 0 && (!hp_v || !hp_c)) return fail(h, PMK_ERR_ARG, "NULL hyperplane arrays");
  h->tree_D = D;
  h->levels = levels;
  h->n_hp = n_hp;
  // The pruned neighbour search compares |t| = |c - u.p| with the radius, which is the reference's norm(z - p) = |t| |u| only
  // for unit normals (setuppartition's are: V[:,1] of an svd), and keeps candidate slots in 16 bits.  Anything else takes the
  // reference's own scan over all hyperplanes.
  h->force_full_scan = n_hp > 65535;
  for (int k = 0; k < n_hp && !h->force_full_scan; ++k) {
    double nn = 0.0;
    for (int d = 0; d < D; ++d) nn += hp_v[(size_t)k * D + d] * hp_v[(size_t)k * D + d];
    if (!(std::fabs(std::sqrt(nn) - 1.0) <= 1e-9)) h->force_full_scan = true;
  }
  if (n_hp > 0) {
    std::vector<double> soa((size_t)D * n_hp);
    for (int k = 0; k < n_hp; ++k)
      for (int d = 0; d < D; ++d) soa[(size_t)d * n_hp + k] = hp_v[(size_t)k * D + d];
    CU(h, h->d_hv.ensure(sizeof(double) * soa.size()));
    CU(h, h->d_hc.ensure(sizeof(double) * n_hp));
    CU(h, cudaMemcpyAsync(h->d_hv.p, soa.data(), sizeof(double) * soa.size(), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->d_hc.p, hp_c, sizeof(double) * n_hp, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  h->tree.levels = levels;
  h->tree.n_hp = n_hp;
  h->tree.hv = h->d_hv.as<double>();
  h->tree.hc = h->d_hc.as<double>();
  h->tree_set = true;
  h->plan_valid = false;
  return PMK_OK;
}

int pmk_find_partition(pmk_handle* h, int64_t Nq, const double* Xq, int32_t* leaf_out) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->tree_set) return fail(h, PMK_ERR_STATE, "pmk_set_tree has not been called");
  if (Nq < 0) return fail(h, PMK_ERR_ARG, "Nq < 0");
  if (Nq == 0) return PMK_OK;
  if (!Xq || !leaf_out) return fail(h, PMK_ERR_ARG, "NULL pointer");
  const int D = h->tree_D;
  CU(h, h->d_Xq.ensure(sizeof(double) * Nq * D));
  CU(h, h->d_home.ensure(sizeof(int32_t) * Nq));
  CU(h, cudaMemcpyAsync(h->d_Xq.p, Xq, sizeof(double) * Nq * D, cudaMemcpyHostToDevice, h->stream));
  launch_home(D, h->tree, Nq, h->d_Xq.as<double>(), h->d_home.as<int32_t>(), nullptr, h->stream);
  KCHECK(h, "k_home");
  CU(h, cudaMemcpyAsync(leaf_out, h->d_home.p, sizeof(int32_t) * Nq, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  h->plan_valid = false;
  return PMK_OK;
}

// ---------------------------------------------------------------------------------------------
// organizetrainingsets on the device
int pmk_organize_training_sets(pmk_handle* h, int64_t N, const double* X0, double eps, int64_t* leaf_off_out,
                               int64_t* total_out) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->tree_set || h->n_hp == 0) return fail(h, PMK_ERR_STATE, "pmk_set_tree (levels >= 2) has not been called");
  if (N < 1 || N > INT32_MAX - 1 || !X0) return fail(h, PMK_ERR_ARG, "bad N or NULL X0");
  const int D = h->tree_D;
  const int64_t TL = (int64_t)h->n_hp + 1;
  CU(h, h->o_X.ensure(sizeof(double) * N * D));
  CU(h, h->o_counts.ensure(sizeof(int32_t) * (N + 1)));
  CU(h, h->o_off.ensure(sizeof(int64_t) * (N + 1)));
  CU(h, h->o_lcount.ensure(sizeof(int32_t) * TL));
  CU(h, h->o_lstart.ensure(sizeof(int64_t) * (TL + 1)));
  CU(h, cudaMemcpyAsync(h->o_X.p, X0, sizeof(double) * N * D, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemsetAsync(h->o_counts.as<int32_t>() + N, 0, sizeof(int32_t), h->stream));
  CU(h, cudaMemsetAsync(h->o_lcount.p, 0, sizeof(int32_t) * TL, h->stream));
  launch_eps_partitions(D, false, h->tree, N, h->o_X.as<double>(), eps, h->o_counts.as<int32_t>(), nullptr, nullptr, nullptr,
                        nullptr, h->stream);
  KCHECK(h, "k_eps_partitions<count>");
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveScan((void*)nullptr, tb, h->o_counts.as<int32_t>(), h->o_off.as<int64_t>(), cub::Sum(), (int64_t)0,
                                   (int)(N + 1), h->stream);
    CU(h, h->d_cub.ensure(tb));
    CU(h, cub::DeviceScan::ExclusiveScan(h->d_cub.p, tb, h->o_counts.as<int32_t>(), h->o_off.as<int64_t>(), cub::Sum(),
                                         (int64_t)0, (int)(N + 1), h->stream));
  }
  int64_t total = 0;
  CU(h, cudaMemcpyAsync(&total, h->o_off.as<int64_t>() + N, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  if (total < N || total > INT32_MAX) return fail(h, PMK_ERR_UNSUPPORTED, "pair count %lld out of range", (long long)total);
  CU(h, h->o_pl.ensure(sizeof(int32_t) * total));
  CU(h, h->o_pp.ensure(sizeof(int32_t) * total));
  CU(h, h->o_sl.ensure(sizeof(int32_t) * total));
  CU(h, h->o_sp.ensure(sizeof(int32_t) * total));
  launch_eps_partitions(D, true, h->tree, N, h->o_X.as<double>(), eps, nullptr, h->o_off.as<int64_t>(), h->o_pl.as<int32_t>(),
                        h->o_pp.as<int32_t>(), h->o_lcount.as<int32_t>(), h->stream);
  KCHECK(h, "k_eps_partitions<fill>");
  launch_scan_small(h->o_lcount.as<int32_t>(), h->o_lstart.as<int64_t>(), (int)TL, h->stream);
  KCHECK(h, "k_scan_small");
  {
    int end_bit = 1;
    while ((1ll << end_bit) <= TL) ++end_bit;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs((void*)nullptr, tb, h->o_pl.as<int32_t>(), h->o_sl.as<int32_t>(), h->o_pp.as<int32_t>(),
                                    h->o_sp.as<int32_t>(), (int)total, 0, end_bit, h->stream);
    CU(h, h->d_cub.ensure(tb));
    CU(h, cub::DeviceRadixSort::SortPairs(h->d_cub.p, tb, h->o_pl.as<int32_t>(), h->o_sl.as<int32_t>(), h->o_pp.as<int32_t>(),
                                          h->o_sp.as<int32_t>(), (int)total, 0, end_bit, h->stream));
  }
  if (leaf_off_out)
    CU(h, cudaMemcpyAsync(leaf_off_out, h->o_lstart.p, sizeof(int64_t) * (TL + 1), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  h->o_N = N;
  h->o_total = total;
  h->o_leaves = TL;
  if (total_out) *total_out = total;
  return PMK_OK;
}

int pmk_organize_fetch(pmk_handle* h, int32_t* inds_out, int64_t* point_off_out, int32_t* point_leaves_out) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (h->o_total == 0) return fail(h, PMK_ERR_STATE, "pmk_organize_training_sets has not been called");
  cudaStream_t s = h->stream;
  if (inds_out) CU(h, cudaMemcpyAsync(inds_out, h->o_sp.p, sizeof(int32_t) * h->o_total, cudaMemcpyDeviceToHost, s));
  if (point_off_out) CU(h, cudaMemcpyAsync(point_off_out, h->o_off.p, sizeof(int64_t) * (h->o_N + 1), cudaMemcpyDeviceToHost, s));
  if (point_leaves_out) CU(h, cudaMemcpyAsync(point_leaves_out, h->o_pl.p, sizeof(int32_t) * h->o_total, cudaMemcpyDeviceToHost, s));
  CU(h, cudaStreamSynchronize(s));
  return PMK_OK;
}

// ---------------------------------------------------------------------------------------------
// query
int pmk_query_plan_dev(pmk_handle* h, int64_t Nq, const double* dXq, double radius, double delta, int wkernel_id,
                       const double* wparams, int nw, int64_t* n_pairs_out) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  h->plan_valid = false;
  if (!h->fitted) return fail(h, PMK_ERR_STATE, "query before fit");
  if (!h->tree_set) return fail(h, PMK_ERR_STATE, "query before pmk_set_tree");
  if (h->tree_D != h->D) return fail(h, PMK_ERR_ARG, "tree dimension %d != model dimension %d", h->tree_D, h->D);
  if ((int64_t)(h->n_hp + 1) != h->total_leaves)
    return fail(h, PMK_ERR_ARG, "tree has %d leaves but the model has %lld (length(hps) == length(X_parts)-1)", h->n_hp + 1,
                (long long)h->total_leaves);
  if (Nq < 1 || Nq > INT32_MAX) return fail(h, PMK_ERR_ARG, "Nq=%lld out of range (the reference asserts !isempty(Xq))", (long long)Nq);
  if (!dXq) return fail(h, PMK_ERR_ARG, "NULL pointer");
  KParams wk;
  if (int rc = parse_kernel(h, wkernel_id, wparams, nw, &wk)) return rc;
  if (h->n_hp > 0 && !is_stationary_host(wk.kind))
    return fail(h, PMK_ERR_ARG, "weight kernel must be stationary (evalkernel(abs(t), weight_theta), mixtureGP.jl:231)");
  const int D = h->D;
  const int64_t TL = h->total_leaves;
  CU(h, h->d_home.ensure(sizeof(int32_t) * Nq));
  CU(h, h->d_npairs.ensure(sizeof(int32_t) * (Nq + 1)));
  CU(h, h->d_pair_off.ensure(sizeof(int64_t) * (Nq + 1)));
  CU(h, h->d_leaf_count.ensure(sizeof(int32_t) * TL));
  CU(h, h->d_leaf_pair_start.ensure(sizeof(int64_t) * (TL + 1)));

  QueryPlan& q = h->plan;
  q.Nq = Nq;
  q.Xq = dXq;
  q.home = h->d_home.as<int32_t>();
  q.npairs = h->d_npairs.as<int32_t>();
  q.pair_off = h->d_pair_off.as<int64_t>();

  Timer tt(h, PMK_T_Q_TREE);
  const bool pruned = !h->full_scan && !h->force_full_scan && h->n_hp > 0;
  int end_bit = 1;
  while ((1ll << end_bit) <= TL) ++end_bit;
  CU(h, cudaMemsetAsync(h->d_npairs.as<int32_t>() + Nq, 0, sizeof(int32_t), h->stream));
  CU(h, cudaMemsetAsync(h->d_leaf_count.p, 0, sizeof(int32_t) * TL, h->stream));
  if (pruned) {
    CU(h, h->d_leaf_qcount.ensure(sizeof(int32_t) * TL));
    CU(h, h->d_leaf_qstart.ensure(sizeof(int64_t) * (TL + 1)));
    CU(h, h->d_qperm.ensure(sizeof(int32_t) * Nq));
    CU(h, h->d_qkeys.ensure(sizeof(int32_t) * Nq));
    CU(h, h->d_bbox.ensure(sizeof(double) * TL * 2 * D));
    CU(h, h->d_cand_count.ensure(sizeof(int32_t) * TL));
    CU(h, h->d_cand_start.ensure(sizeof(int64_t) * (TL + 1)));
    CU(h, h->d_kept.ensure(sizeof(uint16_t) * 8 * Nq));
    if (int rc = ensure_iota(h, Nq)) return rc;
    CU(h, cudaMemsetAsync(h->d_leaf_qcount.p, 0, sizeof(int32_t) * TL, h->stream));
  }
  launch_home(D, h->tree, Nq, dXq, q.home, pruned ? h->d_leaf_qcount.as<int32_t>() : nullptr, h->stream);
  KCHECK(h, "k_home");
  if (pruned) {
    launch_scan_small(h->d_leaf_qcount.as<int32_t>(), h->d_leaf_qstart.as<int64_t>(), (int)TL, h->stream);
    KCHECK(h, "k_scan_small");
    {   // queries sorted (stably) by home leaf
      size_t tb = 0;
      cub::DeviceRadixSort::SortPairs((void*)nullptr, tb, q.home, h->d_qkeys.as<int32_t>(), h->d_iota.as<int32_t>(),
                                      h->d_qperm.as<int32_t>(), (int)Nq, 0, end_bit, h->stream);
      CU(h, h->d_cub.ensure(tb));
      CU(h, cub::DeviceRadixSort::SortPairs(h->d_cub.p, tb, q.home, h->d_qkeys.as<int32_t>(), h->d_iota.as<int32_t>(),
                                            h->d_qperm.as<int32_t>(), (int)Nq, 0, end_bit, h->stream));
    }
    launch_leaf_bbox(D, (int)TL, Nq, dXq, h->d_qperm.as<int32_t>(), h->d_leaf_qstart.as<int64_t>(), h->d_bbox.as<double>(), h->stream);
    KCHECK(h, "k_leaf_bbox");
    launch_leaf_candidates(D, false, (int)TL, h->tree, h->d_bbox.as<double>(), radius, h->d_cand_count.as<int32_t>(), nullptr,
                           nullptr, h->stream);
    KCHECK(h, "k_leaf_candidates<count>");
    launch_scan_small(h->d_cand_count.as<int32_t>(), h->d_cand_start.as<int64_t>(), (int)TL, h->stream);
    KCHECK(h, "k_scan_small");
    int64_t n_cand = 0;
    CU(h, cudaMemcpyAsync(&n_cand, h->d_cand_start.as<int64_t>() + TL, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, h->d_cand.ensure(sizeof(int32_t) * std::max<int64_t>(n_cand, 1)));
    launch_leaf_candidates(D, true, (int)TL, h->tree, h->d_bbox.as<double>(), radius, h->d_cand_count.as<int32_t>(),
                           h->d_cand_start.as<int64_t>(), h->d_cand.as<int32_t>(), h->stream);
    KCHECK(h, "k_leaf_candidates<fill>");
  }
  launch_neighbours(D, false, pruned, (int)TL, h->tree, q, radius, delta, wk.kind, wk.p, h->d_leaf_count.as<int32_t>(),
                    h->d_qperm.as<int32_t>(), h->d_leaf_qstart.as<int64_t>(), h->d_cand_start.as<int64_t>(),
                    h->d_cand.as<int32_t>(), h->d_kept.as<uint16_t>(), h->stream);
  KCHECK(h, "k_neighbours<count>");
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveScan((void*)nullptr, tb, q.npairs, q.pair_off, cub::Sum(), (int64_t)0, (int)(Nq + 1), h->stream);
    CU(h, h->d_cub.ensure(tb));
    CU(h, cub::DeviceScan::ExclusiveScan(h->d_cub.p, tb, q.npairs, q.pair_off, cub::Sum(), (int64_t)0, (int)(Nq + 1), h->stream));
  }
  int64_t n_pairs = 0;
  CU(h, cudaMemcpyAsync(&n_pairs, q.pair_off + Nq, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  if (n_pairs < Nq || n_pairs > INT32_MAX) return fail(h, PMK_ERR_UNSUPPORTED, "pair count %lld out of range", (long long)n_pairs);
  q.n_pairs = n_pairs;
  CU(h, h->d_pair_leaf.ensure(sizeof(int32_t) * n_pairs));
  CU(h, h->d_pair_q.ensure(sizeof(int32_t) * n_pairs));
  CU(h, h->d_pair_hp.ensure(sizeof(int32_t) * n_pairs));
  CU(h, h->d_pair_t.ensure(sizeof(double) * n_pairs));
  CU(h, h->d_pair_w.ensure(sizeof(double) * n_pairs));
  CU(h, h->d_sorted_pair.ensure(sizeof(int32_t) * n_pairs));
  CU(h, h->d_keys_out.ensure(sizeof(int32_t) * n_pairs));
  q.pair_leaf = h->d_pair_leaf.as<int32_t>();
  q.pair_q = h->d_pair_q.as<int32_t>();
  q.pair_hp = h->d_pair_hp.as<int32_t>();
  q.pair_t = h->d_pair_t.as<double>();
  q.pair_w = h->d_pair_w.as<double>();
  launch_neighbours(D, true, pruned, (int)TL, h->tree, q, radius, delta, wk.kind, wk.p, h->d_leaf_count.as<int32_t>(),
                    h->d_qperm.as<int32_t>(), h->d_leaf_qstart.as<int64_t>(), h->d_cand_start.as<int64_t>(),
                    h->d_cand.as<int32_t>(), h->d_kept.as<uint16_t>(), h->stream);
  KCHECK(h, "k_neighbours<fill>");
  launch_scan_small(h->d_leaf_count.as<int32_t>(), h->d_leaf_pair_start.as<int64_t>(), (int)TL, h->stream);
  KCHECK(h, "k_scan_small");
  // stable sort of pair ids by leaf
  {
    if (int rc = ensure_iota(h, n_pairs)) return rc;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs((void*)nullptr, tb, q.pair_leaf, h->d_keys_out.as<int32_t>(), h->d_iota.as<int32_t>(),
                                    h->d_sorted_pair.as<int32_t>(), (int)n_pairs, 0, end_bit, h->stream);
    CU(h, h->d_cub.ensure(tb));
    CU(h, cub::DeviceRadixSort::SortPairs(h->d_cub.p, tb, q.pair_leaf, h->d_keys_out.as<int32_t>(), h->d_iota.as<int32_t>(),
                                          h->d_sorted_pair.as<int32_t>(), (int)n_pairs, 0, end_bit, h->stream));
  }
  if (n_pairs_out) *n_pairs_out = n_pairs;
  h->last_radius = radius;
  h->last_delta = delta;
  h->plan_valid = true;
  return PMK_OK;
}

// The fused pair kernel over a pair list binned by leaf: q.Xq / q.pair_q give every pair's query point, sorted_pair the pair
// ids sorted by (global) leaf, leaf_pair_start the start of every global leaf's run.  Only this handle's leaves
// [leaf_base, leaf_base + n_leaves) may occur.  Writes u, v of pair id i to pu[i], pv[i].
static int run_pair_kernels(pmk_handle* h, const QueryPlan& q, const int64_t* leaf_pair_start, const int32_t* sorted_pair, int flags,
                            double* pu, double* pv) {
  const int mean_only = flags & 3;   // bit0: mean only; bit1: variance without the 1e-12 clamp
  const int solver = effective_solver(h);
  if (!(flags & 1)) {     // variance wanted: the pair kernel streams M (substitution) or P = inv(L), built once per fit
    const bool use_P = uses_inverse(h);
    if (int rc = build_operands(h, !use_P, use_P)) return rc;
  }
  if (h->lt.M == nullptr) h->lt.M = h->lt.L;   // mean-only queries never touch the factor
  Timer tt(h, PMK_T_Q_PAIRS);
  for (int c = 0; c < kNumClasses; ++c) {
    h->ev_used[PMK_T_Q_PAIRS_CLASS0 + c] = false;
    h->ms[PMK_T_Q_PAIRS_CLASS0 + c] = 0.0;
    if (h->n_class[c] == 0) continue;
    PairWork w;
    w.class_leaves = h->d_class_leaves[c].as<int>();
    w.n_class_leaves = h->n_class[c];
    w.tile_off = h->d_tile_off[c].as<int64_t>();
    w.leaf_pair_start = leaf_pair_start;
    w.sorted_pair = sorted_pair;
    w.leaf_base = h->leaf_base;
    const int mq = query_class_mq(c);
    launch_class_tiles(w, mq, h->d_class_tiles[c].as<int32_t>(), h->stream);
    KCHECK(h, "k_class_tiles");
    launch_scan_small(h->d_class_tiles[c].as<int32_t>(), h->d_tile_off[c].as<int64_t>(), h->n_class[c], h->stream);
    KCHECK(h, "k_scan_small");
    // upper bound on the tile count without a host round trip: excess CTAs exit at once
    const int64_t ub = q.n_pairs / mq + h->n_class[c];
    {
      Timer tc(h, PMK_T_Q_PAIRS_CLASS0 + c);
      if (!(mean_only & 1) && solver == 0) {
        if (!launch_query_rowp(h->D, c, h->lt, w, q, h->kp, mean_only, h->class_max_npad[c], pu, pv, h->stream))
          return fail(h, PMK_ERR_UNSUPPORTED, "row-panel pair kernel: no shared-memory configuration for size class %d", c);
      } else
        launch_query_pairs(h->D, c, (unsigned)ub, h->lt, w, q, h->kp, mean_only, pu, pv, h->stream);
    }
    KCHECK(h, "k_query_pairs");
  }
  return PMK_OK;
}

int pmk_query_pairs_dev(pmk_handle* h, int flags, double* d_pair_u, double* d_pair_v) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->plan_valid) return fail(h, PMK_ERR_STATE, "pmk_query_plan_dev has not been called");
  if (!d_pair_u || !d_pair_v) return fail(h, PMK_ERR_ARG, "NULL pointer");
  if (h->leaf_base != 0 || h->n_leaves != h->total_leaves)
    return fail(h, PMK_ERR_STATE, "this handle owns leaves [%lld, %lld) of %lld: route the pairs to their owners (pmk_query_plan_segments / "
                "pmk_query_pairs_routed_dev, or pmk_multi_query)", (long long)h->leaf_base, (long long)(h->leaf_base + h->n_leaves),
                (long long)h->total_leaves);
  h->last_flags = flags;
  return run_pair_kernels(h, h->plan, h->d_leaf_pair_start.as<int64_t>(), h->d_sorted_pair.as<int32_t>(), flags, d_pair_u, d_pair_v);
}

// ---- sub-tree ownership: the pairs of a plan travel to the handles that own their leaves ------------------------------------
int pmk_set_leaf_base(pmk_handle* h, int64_t leaf_base, int64_t total_leaves) {
  if (!h) return PMK_ERR_ARG;
  if (leaf_base < 0 || total_leaves < 0 || (total_leaves > 0 && leaf_base >= total_leaves)) return fail(h, PMK_ERR_ARG, "bad leaf base");
  h->leaf_base = leaf_base;
  h->total_opt = total_leaves;
  h->fitted = false;
  h->plan_valid = false;
  return PMK_OK;
}

int pmk_query_plan_segments(pmk_handle* h, int n_owners, const int64_t* owner_first_leaf, int64_t* seg_off) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->plan_valid) return fail(h, PMK_ERR_STATE, "pmk_query_plan_dev has not been called");
  if (n_owners < 1 || !owner_first_leaf || !seg_off) return fail(h, PMK_ERR_ARG, "NULL pointer or n_owners < 1");
  const int64_t TL = h->total_leaves;
  for (int o = 0; o <= n_owners; ++o) {
    const int64_t f = owner_first_leaf[o];
    if (f < 0 || f > TL || (o > 0 && f < owner_first_leaf[o - 1]) || (o == 0 && f != 0) || (o == n_owners && f != TL))
      return fail(h, PMK_ERR_ARG, "owner_first_leaf must ascend from 0 to the number of leaves (%lld)", (long long)TL);
  }
  // leaf_pair_start has TL + 1 entries: the start of every leaf's run in the leaf-sorted pair list
  for (int o = 0; o <= n_owners; ++o)
    CU(h, cudaMemcpyAsync(seg_off + o, h->d_leaf_pair_start.as<int64_t>() + owner_first_leaf[o], sizeof(int64_t), cudaMemcpyDeviceToHost,
                          h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return PMK_OK;
}

int pmk_query_plan_pack_dev(pmk_handle* h, double* d_X_sorted, int32_t* d_leaf_sorted) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->plan_valid) return fail(h, PMK_ERR_STATE, "pmk_query_plan_dev has not been called");
  if (!d_X_sorted || !d_leaf_sorted) return fail(h, PMK_ERR_ARG, "NULL pointer");
  const QueryPlan& q = h->plan;
  launch_pack_sorted_pairs(h->D, q, h->d_sorted_pair.as<int32_t>(), d_X_sorted, d_leaf_sorted, h->stream);
  KCHECK(h, "k_pack_sorted_pairs");
  return PMK_OK;
}

int pmk_query_plan_unpack_dev(pmk_handle* h, const double* d_u_sorted, const double* d_v_sorted, double* d_pair_u, double* d_pair_v) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->plan_valid) return fail(h, PMK_ERR_STATE, "pmk_query_plan_dev has not been called");
  if (!d_u_sorted || !d_pair_u) return fail(h, PMK_ERR_ARG, "NULL pointer");
  launch_unpack_sorted_pairs(h->plan.n_pairs, h->d_sorted_pair.as<int32_t>(), d_u_sorted, d_v_sorted, d_pair_u, d_pair_v, h->stream);
  KCHECK(h, "k_unpack_sorted_pairs");
  return PMK_OK;
}

int pmk_query_set_flags(pmk_handle* h, int flags) {
  if (!h) return PMK_ERR_ARG;
  h->last_flags = flags;
  return PMK_OK;
}

int pmk_query_pairs_routed_dev(pmk_handle* h, int64_t R, const double* d_X, const int32_t* d_leaf, int flags, double* d_u, double* d_v) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->fitted) return fail(h, PMK_ERR_STATE, "query before fit");
  if (R < 0 || R > INT32_MAX) return fail(h, PMK_ERR_ARG, "R=%lld out of range", (long long)R);
  if (R == 0) return PMK_OK;
  if (!d_X || !d_leaf || !d_u || (!(flags & 1) && !d_v)) return fail(h, PMK_ERR_ARG, "NULL pointer");
  const int64_t TL = h->total_leaves;
  CU(h, h->r_keys.ensure(sizeof(int32_t) * R));
  CU(h, h->r_sorted.ensure(sizeof(int32_t) * R));
  CU(h, h->r_leaf_start.ensure(sizeof(int64_t) * (TL + 1)));
  if (int rc = ensure_iota(h, R)) return rc;
  int end_bit = 1;
  while ((1ll << end_bit) <= TL) ++end_bit;
  {
    Timer tt(h, PMK_T_Q_ROUTE_SORT);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs((void*)nullptr, tb, d_leaf, h->r_keys.as<int32_t>(), h->d_iota.as<int32_t>(), h->r_sorted.as<int32_t>(),
                                    (int)R, 0, end_bit, h->stream);
    CU(h, h->d_cub.ensure(tb));
    CU(h, cub::DeviceRadixSort::SortPairs(h->d_cub.p, tb, d_leaf, h->r_keys.as<int32_t>(), h->d_iota.as<int32_t>(), h->r_sorted.as<int32_t>(),
                                          (int)R, 0, end_bit, h->stream));
    ++h->launches;
    // start of every global leaf's run in the sorted keys; flags any leaf outside [leaf_base, leaf_base + n_leaves)
    CU(h, cudaMemsetAsync(h->d_info2.p, 0, sizeof(int), h->stream));
    launch_run_starts(h->r_keys.as<int32_t>(), R, TL, h->leaf_base, h->n_leaves, h->r_leaf_start.as<int64_t>(), h->d_info2.as<int>(), h->stream);
    KCHECK(h, "k_run_starts");
  }
  QueryPlan q{};
  q.Nq = R;
  q.Xq = d_X;
  q.n_pairs = R;
  q.pair_q = h->d_iota.as<int32_t>();
  if (int rc = run_pair_kernels(h, q, h->r_leaf_start.as<int64_t>(), h->r_sorted.as<int32_t>(), flags, d_u, d_v)) return rc;
  int foreign = 0;
  CU(h, cudaMemcpyAsync(&foreign, h->d_info2.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  if (foreign) return fail(h, PMK_ERR_ARG, "routed pairs name leaves this handle does not own ([%lld, %lld) of %lld)", (long long)h->leaf_base + 1,
                           (long long)(h->leaf_base + h->n_leaves + 1), (long long)TL);
  return PMK_OK;
}

int pmk_query_combine_dev(pmk_handle* h, const double* d_pair_u, const double* d_pair_v, double* dYq, double* dVq) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->plan_valid) return fail(h, PMK_ERR_STATE, "pmk_query_plan_dev has not been called");
  const int mean_only = h->last_flags & 1;
  if (!d_pair_u || !dYq || (!mean_only && (!d_pair_v || !dVq))) return fail(h, PMK_ERR_ARG, "NULL pointer");
  const QueryPlan& q = h->plan;
  Timer tt(h, PMK_T_Q_COMBINE);
  launch_combine(q.Nq, q.pair_off, q.pair_w, d_pair_u, d_pair_v, dYq, dVq, mean_only, h->stream);
  KCHECK(h, "k_combine");
  return PMK_OK;
}

int pmk_query_dev(pmk_handle* h, int64_t Nq, const double* dXq, double radius, double delta, int wkernel_id,
                  const double* wparams, int nw, int flags, double* dYq, double* dVq) {
  if (!h) return PMK_ERR_ARG;
  int64_t np = 0;
  if (int rc = pmk_query_plan_dev(h, Nq, dXq, radius, delta, wkernel_id, wparams, nw, &np)) return rc;
  CU(h, h->d_pair_u.ensure(sizeof(double) * np));
  CU(h, h->d_pair_v.ensure(sizeof(double) * np));
  if (int rc = pmk_query_pairs_dev(h, flags, h->d_pair_u.as<double>(), h->d_pair_v.as<double>())) return rc;
  return pmk_query_combine_dev(h, h->d_pair_u.as<double>(), h->d_pair_v.as<double>(), dYq, dVq);
}

int pmk_query(pmk_handle* h, int64_t Nq, const double* Xq, double radius, double delta, int wkernel_id, const double* wparams,
              int nw, int flags, double* Yq, double* Vq) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->fitted) return fail(h, PMK_ERR_STATE, "query before fit");
  if (Nq < 1) return fail(h, PMK_ERR_ARG, "Nq < 1 (the reference asserts !isempty(Xq))");
  const int mean_only = flags & 1;
  if (!Xq || !Yq || (!mean_only && !Vq)) return fail(h, PMK_ERR_ARG, "NULL pointer");
  const int D = h->D;
  CU(h, h->d_Xq.ensure(sizeof(double) * Nq * D));
  CU(h, h->d_Yq.ensure(sizeof(double) * Nq));
  CU(h, h->d_Vq.ensure(sizeof(double) * Nq));
  CU(h, cudaMemcpyAsync(h->d_Xq.p, Xq, sizeof(double) * Nq * D, cudaMemcpyHostToDevice, h->stream));
  if (int rc = pmk_query_dev(h, Nq, h->d_Xq.as<double>(), radius, delta, wkernel_id, wparams, nw, flags, h->d_Yq.as<double>(),
                             h->d_Vq.as<double>()))
    return rc;
  CU(h, cudaMemcpyAsync(Yq, h->d_Yq.p, sizeof(double) * Nq, cudaMemcpyDeviceToHost, h->stream));
  if (!mean_only) CU(h, cudaMemcpyAsync(Vq, h->d_Vq.p, sizeof(double) * Nq, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return PMK_OK;
}

int pmk_last_query_pairs(pmk_handle* h, int64_t* n_pairs) {
  if (!h) return PMK_ERR_ARG;
  if (!h->plan_valid) return fail(h, PMK_ERR_STATE, "no query plan");
  if (n_pairs) *n_pairs = h->plan.n_pairs;
  return PMK_OK;
}

int pmk_last_query_debug(pmk_handle* h, int32_t* home, int64_t* pair_off, int32_t* pair_leaf, int32_t* pair_hp, double* pair_t,
                         double* pair_w, double* pair_u, double* pair_v) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->plan_valid) return fail(h, PMK_ERR_STATE, "no query plan");
  const QueryPlan& q = h->plan;
  const int64_t np = q.n_pairs;
  cudaStream_t s = h->stream;
  if (home) CU(h, cudaMemcpyAsync(home, q.home, sizeof(int32_t) * q.Nq, cudaMemcpyDeviceToHost, s));
  if (pair_off) CU(h, cudaMemcpyAsync(pair_off, q.pair_off, sizeof(int64_t) * (q.Nq + 1), cudaMemcpyDeviceToHost, s));
  if (pair_leaf) CU(h, cudaMemcpyAsync(pair_leaf, q.pair_leaf, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, s));
  if (pair_hp) CU(h, cudaMemcpyAsync(pair_hp, q.pair_hp, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, s));
  if (pair_t) CU(h, cudaMemcpyAsync(pair_t, q.pair_t, sizeof(double) * np, cudaMemcpyDeviceToHost, s));
  if (pair_w) CU(h, cudaMemcpyAsync(pair_w, q.pair_w, sizeof(double) * np, cudaMemcpyDeviceToHost, s));
  if (pair_u) {
    if (!h->d_pair_u.p) return fail(h, PMK_ERR_STATE, "pair_u is caller-owned in the split query");
    CU(h, cudaMemcpyAsync(pair_u, h->d_pair_u.p, sizeof(double) * np, cudaMemcpyDeviceToHost, s));
  }
  if (pair_v) {
    if (!h->d_pair_v.p) return fail(h, PMK_ERR_STATE, "pair_v is caller-owned in the split query");
    CU(h, cudaMemcpyAsync(pair_v, h->d_pair_v.p, sizeof(double) * np, cudaMemcpyDeviceToHost, s));
  }
  CU(h, cudaStreamSynchronize(s));
  return PMK_OK;
}

int pmk_last_query_debug_dense(pmk_handle* h, int64_t first_query, int64_t n_queries, uint8_t* keep_flags, double* ts, double* zs) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->plan_valid) return fail(h, PMK_ERR_STATE, "no query plan");
  const QueryPlan& q = h->plan;
  if (first_query < 0 || n_queries < 0 || first_query + n_queries > q.Nq) return fail(h, PMK_ERR_ARG, "query range out of bounds");
  const int64_t total = n_queries * (int64_t)h->n_hp;
  if (total == 0) return PMK_OK;
  if (total > ((int64_t)1 << 28))
    return fail(h, PMK_ERR_UNSUPPORTED, "dense debug arrays of %lld queries x %d hyperplanes: ask for at most 2^28 entries per call",
                (long long)n_queries, h->n_hp);
  const int D = h->D;
  const size_t bt = (size_t)total * sizeof(double), bz = bt * D, bk = ((size_t)total + 7) / 8 * 8;
  CU(h, h->d_scratch.ensure(bt + bz + bk));
  double* dts = h->d_scratch.as<double>();
  double* dzs = dts + total;
  uint8_t* dk = reinterpret_cast<uint8_t*>(dzs + (size_t)total * D);
  QueryPlan sub = q;
  sub.Nq = n_queries;
  sub.Xq = q.Xq + first_query * D;
  sub.home = q.home + first_query;
  launch_dense_debug(D, h->tree, sub, h->last_radius, h->last_delta, keep_flags ? dk : nullptr, ts ? dts : nullptr, zs ? dzs : nullptr,
                     h->stream);
  KCHECK(h, "k_dense_debug");
  if (ts) CU(h, cudaMemcpyAsync(ts, dts, bt, cudaMemcpyDeviceToHost, h->stream));
  if (zs) CU(h, cudaMemcpyAsync(zs, dzs, bz, cudaMemcpyDeviceToHost, h->stream));
  if (keep_flags) CU(h, cudaMemcpyAsync(keep_flags, dk, (size_t)total, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return PMK_OK;
}

int pmk_measure_fp64_peak(pmk_handle* h, double* dmma_tflops) {
  if (!h || !dmma_tflops) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  CU(h, h->d_scratch.ensure(1024));
  cudaEvent_t e0, e1;
  CU(h, cudaEventCreate(&e0));
  CU(h, cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {      // first launch warms up; best of the other three
    cudaEventRecord(e0, h->stream);
    const double flops = launch_dmma_peak(h->d_scratch.as<double>(), 4000, h->stream);
    cudaEventRecord(e1, h->stream);
    if (cudaGetLastError() != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) {
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      return fail(h, PMK_ERR_CUDA, "k_dmma_peak failed");
    }
    ++h->launches;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *dmma_tflops = best;
  return PMK_OK;
}

int pmk_condition_estimate(pmk_handle* h, double* cond_lower_bound, int* solver_in_use) {
  if (!h) return PMK_ERR_ARG;
  if (h->n_leaves == 0) return fail(h, PMK_ERR_STATE, "no model laid out (call pmk_fit first)");
  if (cond_lower_bound) *cond_lower_bound = h->cond_est;
  if (solver_in_use) *solver_in_use = effective_solver(h);
  return PMK_OK;
}

int pmk_last_query_leaf_pairs(pmk_handle* h, int64_t* pairs_per_leaf) {
  if (!h) return PMK_ERR_ARG;
  if (int rc = set_device(h)) return rc;
  if (!h->plan_valid) return fail(h, PMK_ERR_STATE, "no query plan");
  if (!pairs_per_leaf) return fail(h, PMK_ERR_ARG, "NULL pointer");
  const int64_t TL = h->total_leaves;
  std::vector<int64_t> st((size_t)TL + 1);
  CU(h, cudaMemcpyAsync(st.data(), h->d_leaf_pair_start.p, sizeof(int64_t) * (TL + 1), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  for (int64_t g = 0; g < TL; ++g) pairs_per_leaf[g] = st[g + 1] - st[g];
  return PMK_OK;
}

}  // extern "C"
