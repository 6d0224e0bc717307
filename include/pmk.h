/* pmk.h -- C ABI of libpmk_b200.so: the B200-native (sm_100a) local-GP fit + mixture-query
 * hot path of RoyCCWang/PatchMixtureKriging.
 *
 * The reference is pure Julia and has no FFI of its own; the drop-in boundary is its Julia
 * function surface (src/PatchMixtureKriging.jl:54-71).  Each entry point below names the
 * reference function whose BODY it replaces (paths relative to the reference repo);
 * julia/PatchMixtureKrigingB200.jl and INTEGRATION.md show the `ccall` a maintainer adds.
 *
 * Conventions
 *  - plain C types only; all matrices column-major (Julia layout); points are "point-major":
 *    X is D x n column-major, i.e. exactly src/misc/utilities.jl:25-36 `array2matrix(X)`.
 *  - every index RETURNED is 1-based (leaf ids, hyperplane ids) so it compares bit-for-bit
 *    with the reference; offsets/CSR pointers PASSED IN are 0-based prefix sums.
 *  - the caller owns every host buffer for the duration of the (blocking) call; the library
 *    owns all device memory behind the opaque handle and retains no host pointer.
 *  - return value: PMK_OK or a negative pmk_status; pmk_last_error() gives the text.
 *    No C++ exception crosses this boundary.  There is NO CPU fallback: without a CUDA
 *    device pmk_create fails with PMK_ERR_CUDA.
 *  - one pmk_handle = one fitted model (or one rank's sub-tree of it) on one GPU; one pmk_multi = one model sharded over the
 *    GPUs of a box by sub-tree ownership (below).  Calls on one handle must be serialised.
 *  - *_dev variants take DEVICE pointers (same layouts) and run on the handle's stream
 *    without host synchronisation beyond what is documented.
 */
#ifndef PMK_H
#define PMK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmk_handle pmk_handle;

typedef enum {
  PMK_OK = 0,
  PMK_ERR_CUDA = -1,         /* CUDA runtime error (text in pmk_last_error)                    */
  PMK_ERR_ARG = -2,          /* bad argument / size mismatch  (-> AssertionError/DimensionMismatch,
                                mixtureGP.jl:298, RKHS.jl:18,199-203,225-227)                  */
  PMK_ERR_NOT_POSDEF = -3,   /* a leaf's K+sigma2*I is not PD (-> PosDefException(info),
                                mixtureGP.jl:109); bad_leaf (1-based) and LAPACK-style info set */
  PMK_ERR_STATE = -4,        /* call order: query before fit / set_tree                        */
  PMK_ERR_UNSUPPORTED = -5   /* D > 3, leaf larger than PMK_MAX_LEAF_POINTS, unknown kernel id */
} pmk_status;

/* kernel ids: src/misc/declarations.jl:25-100, evaluated as in src/RKHS/kernel.jl */
typedef enum {
  PMK_KERNEL_SQEXP = 0,     /* GaussianKernel1DType{eps_sq}: exp(-eps_sq*tau^2)      kernel.jl:350-357 */
  PMK_KERNEL_SPLINE34 = 1,  /* Spline34KernelType{a}                                 kernel.jl:299-313 */
  PMK_KERNEL_BB10 = 2,      /* BrownianBridge10: prod_d min(p,q)-p*q                 kernel.jl:156-158,196-198 */
  PMK_KERNEL_BB20 = 3,      /* BrownianBridge20 (iterated, beta=2)                   kernel.jl:218-225 */
  PMK_KERNEL_BB1EPS = 4,    /* BrownianBridge1eps{eps}                               kernel.jl:168-174 */
  PMK_KERNEL_BB2EPS = 5,    /* BrownianBridge2eps{eps}                               kernel.jl:176-193 */
  PMK_KERNEL_SPLINE12 = 6,  /* Spline12KernelType{a}                                 kernel.jl:316-330 */
  PMK_KERNEL_SPLINE32 = 7,  /* Spline32KernelType{a}                                 kernel.jl:333-347 */
  PMK_KERNEL_RQ = 8         /* RationalQuadraticKernelType{a}                        kernel.jl:360-366 */
} pmk_kernel_id;

#define PMK_MAX_DIM 3
#define PMK_MAX_LEAF_POINTS 2048

/* timing slots of pmk_get_timings (milliseconds, CUDA events on the handle's stream) */
enum {
  PMK_T_FIT_PACK = 0,      /* AoS -> padded SoA leaf packing                                   */
  PMK_T_FIT_CHOL = 1,      /* batched blocked Cholesky (K2)                                    */
  PMK_T_FIT_SOLVE = 2,     /* forward/back solves for alpha                                    */
  PMK_T_Q_TREE = 3,        /* home leaf + neighbour search + pair build + binning              */
  PMK_T_Q_PAIRS = 4,       /* fused cross-covariance / mean / variance pair kernel (K3)        */
  PMK_T_Q_COMBINE = 5,     /* convex mixture combine                                           */
  PMK_T_GRAM = 6,          /* standalone Gram kernel (constructkernelmatrix / U_set)           */
  PMK_T_FIT_GRAM = 7,      /* per-leaf Gram tiles of the fit (K1)                              */
  PMK_T_Q_PAIRS_CLASS0 = 8,  /* .. +4: the fused pair kernel per leaf-size class (<=512, <=768, <=1024, <=1536, <=2048) */
  PMK_T_Q_MAKE_M = 13,       /* M_IJ = L_IJ inv(L_JJ), built once per fit by the first variance query               */
  PMK_T_Q_INVERT = 14,       /* P = inv(L) (operand of the explicit-inverse pair kernels), built once per fit       */
  PMK_T_FIT_REFINE = 15,     /* one step of iterative refinement of alpha (ill-conditioned models only)              */
  PMK_T_Q_ROUTE_SORT = 16,   /* owner side of routed pairs: binning the received pairs by leaf                       */
  PMK_T_COUNT = 17
};

/* ---- lifetime ---------------------------------------------------------------------------- */
int pmk_create(pmk_handle** out, int device);
void pmk_destroy(pmk_handle* h);
const char* pmk_last_error(const pmk_handle* h);   /* h may be NULL: error of a failed pmk_create */
int pmk_version(void);

/* ---- Gram matrices ----------------------------------------------------------------------- */
/* constructkernelmatrix(X, theta) (src/RKHS/RKHS.jl:4-34): K_out[i + n*j] = k(X[i], X[j]) (+ sigma2
 * on the diagonal; pass 0 for the plain Gram matrix).  Full symmetric n x n, column-major. */
int pmk_gram(pmk_handle* h, int D, int64_t n, const double* X, int kernel_id, const double* kparams,
             int nparams, double sigma2, double* K_out);
/* constructkernelmatrix(X, Z, theta) (RKHS.jl:95-110): K_out[i + n*j] = k(X[i], Z[j]), n x m. */
int pmk_cross_gram(pmk_handle* h, int D, int64_t n, const double* X, int64_t m, const double* Z,
                   int kernel_id, const double* kparams, int nparams, double* K_out);

/* ---- fit --------------------------------------------------------------------------------- */
/* fitmixtureGP!(eta, y_parts, theta, sigma2) (src/RKHS/mixtureGP.jl:70-118) for all leaves at once;
 * with n_leaves == 1 it is fitRKHS!(eta, y) (RKHS.jl:182-217).
 *   leaf_off : n_leaves+1 prefix offsets (0-based, in points) into X_packed / y_packed
 *   X_packed : D x sum(n_p), the leaves' training inputs back to back (array2matrix of each X_set[p])
 *   y_packed : sum(n_p)
 * On PMK_ERR_NOT_POSDEF, *bad_leaf = first (lowest) failing leaf, 1-based, *info = order of the
 * first non-positive leading minor (LAPACK dpotrf convention).  Either may be NULL. */
int pmk_fit(pmk_handle* h, int D, int64_t n_leaves, const int64_t* leaf_off, const double* X_packed,
            const double* y_packed, int kernel_id, const double* kparams, int nparams, double sigma2,
            int64_t* bad_leaf, int* info);
/* same with device pointers for X_packed / y_packed (leaf_off stays a host array) */
int pmk_fit_dev(pmk_handle* h, int D, int64_t n_leaves, const int64_t* leaf_off, const double* dX_packed,
                const double* dy_packed, int kernel_id, const double* kparams, int nparams, double sigma2,
                int64_t* bad_leaf, int* info);

/* per-leaf state of MixtureGPType (mixtureGP.jl:38-66); leaf is 1-based.
 *   alpha: c_set[leaf] (n_p);  L: L_set[leaf] as dense n_p x n_p column-major, upper triangle zero;
 *   K: U_set[leaf] = Gram WITHOUT sigma2 (mixtureGP.jl:99), recomputed on demand. */
int pmk_leaf_size(pmk_handle* h, int64_t leaf, int64_t* n_out);
int pmk_get_alpha(pmk_handle* h, int64_t leaf, double* out);
/* overwrite a leaf's weights (setupGPquery(c, X, theta, sigma2), src/RKHS/querying.jl:43-58, takes c from the caller) */
int pmk_set_alpha(pmk_handle* h, int64_t leaf, const double* c);
int pmk_get_L(pmk_handle* h, int64_t leaf, double* out);
/* dense n x n column-major inv(L) as the explicit-inverse solver uses it (built on demand) */
int pmk_get_Linv(pmk_handle* h, int64_t leaf, double* out);
int pmk_get_K(pmk_handle* h, int64_t leaf, double* out);

/* ---- tree -------------------------------------------------------------------------------- */
/* the BSP built by setuppartition (src/patchwork/partition.jl:106-129), flattened by the host wrapper:
 * hyperplanes of the 2^(levels-1)-1 internal nodes in fetchhyperplanes order = PreOrderDFS
 * (mixtureGP.jl:322-334): hp_v is D x n_hp column-major, hp_c has n_hp entries.  levels == 1 means
 * "no tree" (a single GP: every query's home leaf is 1). */
int pmk_set_tree(pmk_handle* h, int D, int levels, const double* hp_v, const double* hp_c);

/* findpartition (partition.jl:248-262) for many points: leaf_out[j] 1-based. */
int pmk_find_partition(pmk_handle* h, int64_t Nq, const double* Xq, int32_t* leaf_out);

/* organizetrainingsets(root, levels, X0, eps) (partition.jl:301-357; findεpartitions! :269-298) on the device, for the
 * tree given to pmk_set_tree.  X0 is D x N.  leaf_off_out gets n_leaves+1 0-based prefix offsets, *total_out the number of
 * (leaf, point) memberships.  pmk_organize_fetch then returns
 *   inds_out[total]          X_set_inds, leaf after leaf, ascending 1-based global point ids inside each leaf
 *   point_off_out[N+1], point_leaves_out[total]   regions_list_set: every point's leaves, left-to-right, 1-based
 * (any pointer may be NULL).  Bit-exact against the reference's comparisons (v.x < c+eps / v.x > c-eps, un-fused). */
int pmk_organize_training_sets(pmk_handle* h, int64_t N, const double* X0, double eps, int64_t* leaf_off_out,
                               int64_t* total_out);
int pmk_organize_fetch(pmk_handle* h, int32_t* inds_out, int64_t* point_off_out, int32_t* point_leaves_out);

/* ---- query ------------------------------------------------------------------------------- */
/* querymixtureGP!(Yq, Vq, Xq, eta, root, levels, radius, delta, theta, sigma2, weight_theta, ...)
 * (mixtureGP.jl:159-294; inner: queryinner! :296-316, findneighbourpartitions :339-405,
 * findpartition partition.jl:248-262).  Xq is D x Nq.  wkernel_id/wparams = weight_theta (a stationary
 * kernel evaluated at abs(t)).  theta and sigma2 are those given to pmk_fit.
 * flags: bit0 = mean only (Vq untouched; this is query!(Yq,Xq,eta), RKHS.jl:220-247, when levels==1);
 *        bit1 = variance without the clamp(., 1e-12, Inf) of queryinner! (evalqueryGP!, querying.jl:60-79). */
int pmk_query(pmk_handle* h, int64_t Nq, const double* Xq, double radius, double delta, int wkernel_id,
              const double* wparams, int nw, int flags, double* Yq, double* Vq);
int pmk_query_dev(pmk_handle* h, int64_t Nq, const double* dXq, double radius, double delta, int wkernel_id,
                  const double* wparams, int nw, int flags, double* dYq, double* dVq);

/* debug_flag=true outputs (MixtureGPDebugType, mixtureGP.jl:5-35,242-260) of the LAST query:
 *   home[Nq]           p_region_ind_set (1-based)
 *   pair_off[Nq+1]     CSR offsets: query j owns slots pair_off[j] .. pair_off[j+1]-1; the last slot is
 *                      the home leaf (w_tilde = 1), the others the kept neighbours in hyperplane order
 *   pair_leaf          region_inds_set (+ home last), 1-based
 *   pair_hp            1-based index into hps of the hyperplane that produced the slot (0 for home)
 *   pair_t             ts[hps_keep_flags]  (0 for home)
 *   pair_w, pair_u, pair_v   w_tilde_set, u_set, v_set
 * Call pmk_last_query_pairs first to size the pair arrays.  Any output pointer may be NULL. */
int pmk_last_query_pairs(pmk_handle* h, int64_t* n_pairs);
int pmk_last_query_debug(pmk_handle* h, int32_t* home, int64_t* pair_off, int32_t* pair_leaf, int32_t* pair_hp,
                         double* pair_t, double* pair_w, double* pair_u, double* pair_v);

/* debug_flag=true outputs that are dense over the hyperplanes (mixtureGP.jl:17-19,256-258: hps_keep_flags_set, zs_set, ts_set)
 * for queries [first_query, first_query + n_queries) of the LAST query: for every hyperplane i (fetchhyperplanes order)
 *   ts[j*n_hp + i] = t_i = -dot(u_i, p_j) + c_i,  zs[(j*n_hp + i)*D + d] = p_j + t_i u_i,  keep_flags[j*n_hp + i] = 0 / 1
 * (findneighbourpartitions, mixtureGP.jl:347-352,361-365,389).  n_queries * n_hp <= 2^28 per call; any pointer may be NULL. */
int pmk_last_query_debug_dense(pmk_handle* h, int64_t first_query, int64_t n_queries, uint8_t* keep_flags, double* ts, double* zs);
/* number of (query, leaf) pairs of the last plan per leaf (n_leaves of the whole model entries) */
int pmk_last_query_leaf_pairs(pmk_handle* h, int64_t* pairs_per_leaf);

/* ---- query in stages on DEVICE buffers ---------------------------------------------------- */
/* pmk_query_dev = the three in sequence:
 *   pmk_query_plan_dev   : home leaves, neighbours, weights, pair list binned by leaf; returns n_pairs
 *   pmk_query_pairs_dev  : fused pair kernel; writes u,v of every pair into n_pairs-long device arrays
 *   pmk_query_combine_dev: convex combination -> Yq, Vq. */
int pmk_query_plan_dev(pmk_handle* h, int64_t Nq, const double* dXq, double radius, double delta,
                       int wkernel_id, const double* wparams, int nw, int64_t* n_pairs);
int pmk_query_pairs_dev(pmk_handle* h, int flags, double* d_pair_u, double* d_pair_v);
int pmk_query_combine_dev(pmk_handle* h, const double* d_pair_u, const double* d_pair_v, double* dYq, double* dVq);
/* The pair kernel's operands (P = inv(L) for the explicit-inverse solver, M_IJ = L_IJ inv(L_JJ) for substitution) are built by
 * the first variance query after a fit; pmk_build_M builds them now (so that the cost is the fit's, not the first query's). */
int pmk_build_M(pmk_handle* h);

/* ---- sub-tree ownership: building blocks (pmk_multi below is the ready-made single-box form) ------------------------------- */
/* Leaves are independent (reference: one GP per leaf, mixtureGP.jl:92-115).  A handle may own a contiguous range of a larger
 * model's leaves -- a sub-tree of the BSP: pmk_set_leaf_base(h, first leaf 0-based, leaves of the whole model) BEFORE pmk_fit,
 * which then receives only the owned leaves' inputs.  Leaf ids stay global everywhere (pmk_get_alpha(h, leaf) ...).  Every
 * handle holds the (tiny) whole tree, so any of them can PLAN any slice of the queries; the (query, leaf) pairs of a plan are
 * then answered by the owners of their leaves:
 *   pmk_query_plan_segments   : the plan's pairs sorted by leaf form one contiguous segment per owner; seg_off[o] .. seg_off[o+1]
 *                               (n_owners + 1 entries out) for owners holding leaves owner_first_leaf[o] .. owner_first_leaf[o+1]
 *   pmk_query_plan_pack_dev   : the sorted pairs as what travels: query point (D doubles, point-major) and 1-based leaf id
 *   pmk_query_pairs_routed_dev: OWNER side -- R pairs (points + leaf ids, any order, own leaves only) -> u, v in the same order
 *   pmk_query_plan_unpack_dev : the owners' answers, concatenated in sorted order, back to the plan's pair order
 *   pmk_query_set_flags + pmk_query_combine_dev finish the query on the planning handle.
 * Results are bit-identical to a single handle owning every leaf: a pair's u, v depend on its leaf and point only. */
int pmk_set_leaf_base(pmk_handle* h, int64_t leaf_base, int64_t total_leaves);
int pmk_query_plan_segments(pmk_handle* h, int n_owners, const int64_t* owner_first_leaf, int64_t* seg_off);
int pmk_query_plan_pack_dev(pmk_handle* h, double* d_X_sorted, int32_t* d_leaf_sorted);
int pmk_query_pairs_routed_dev(pmk_handle* h, int64_t R, const double* d_X, const int32_t* d_leaf, int flags, double* d_u, double* d_v);
int pmk_query_plan_unpack_dev(pmk_handle* h, const double* d_u_sorted, const double* d_v_sorted, double* d_pair_u, double* d_pair_v);
int pmk_query_set_flags(pmk_handle* h, int flags);

/* ---- multi-GPU: one model over the GPUs of one box (north star: "patches are partitioned across the 8 GPUs ... by a
 * BSP-leaf-to-rank map") ---------------------------------------------------------------------------------------------------- */
/* One process, one host thread and one stream set per GPU, CUDA peer copies over NVLink between them.  Rank r OWNS a
 * contiguous leaf range (pmk_multi_owned_range: ranges of equal cost; with equal leaves the n sub-trees below the top log2(n)
 * levels of the BSP) and keeps X, alpha, L and
 * the query operand for those leaves only; nothing is replicated but the tree.  fitmixtureGP! (mixtureGP.jl:70) = every rank
 * fits its leaves, no exchange.  querymixtureGP! (mixtureGP.jl:159) = every rank plans a contiguous slice of the queries
 * (pmk_multi_query_range), the pairs travel to the owners of their leaves (D doubles + 4 bytes each), the owners run the fused
 * pair kernel, u and v travel back (16 bytes per pair), the planning rank combines in the reference's order and its slice of
 * Yq, Vq goes straight into the caller's host arrays.  Results are bit-identical to one GPU.
 * device_ids: n_devices CUDA ordinals (NULL = 0 .. n_devices-1); an ordinal may repeat (several ranks on one GPU: tests).
 * The *_staged forms keep inputs and results resident in HBM between calls (timing without host copies):
 *   pmk_multi_stage_training + pmk_multi_fit_staged = pmk_multi_fit;
 *   pmk_multi_stage_queries + pmk_multi_query_staged + pmk_multi_fetch_results = pmk_multi_query. */
typedef struct pmk_multi pmk_multi;
int pmk_multi_create(pmk_multi** out, int n_devices, const int* device_ids);
void pmk_multi_destroy(pmk_multi* m);
const char* pmk_multi_last_error(const pmk_multi* m);     /* m may be NULL: error of a failed pmk_multi_create */
int pmk_multi_size(const pmk_multi* m);
/* the equal-count split of [0, total) over n ranks (host only, no GPU needed): [first, first + count).  pmk_multi slices the
 * QUERIES this way (pmk_multi_query_range); pmk_multi_leaf_range is the same split for a host layer that deals leaves itself
 * (patchmixturekriging_b200/sharding.py).  pmk_multi's own leaf -> rank map is cost-balanced: see pmk_multi_owned_range. */
int pmk_multi_leaf_range(int n_ranks, int64_t n_leaves, int rank, int64_t* first, int64_t* count);
int pmk_multi_query_range(int n_ranks, int64_t Nq, int rank, int64_t* first, int64_t* count);
/* the leaves rank owns, 0-based [first, first + count), valid once the training data is staged (pmk_multi_stage_training /
 * pmk_multi_fit): contiguous ranges of nearly equal cost sum(n^3) -- a leaf's share of the factorisation, the operand build
 * and, for queries spread like the training points, the pair kernel */
int pmk_multi_owned_range(const pmk_multi* m, int rank, int64_t* first, int64_t* count);
/* the same deal as a host-only function (no GPU needed), for a host layer that runs one process per GPU and wants pmk_multi's map:
 * first[0 .. n_ranks] = the boundaries of n_ranks contiguous leaf ranges of nearly equal sum(n^3); n_leaves >= n_ranks */
int pmk_multi_balanced_ranges(int n_ranks, int64_t n_leaves, const int64_t* leaf_off, int64_t* first);
/* rank's own handle, for inspection of the leaves it owns (pmk_get_L, pmk_get_alpha, pmk_condition_estimate ...) */
int pmk_multi_handle(pmk_multi* m, int rank, pmk_handle** h);
int pmk_multi_set_option(pmk_multi* m, int option, int64_t value);
int pmk_multi_fit(pmk_multi* m, int D, int64_t n_leaves, const int64_t* leaf_off, const double* X_packed, const double* y_packed,
                  int kernel_id, const double* kparams, int nparams, double sigma2, int64_t* bad_leaf, int* info);
int pmk_multi_set_tree(pmk_multi* m, int D, int levels, const double* hp_v, const double* hp_c);
int pmk_multi_query(pmk_multi* m, int64_t Nq, const double* Xq, double radius, double delta, int wkernel_id, const double* wparams,
                    int nw, int flags, double* Yq, double* Vq);
int pmk_multi_stage_training(pmk_multi* m, int D, int64_t n_leaves, const int64_t* leaf_off, const double* X_packed,
                             const double* y_packed);
int pmk_multi_fit_staged(pmk_multi* m, int kernel_id, const double* kparams, int nparams, double sigma2, int64_t* bad_leaf, int* info);
int pmk_multi_stage_queries(pmk_multi* m, int64_t Nq, const double* Xq);
int pmk_multi_query_staged(pmk_multi* m, double radius, double delta, int wkernel_id, const double* wparams, int nw, int flags);
int pmk_multi_fetch_results(pmk_multi* m, double* Yq, double* Vq);
/* (query, leaf) pairs of the last query per leaf, summed over the ranks (n_leaves entries) */
int pmk_multi_leaf_pairs(pmk_multi* m, int64_t* pairs_per_leaf);
/* Device times of the last staged fit / query in milliseconds, CUDA events on each rank's stream, MAXIMUM over the ranks:
 * a phase's time on a rank runs from the rank's first event of the call to the end of the phase, waits for peers included. */
enum {
  PMK_MT_FIT = 0,          /* whole fit: pack + Gram + Cholesky + alpha + the query operand                  */
  PMK_MT_QUERY = 1,        /* whole query: plan .. combine                                                   */
  PMK_MT_Q_PLAN = 2,       /* home leaves, neighbours, pair list (per-rank duration, max over ranks)         */
  PMK_MT_Q_ROUTE = 3,      /* pack + peer copies of the pairs to their owners                                */
  PMK_MT_Q_PAIRS = 4,      /* owners: binning + fused pair kernel                                            */
  PMK_MT_Q_RETURN = 5,     /* peer copies of u, v back + unpack + combine                                    */
  PMK_MT_COUNT = 8
};
int pmk_multi_get_timings(pmk_multi* m, double* ms /* PMK_MT_COUNT */, double* per_rank_ms /* n x PMK_T_COUNT, or NULL */);
int64_t pmk_multi_launch_count(const pmk_multi* m);

/* ---- options ----------------------------------------------------------------------------- */
/* PMK_OPT_FULL_HYPERPLANE_SCAN: 1 = findneighbourpartitions scans ALL hyperplanes per query exactly as the
 * reference loop does (mixtureGP.jl:354); 0 (default) = exact per-leaf candidate lists (same result). */
/* PMK_OPT_QUERY_SOLVER: how queryinner!'s v = L \ kq (mixtureGP.jl:311) is carried out for a tile of queries:
 *  -1 (default) = by conditioning: 0 unless the fit's lower bound of cond(K + sigma2 I), (max diag L / min diag L)^2 over the
 *                 leaves (pmk_condition_estimate), reaches 1e4 -- then 1.  Measured against the reference's dtrsv
 *                 (profiles/parity_r02.json): the explicit inverse stays below 1e-9 of the variance up to cond ~ 1e5 and
 *                 reaches 3e-9 at cond 3e6, substitution stays at the level at which dtrsv and dtrsm differ from each other;
 *   0           = s = P kq with P = inv(L) formed once per fit (PMK_OPT_INVERSE_BUILDER), as a ROW-PANEL product: the
 *                 cross-covariance tile is evaluated once into shared memory, every warp streams its own rows of P and
 *                 keeps only ||s||^2 -- no dependency between warps, so the tensor pipe never waits.  Every kernel
 *                 function (the squared exponential with an inlined table-driven exp).  Measured vs dtrsv: <= 1.1e-11 at
 *                 sigma2 = 1e-3, 2.2e-10 / 7.1e-10 at sigma2 = 1e-5 (squared exponential cond 3e6 / Spline34 cond 1.5e7);
 *   1           = blocked forward substitution with 32x32 diagonal-block inverses (closest to dtrsv; what -1 resolves to for
 *                 ill-conditioned models; also serves mean-only queries). */
/* PMK_OPT_INVERSE_BUILDER: how P = inv(L) is formed (once per fit) for the explicit-inverse solvers:
 *   0 (default) = recursive doubling on the packed tiles, P21 = -inv(B) C inv(A), every flop a DMMA GEMM (pmk_invert.cu);
 *   1           = the substitution pair kernel run on identity right-hand sides (round-1 builder). */
/* PMK_OPT_ALPHA_REFINE: one step of iterative refinement of alpha = (K + sigma2 I)^-1 y after the Cholesky solve (the reference
 * solves U\y by LU, mixtureGP.jl:106): -1 (default) = for models flagged by the same conditioning estimate, 0 = never,
 * 1 = always.  Set before pmk_fit. */
/* PMK_OPT_CHOL_VARIANT: the batched Cholesky of the fit: -1 (default) = by leaf size -- leaves of 768 padded rows and more advance
 * 64-column panel by panel through the level-synchronous kernels (the serial factorisation of the 64x64 diagonal blocks in one
 * launch, k_chol_factor64, the DMMA panel updates with TMA-staged operands in the next, k_chol_panel64), smaller leaves take one CTA
 * per leaf running its panels to the end (k_chol); 0 = level-synchronous for every leaf; 1 = one CTA per leaf for every leaf.  Same L to rounding; the default's
 * choice depends on the leaf alone, so a leaf's factor does not depend on what else is fitted with it. */
/* PMK_OPT_GRAM_FAST_EXP: pmk_gram / pmk_cross_gram with the squared exponential evaluated as exp(-eps_sq |x - z|^2) by the
 * table-driven exp of the fit and query kernels instead of the reference's sqrt, re-square and libm exp (kernel.jl:277-287,
 * 350-357): 0 (default) = the reference's operation order (entries within 5e-15 of the oracle), 1 = fast: the exp within
 * 1.3 ulp, the value within (3 + 6 |log K|) ulp of the reference-order entry because the argument is rounded differently
 * (< 1e-13 relative wherever K > 1e-30); a third of the FP64 work of a kernel that is otherwise bound by it, not by its 8 n^2 bytes. */
enum { PMK_OPT_FULL_HYPERPLANE_SCAN = 1, PMK_OPT_QUERY_SOLVER = 2, PMK_OPT_INVERSE_BUILDER = 3, PMK_OPT_ALPHA_REFINE = 4,
       PMK_OPT_CHOL_VARIANT = 5, PMK_OPT_GRAM_FAST_EXP = 6 };
int pmk_set_option(pmk_handle* h, int option, int64_t value);
/* lower bound of the worst leaf's cond(K + sigma2 I) from the last fit, and the query solver PMK_OPT_QUERY_SOLVER = -1 resolves to */
int pmk_condition_estimate(pmk_handle* h, double* cond_lower_bound, int* solver_in_use);

/* ---- setuppartition on the device ------------------------------------------ */
/* setuppartition(X, levels) (reference src/patchwork/partition.jl:106-129; gethyperplane :86-100, splitpoints :64-83,
 * createchildren :166-217) with every O(N) step on the GPU, one tree level per call pair:
 *
 *   pmk_partition_begin(h, D, N, X, levels)            X: N x D point-major host array (array2matrix(X), utilities.jl:25-36)
 *   for depth = 0 .. levels-2:
 *     pmk_partition_level_z(h, depth, z)                z[j*D..] = X_j[1] - mean(X_j) for the 2^depth nodes of this depth,
 *                                                       left to right (partition.jl:89-90; Base's pairwise summation order)
 *     v_j = V[:,1] of svd(z_j')                         ON THE HOST, by the caller's own LinearAlgebra (partition.jl:93-94):
 *                                                       the bits of v are LAPACK dgesdd's and no restatement can promise
 *                                                       them, so the reference's own call stays where it is
 *     pmk_partition_level_split(h, depth, v, c)         f = dot(v_j, x), c_j = median(f) (returned), left iff f < c_j; every
 *                                                       node's points are split in place, order kept (partition.jl:64-83)
 *   pmk_partition_fetch(h, leaf_off, inds)              leaves in AbstractTrees.Leaves order (= labelleafnodes' numbering,
 *                                                       partition.jl:131-159): leaf p owns inds[leaf_off[p] .. leaf_off[p+1]),
 *                                                       ascending 1-based global ids = X_parts_inds; 2^(levels-1)+1 offsets
 *
 * The node at depth d with left-to-right index j has pre-order (fetchhyperplanes) index sum over its path bits b_i
 * (MSB first) of (b_i ? 2^(levels-2-i) : 1).  Bit-exact contract: z, c, leaf_off, inds equal the oracle's for the same v.
 * Errors: PMK_ERR_STATE (call order), PMK_ERR_ARG (a child without points: the reference fails in mean() there).
 * pmk_partition_sum_plan is host-only (no GPU needed): the sequential blocks (start, length, depth in the halving tree)
 * Base.mapreduce_impl(+, A, 1, n, 1024) sums a range of n elements in -- the order k_part_block_sums / k_part_node_z follow. */
int pmk_partition_begin(pmk_handle* h, int D, int64_t N, const double* X, int levels);
int pmk_partition_level_z(pmk_handle* h, int depth, double* z_out);
int pmk_partition_level_split(pmk_handle* h, int depth, const double* v, double* c_out);
int pmk_partition_fetch(pmk_handle* h, int64_t* leaf_off_out, int32_t* inds_out);
int pmk_partition_sum_plan(int64_t n, int64_t max_blocks, int64_t* blk_start, int32_t* blk_len, int32_t* blk_depth, int64_t* n_blocks);

/* ---- checkpoint -------------------------------------------------------------------------- */
/* The reference keeps a fitted MixtureGPType in memory only (src/RKHS/mixtureGP.jl:38-66; no serialisation anywhere in
 * src/).  pmk_save_model writes the handle's fitted model -- padded training inputs, y, alpha (c_set), the packed factors L
 * (L_set) with their diagonal-block inverses, kernel id/parameter, sigma2 and, when set, the flattened tree -- to one flat
 * binary file in the handle's own HBM layout (little-endian, "PMKB200" magic, version 1).  pmk_load_model lays the same
 * model out in a handle (any previous model of that handle is replaced, its fit range reset to all leaves) and copies the
 * buffers back: queries on the loaded handle return bit-identical results to the handle that saved.  The pair-kernel
 * operand P = inv(L) is not stored; the first variance query rebuilds it, as after pmk_fit.
 * Errors: PMK_ERR_STATE (save before fit), PMK_ERR_ARG (cannot open / truncated / not a model file),
 * PMK_ERR_UNSUPPORTED (file version or layout of another library version). */
/* What a host layer needs to rebuild its view of a loaded model: dimensions and hyper-parameters (any out pointer may be
 * NULL; *levels = 0 when no tree is set), a leaf's training inputs (n x D point-major = X_parts[leaf], mixtureGP.jl:40),
 * and the hyperplanes in pmk_set_tree's layout (fetchhyperplanes order, mixtureGP.jl:322-334). */
int pmk_model_info(pmk_handle* h, int* D, int64_t* n_leaves, int* kernel_id, double* kparam, double* sigma2, int* levels);
int pmk_get_X(pmk_handle* h, int64_t leaf, double* out);
int pmk_get_tree(pmk_handle* h, double* hp_v, double* hp_c);
int pmk_save_model(pmk_handle* h, const char* path);
int pmk_load_model(pmk_handle* h, const char* path);

/* ---- instrumentation --------------------------------------------------------------------- */
int pmk_get_timings(pmk_handle* h, double* ms /* PMK_T_COUNT entries */);
/* cycle counters of the fit kernel, summed over CTAs (warp 0): total, gram-init, update loop, diagonal factor,
 * panel solve, barrier wait, #CTAs, reserved.  flags: bit0 = zero them after reading; bit1 = read the pair kernel's
 * counters instead (total, cross-covariance init, publish+barrier, diagonal solve+barrier, update, #CTAs). */
int pmk_debug_counters(pmk_handle* h, uint64_t* out8, int flags);
/* The recursion plan pmk_invert.cu uses to form P = inv(L) for a leaf of n_blocks 32-row blocks (host only, no GPU needed):
 * nodes4[4k .. 4k+3] = {lo, mid, hi, height} of node k, in launch order (ascending height).  A node inverts the block range
 * [lo, hi) from its children [lo, mid) and [mid, hi): P[mid:hi, lo:mid] = -P[mid:hi, mid:hi] L[mid:hi, lo:mid] P[lo:mid, lo:mid]. */
int pmk_inverse_plan(int n_blocks, int max_nodes, int16_t* nodes4, int* n_nodes);
/* The FP64 tensor-pipe rate of the handle's GPU, measured now: a register-resident loop of independent mma.sync.m8n8k4.f64
 * (SASS DMMA.8x8x4) chains on every SM, best of three ~2 ms launches, in TFLOP/s.  bench.py quotes every FP64 fraction
 * against this number (MEASURED_PEAKS.json carries no FP64 entry). */
int pmk_measure_fp64_peak(pmk_handle* h, double* dmma_tflops);
/* number of kernel launches issued by this handle since creation */
int64_t pmk_launch_count(const pmk_handle* h);
/* stream the handle launches on, as a cudaStream_t cast to void* (for event timing by the caller) */
void* pmk_stream(pmk_handle* h);
int pmk_synchronize(pmk_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* PMK_H */
