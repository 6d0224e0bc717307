// K3 dispatch: size classes and the per-D launchers (kernel in pmk_query_trsm.cuh).
#include "pmk_internal.cuh"

namespace pmk {

void read_query_cycles_d1(unsigned long long* out, bool reset);
void read_query_cycles_d2(unsigned long long* out, bool reset);
void read_query_cycles_d3(unsigned long long* out, bool reset);
void read_query_cycles(int D, unsigned long long* out, bool reset) {
  if (D == 1) read_query_cycles_d1(out, reset);
  else if (D == 2) read_query_cycles_d2(out, reset);
  else read_query_cycles_d3(out, reset);
}

__global__ void k_class_tiles(PairWork w, int mq, int32_t* __restrict__ tiles) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= w.n_class_leaves) return;
  const int64_t gl = w.leaf_base + w.class_leaves[s];
  const int64_t cnt = w.leaf_pair_start[gl + 1] - w.leaf_pair_start[gl];
  tiles[s] = (int32_t)((cnt + mq - 1) / mq);
}


void launch_pairs_d1(int, unsigned, const LeafTable&, const PairWork&, const QueryPlan&, KParams, int, double*, double*, cudaStream_t);
void launch_pairs_d2(int, unsigned, const LeafTable&, const PairWork&, const QueryPlan&, KParams, int, double*, double*, cudaStream_t);
void launch_pairs_d3(int, unsigned, const LeafTable&, const PairWork&, const QueryPlan&, KParams, int, double*, double*, cudaStream_t);

// size classes: n_pad <= 8 * NT * NW
int query_class_of(int npad) { return npad <= 512 ? 0 : (npad <= 768 ? 1 : (npad <= 1024 ? 2 : (npad <= 1536 ? 3 : 4))); }
int query_class_mq(int cls) { return cls == 0 ? 32 : (cls == 1 ? 24 : (cls == 2 ? 16 : 8)); }

void launch_query_pairs(int D, int cls, unsigned grid, const LeafTable& lt, const PairWork& w, const QueryPlan& q,
                        KParams kp, int mean_only, double* pu, double* pv, cudaStream_t s) {
  if (grid == 0) return;
  switch (D) {
    case 1: launch_pairs_d1(cls, grid, lt, w, q, kp, mean_only, pu, pv, s); break;
    case 2: launch_pairs_d2(cls, grid, lt, w, q, kp, mean_only, pu, pv, s); break;
    case 3: launch_pairs_d3(cls, grid, lt, w, q, kp, mean_only, pu, pv, s); break;
    default: break;
  }
}

bool launch_rowp_d1(int, const LeafTable&, const PairWork&, const QueryPlan&, KParams, int, int, double*, double*, cudaStream_t);
bool launch_rowp_d2(int, const LeafTable&, const PairWork&, const QueryPlan&, KParams, int, int, double*, double*, cudaStream_t);
bool launch_rowp_d3(int, const LeafTable&, const PairWork&, const QueryPlan&, KParams, int, int, double*, double*, cudaStream_t);

// explicit-inverse pair kernel, row-panel product (pmk_query_rowp.cuh); persistent, reads its tile count on the device
bool launch_query_rowp(int D, int cls, const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp, int flags,
                       int npmax, double* pu, double* pv, cudaStream_t s) {
  switch (D) {
    case 1: return launch_rowp_d1(cls, lt, w, q, kp, flags, npmax, pu, pv, s);
    case 2: return launch_rowp_d2(cls, lt, w, q, kp, flags, npmax, pu, pv, s);
    case 3: return launch_rowp_d3(cls, lt, w, q, kp, flags, npmax, pu, pv, s);
    default: return false;
  }
}

void launch_class_tiles(const PairWork& w, int mq, int32_t* tiles, cudaStream_t s) {
  if (w.n_class_leaves == 0) return;
  k_class_tiles<<<(w.n_class_leaves + 255) / 256, 256, 0, s>>>(w, mq, tiles);
}

}  // namespace pmk
