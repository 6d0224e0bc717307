// K3 instantiations for D = 1 (see pmk_query_trsm.cuh, pmk_query_rowp.cuh)
#include "pmk_query_rowp.cuh"
namespace pmk {
void launch_pairs_d1(int cls, unsigned grid, const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp,
                      int mean_only, double* pu, double* pv, cudaStream_t s) {
  launch_pairs_d<1>(cls, grid, lt, w, q, kp, mean_only, pu, pv, s);
}
bool launch_rowp_d1(int cls, const LeafTable& lt, const PairWork& w, const QueryPlan& q, KParams kp, int flags, int npmax,
                    double* pu, double* pv, cudaStream_t s) {
  return launch_rowp_d<1>(cls, lt, w, q, kp, flags, npmax, pu, pv, s);
}
void read_query_cycles_d1(unsigned long long* out, bool reset) { read_query_cycles_tu(out, reset); }
}  // namespace pmk
