// pmk_multi: one mixture-GP model over the GPUs of one box by SUB-TREE OWNERSHIP (include/pmk.h).
//
// The reference is single-threaded (fitmixtureGP! loops over the leaves, mixtureGP.jl:92; querymixtureGP! over the queries,
// :203); its leaves are independent GPs, so rank r owns the contiguous leaf range [n_leaves r / n, n_leaves (r+1) / n) -- with
// n a power of two these are the n sub-trees below the top log2(n) levels of the BSP -- and nothing but the tree is replicated.
//   fit   : every rank fits its own leaves (and builds its query operand); no exchange.
//   query : every rank PLANS a contiguous slice of the queries (home leaf, neighbours, weights: needs only the tree), then the
//           (query, leaf) pairs travel to the owners of their leaves -- the plan's leaf-sorted pair list is one contiguous
//           segment per owner, pulled by the owner with one peer copy per source (D doubles + 4 bytes per pair over NVLink) --,
//           the owners run the fused pair kernel on what they received, the planners pull u, v back (16 bytes per pair),
//           combine in the reference's slot order and copy their slice of Yq, Vq into the caller's host arrays.
// One host thread per rank (CUDA's current device is per thread), host barriers at the two exchange points, every transfer a
// contiguous cudaMemcpyPeerAsync on the PULLING rank's stream -- so the only cross-device ordering needed is "the producer
// has synchronised before the barrier".  Written against the single-GPU C ABI (include/pmk.h) only.
#include <algorithm>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/pmk.h"

namespace {

thread_local std::string g_multi_create_error;

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
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

// all ranks meet; the call returns true iff EVERY rank arrived with ok == true (a failed rank keeps meeting the others,
// so nobody waits for ever)
class Meet {
 public:
  explicit Meet(int n) : n_(n) {}
  bool sync(bool ok) {
    std::unique_lock<std::mutex> lk(mu_);
    if (!ok) bad_ = true;
    const unsigned long long gen = gen_;
    if (++arrived_ == n_) {
      arrived_ = 0;
      result_ = !bad_;
      ++gen_;
      cv_.notify_all();
      return result_;
    }
    cv_.wait(lk, [&] { return gen_ != gen; });
    return result_;
  }
  void reset() { bad_ = false; }

 private:
  int n_, arrived_ = 0;
  unsigned long long gen_ = 0;
  bool bad_ = false, result_ = true;
  std::mutex mu_;
  std::condition_variable cv_;
};

enum { EV_START = 0, EV_PLAN, EV_ROUTE, EV_PAIRS, EV_END, EV_FIT0, EV_FIT1, EV_COUNT };

struct Rank {
  int device = 0;
  pmk_handle* h = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[EV_COUNT] = {};
  // leaves [leaf_first, leaf_first + leaf_count) and their staged inputs
  int64_t leaf_first = 0, leaf_count = 0;
  std::vector<int64_t> leaf_off;       // leaf_count + 1, rebased to 0
  Buf dX, dy;
  // query slice [q_first, q_first + q_count)
  int64_t q_first = 0, q_count = 0;
  Buf dXq, dYq, dVq;
  // as PLANNER: the plan's pairs in leaf-sorted order (tx_*), the owners' answers in the same order (tu, tv), in pair order (pu, pv)
  int64_t n_pairs = 0;
  std::vector<int64_t> seg;            // n + 1: this plan's segment per owner
  Buf tx_X, tx_leaf, tu, tv, pu, pv;
  // as OWNER: what the planners sent (rx_*), per-planner offsets, answers
  int64_t n_rx = 0;
  std::vector<int64_t> rx_off;         // n + 1
  Buf rx_X, rx_leaf, rx_u, rx_v;
  int rc = PMK_OK;
  std::string err;
  double ms[PMK_MT_COUNT] = {};
};

}  // namespace

struct pmk_multi {
  int n = 0;
  std::vector<Rank> rk;
  std::string err;
  Meet* meet = nullptr;
  // one persistent host thread per rank (n > 1): a call posts a job, every worker runs it for its rank, the caller waits.
  // Creating the threads per call cost ~0.3 ms per call at 8 ranks, five calls per fit + query step.
  std::vector<std::thread> workers;
  std::mutex jm;
  std::condition_variable jcv, dcv;
  std::function<void(int)> job;
  unsigned long long job_gen = 0;
  int pending = 0;
  bool stop = false;
  int D = 0;
  int64_t n_leaves = 0;
  bool staged_training = false, fitted = false, tree_set = false, staged_queries = false, results_ready = false;
  int64_t Nq = 0;
  int last_flags = 0;
  double ms[PMK_MT_COUNT] = {};
};

namespace {

int mfail(pmk_multi* m, int code, const char* fmt, ...) {
  char buf[640];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (m) m->err = buf; else g_multi_create_error = buf;
  return code;
}

// error of rank r from a single-GPU call
int rfail(Rank& r, int code) {
  r.rc = code;
  r.err = pmk_last_error(r.h);
  return code;
}
int rcuda(Rank& r, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return PMK_OK;
  r.rc = PMK_ERR_CUDA;
  r.err = std::string(what) + " failed: " + cudaGetErrorString(e);
  return PMK_ERR_CUDA;
}
#define RK(r, call)                                       \
  do {                                                    \
    const int rc_ = (call);                               \
    if (rc_ != PMK_OK) return rfail(r, rc_);              \
  } while (0)
#define RC(r, expr)                                       \
  do {                                                    \
    if (int rc_ = rcuda(r, (expr), #expr)) return rc_;    \
  } while (0)

// the calling thread's current device is the caller's business (torch tracks it): put it back on the way out
struct DeviceGuard {
  int dev = -1;
  DeviceGuard() {
    if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
  }
  ~DeviceGuard() {
    if (dev >= 0) cudaSetDevice(dev);
  }
};

// run fn(rank) on one host thread per rank; first failure (lowest rank) becomes the call's status
template <class F>
int run_ranks(pmk_multi* m, F&& fn) {
  for (Rank& r : m->rk) {
    r.rc = PMK_OK;
    r.err.clear();
  }
  m->meet->reset();
  DeviceGuard guard;
  if (m->n == 1) {
    cudaSetDevice(m->rk[0].device);
    fn(0);
  } else {
    std::unique_lock<std::mutex> lk(m->jm);
    m->job = [&](int i) { fn(i); };
    m->pending = m->n;
    ++m->job_gen;
    m->jcv.notify_all();
    m->dcv.wait(lk, [&] { return m->pending == 0; });
    m->job = nullptr;
  }
  for (int i = 0; i < m->n; ++i)
    if (m->rk[i].rc != PMK_OK) {
      m->err = "rank " + std::to_string(i) + " (device " + std::to_string(m->rk[i].device) + "): " + m->rk[i].err;
      return m->rk[i].rc;
    }
  return PMK_OK;
}

void worker_main(pmk_multi* m, int i) {
  cudaSetDevice(m->rk[i].device);
  unsigned long long seen = 0;
  for (;;) {
    std::function<void(int)> job;
    {
      std::unique_lock<std::mutex> lk(m->jm);
      m->jcv.wait(lk, [&] { return m->stop || m->job_gen != seen; });
      if (m->stop) return;
      seen = m->job_gen;
      job = m->job;
    }
    job(i);
    {
      std::lock_guard<std::mutex> lk(m->jm);
      if (--m->pending == 0) m->dcv.notify_all();
    }
  }
}

void range_of(int n, int64_t total, int r, int64_t* first, int64_t* count) {
  const int64_t a = total * r / n, b = total * (r + 1) / n;
  *first = a;
  *count = b - a;
}

// Ownership: contiguous leaf ranges of (nearly) equal COST, cost of a leaf = n^3 -- its share of the factorisation and of the
// operand build, and with queries spread like the training points also of the pair kernel (pairs ~ n, flops per pair ~ n^2).
// Boundary i is the leaf index whose cost prefix is nearest to i / n of the total, every rank keeping at least one leaf.
// (Equal leaf COUNTS left the slowest of 8 owners 6 % behind the mean on C4: leaf sizes 700 .. 1400.)
void balanced_bounds(int n, int64_t n_leaves, const int64_t* leaf_off, int64_t* bnd /* n + 1 */) {
  std::vector<double> pre((size_t)n_leaves + 1, 0.0);
  for (int64_t p = 0; p < n_leaves; ++p) {
    const double np = (double)(leaf_off[p + 1] - leaf_off[p]);
    pre[p + 1] = pre[p] + np * np * np;
  }
  bnd[0] = 0;
  bnd[n] = n_leaves;
  for (int i = 1; i < n; ++i) {
    const double target = pre[n_leaves] * (double)i / (double)n;
    int64_t k = std::lower_bound(pre.begin(), pre.end(), target) - pre.begin();
    if (k > 0 && target - pre[k - 1] < pre[k] - target) --k;
    k = std::max<int64_t>(k, bnd[i - 1] + 1);
    k = std::min<int64_t>(k, n_leaves - (n - i));
    bnd[i] = k;
  }
}

float elapsed(cudaEvent_t a, cudaEvent_t b) {
  float t = 0.f;
  return cudaEventElapsedTime(&t, a, b) == cudaSuccess ? t : 0.f;
}

}  // namespace

extern "C" {

int pmk_multi_leaf_range(int n_ranks, int64_t n_leaves, int rank, int64_t* first, int64_t* count) {
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || n_leaves < 0 || !first || !count) return PMK_ERR_ARG;
  range_of(n_ranks, n_leaves, rank, first, count);
  return PMK_OK;
}

int pmk_multi_query_range(int n_ranks, int64_t Nq, int rank, int64_t* first, int64_t* count) {
  return pmk_multi_leaf_range(n_ranks, Nq, rank, first, count);
}


int pmk_multi_balanced_ranges(int n_ranks, int64_t n_leaves, const int64_t* leaf_off, int64_t* first) {
  if (n_ranks < 1 || n_leaves < n_ranks || !leaf_off || !first) return PMK_ERR_ARG;
  for (int64_t p = 0; p < n_leaves; ++p)
    if (leaf_off[p + 1] < leaf_off[p]) return PMK_ERR_ARG;
  balanced_bounds(n_ranks, n_leaves, leaf_off, first);
  return PMK_OK;
}

const char* pmk_multi_last_error(const pmk_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

int pmk_multi_size(const pmk_multi* m) { return m ? m->n : 0; }

void pmk_multi_destroy(pmk_multi* m) {
  if (!m) return;
  DeviceGuard guard;
  if (!m->workers.empty()) {
    {
      std::lock_guard<std::mutex> lk(m->jm);
      m->stop = true;
    }
    m->jcv.notify_all();
    for (std::thread& t : m->workers) t.join();
    m->workers.clear();
  }
  for (Rank& r : m->rk) {
    if (!r.h) continue;             // a rank that was never created (pmk_multi_create failed on the way) owns nothing
    cudaSetDevice(r.device);
    if (r.stream) cudaStreamSynchronize(r.stream);
    Buf* bufs[] = {&r.dX, &r.dy, &r.dXq, &r.dYq, &r.dVq, &r.tx_X, &r.tx_leaf, &r.tu, &r.tv, &r.pu, &r.pv, &r.rx_X, &r.rx_leaf, &r.rx_u, &r.rx_v};
    for (Buf* b : bufs) b->release();
    for (cudaEvent_t e : r.ev)
      if (e) cudaEventDestroy(e);
    if (r.h) pmk_destroy(r.h);
  }
  delete m->meet;
  delete m;
}

int pmk_multi_create(pmk_multi** out, int n_devices, const int* device_ids) {
  if (!out) return mfail(nullptr, PMK_ERR_ARG, "pmk_multi_create: out is NULL");
  *out = nullptr;
  if (n_devices < 1 || n_devices > 64) return mfail(nullptr, PMK_ERR_ARG, "pmk_multi_create: n_devices=%d out of range [1, 64]", n_devices);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return mfail(nullptr, PMK_ERR_CUDA, "pmk_multi_create: no CUDA device (%s); libpmk_b200 has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  DeviceGuard guard;
  pmk_multi* m = new (std::nothrow) pmk_multi();
  if (!m) return mfail(nullptr, PMK_ERR_CUDA, "out of host memory");
  m->n = n_devices;
  m->rk.resize(n_devices);
  m->meet = new Meet(n_devices);
  for (int i = 0; i < n_devices; ++i) {
    Rank& r = m->rk[i];
    r.device = device_ids ? device_ids[i] : i;
    if (r.device < 0 || r.device >= ndev) {
      const int bad = r.device;
      pmk_multi_destroy(m);
      return mfail(nullptr, PMK_ERR_ARG, "pmk_multi_create: device %d out of range [0,%d)", bad, ndev);
    }
    const int rc = pmk_create(&r.h, r.device);
    if (rc != PMK_OK) {
      const std::string msg = pmk_last_error(nullptr);
      pmk_multi_destroy(m);
      return mfail(nullptr, rc, "pmk_multi_create: rank %d: %s", i, msg.c_str());
    }
    r.stream = (cudaStream_t)pmk_stream(r.h);
    cudaSetDevice(r.device);
    for (cudaEvent_t& ev : r.ev) cudaEventCreate(&ev);
  }
  // peer access between every pair of distinct devices: the segment copies then go GPU to GPU over NVLink
  for (int i = 0; i < n_devices; ++i) {
    cudaSetDevice(m->rk[i].device);
    for (int j = 0; j < n_devices; ++j) {
      if (m->rk[j].device == m->rk[i].device) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, m->rk[i].device, m->rk[j].device);
      if (can) {
        e = cudaDeviceEnablePeerAccess(m->rk[j].device, 0);
        if (e != cudaSuccess) cudaGetLastError();     // already enabled (another pmk_multi, torch ...): fine
      }
    }
  }
  if (n_devices > 1)
    for (int i = 0; i < n_devices; ++i) m->workers.emplace_back(worker_main, m, i);
  *out = m;
  return PMK_OK;
}

int pmk_multi_owned_range(const pmk_multi* m, int rank, int64_t* first, int64_t* count) {
  if (!m || rank < 0 || rank >= m->n || !first || !count) return PMK_ERR_ARG;
  if (!m->staged_training) return PMK_ERR_STATE;
  *first = m->rk[rank].leaf_first;
  *count = m->rk[rank].leaf_count;
  return PMK_OK;
}

int pmk_multi_handle(pmk_multi* m, int rank, pmk_handle** h) {
  if (!m || !h || rank < 0 || rank >= m->n) return PMK_ERR_ARG;
  *h = m->rk[rank].h;
  return PMK_OK;
}

int pmk_multi_set_option(pmk_multi* m, int option, int64_t value) {
  if (!m) return PMK_ERR_ARG;
  for (Rank& r : m->rk) {
    const int rc = pmk_set_option(r.h, option, value);
    if (rc != PMK_OK) return mfail(m, rc, "%s", pmk_last_error(r.h));
  }
  return PMK_OK;
}

// ---------------------------------------------------------------------------------------------
namespace {

int check_training(pmk_multi* m, int D, int64_t n_leaves, const int64_t* leaf_off, const double* X, const double* y) {
  if (D < 1 || D > PMK_MAX_DIM) return mfail(m, PMK_ERR_UNSUPPORTED, "D=%d unsupported (1..%d)", D, PMK_MAX_DIM);
  if (!leaf_off || !X || !y) return mfail(m, PMK_ERR_ARG, "NULL pointer");
  if (n_leaves < m->n) return mfail(m, PMK_ERR_ARG, "%lld leaves cannot be dealt to %d ranks (every rank owns at least one leaf)", (long long)n_leaves, m->n);
  if (leaf_off[0] != 0) return mfail(m, PMK_ERR_ARG, "leaf_off[0] must be 0");
  return PMK_OK;
}

void deal_leaves(pmk_multi* m, int64_t n_leaves, const int64_t* leaf_off) {
  std::vector<int64_t> bnd((size_t)m->n + 1, 0);
  balanced_bounds(m->n, n_leaves, leaf_off, bnd.data());
  for (int i = 0; i < m->n; ++i) {
    m->rk[i].leaf_first = bnd[i];
    m->rk[i].leaf_count = bnd[i + 1] - bnd[i];
  }
}

// rank i copies the inputs of its leaves to its GPU (asynchronously on its stream; `sync`: wait for the copies)
int stage_training_rank(pmk_multi* m, int i, int D, const int64_t* leaf_off, const double* X, const double* y, bool sync) {
  Rank& r = m->rk[i];
  const int64_t p0 = leaf_off[r.leaf_first], p1 = leaf_off[r.leaf_first + r.leaf_count];
  r.leaf_off.resize(r.leaf_count + 1);
  for (int64_t k = 0; k <= r.leaf_count; ++k) r.leaf_off[k] = leaf_off[r.leaf_first + k] - p0;
  if (p1 - p0 < 1) {
    r.rc = PMK_ERR_ARG;
    r.err = "no training points";
    return r.rc;
  }
  RC(r, r.dX.ensure(sizeof(double) * (size_t)(p1 - p0) * D));
  RC(r, r.dy.ensure(sizeof(double) * (size_t)(p1 - p0)));
  RC(r, cudaMemcpyAsync(r.dX.p, X + p0 * D, sizeof(double) * (size_t)(p1 - p0) * D, cudaMemcpyHostToDevice, r.stream));
  RC(r, cudaMemcpyAsync(r.dy.p, y + p0, sizeof(double) * (size_t)(p1 - p0), cudaMemcpyHostToDevice, r.stream));
  if (sync) RC(r, cudaStreamSynchronize(r.stream));
  return PMK_OK;
}

}  // namespace

int pmk_multi_stage_training(pmk_multi* m, int D, int64_t n_leaves, const int64_t* leaf_off, const double* X, const double* y) {
  if (!m) return PMK_ERR_ARG;
  m->staged_training = false;
  m->fitted = false;
  if (int rc = check_training(m, D, n_leaves, leaf_off, X, y)) return rc;
  m->D = D;
  m->n_leaves = n_leaves;
  deal_leaves(m, n_leaves, leaf_off);
  const int rc = run_ranks(m, [&](int i) -> int { return stage_training_rank(m, i, D, leaf_off, X, y, true); });
  if (rc == PMK_OK) m->staged_training = true;
  return rc;
}

namespace {

// every rank fits its leaves and builds its query operand; with host inputs (X != nullptr) it stages them first, in the same job
int fit_job(pmk_multi* m, const int64_t* leaf_off, const double* X, const double* y, int kernel_id, const double* kparams, int nparams,
            double sigma2, int64_t* bad_leaf, int* info) {
  if (bad_leaf) *bad_leaf = 0;
  if (info) *info = 0;
  m->fitted = false;
  m->results_ready = false;
  std::vector<int64_t> bad(m->n, 0);
  std::vector<int> inf(m->n, 0);
  const int rc = run_ranks(m, [&](int i) -> int {
    Rank& r = m->rk[i];
    if (X) {
      if (int src = stage_training_rank(m, i, m->D, leaf_off, X, y, false)) return src;      // same stream: the fit's kernels follow the copies
    }
    RK(r, pmk_set_leaf_base(r.h, r.leaf_first, m->n_leaves));
    RC(r, cudaEventRecord(r.ev[EV_FIT0], r.stream));
    const int frc = pmk_fit_dev(r.h, m->D, r.leaf_count, r.leaf_off.data(), r.dX.as<double>(), r.dy.as<double>(), kernel_id, kparams, nparams,
                                sigma2, &bad[i], &inf[i]);
    if (frc != PMK_OK) return rfail(r, frc);
    RK(r, pmk_build_M(r.h));        // the query operand is part of the fit at every rank count
    RC(r, cudaEventRecord(r.ev[EV_FIT1], r.stream));
    RC(r, cudaStreamSynchronize(r.stream));
    r.ms[PMK_MT_FIT] = elapsed(r.ev[EV_FIT0], r.ev[EV_FIT1]);
    return PMK_OK;
  });
  if (rc == PMK_ERR_NOT_POSDEF) {      // the reference stops at the first (lowest) failing leaf (mixtureGP.jl:92,109)
    for (int i = 0; i < m->n; ++i)
      if (bad[i] != 0) {
        if (bad_leaf) *bad_leaf = bad[i];
        if (info) *info = inf[i];
        m->err = "rank " + std::to_string(i) + ": " + m->rk[i].err;
        break;
      }
    return rc;
  }
  if (rc != PMK_OK) return rc;
  m->ms[PMK_MT_FIT] = 0.0;
  for (Rank& r : m->rk) m->ms[PMK_MT_FIT] = std::max(m->ms[PMK_MT_FIT], r.ms[PMK_MT_FIT]);
  m->fitted = true;
  return PMK_OK;
}

}  // namespace

int pmk_multi_fit_staged(pmk_multi* m, int kernel_id, const double* kparams, int nparams, double sigma2, int64_t* bad_leaf, int* info) {
  if (!m) return PMK_ERR_ARG;
  if (!m->staged_training) {
    if (bad_leaf) *bad_leaf = 0;
    if (info) *info = 0;
    return mfail(m, PMK_ERR_STATE, "pmk_multi_stage_training has not been called");
  }
  return fit_job(m, nullptr, nullptr, nullptr, kernel_id, kparams, nparams, sigma2, bad_leaf, info);
}

// stage + fit in ONE job per rank: the host-to-device copies and the fit's kernels are queued on the rank's stream back to back
// (no synchronize, no second wake-up of the rank threads in between)
int pmk_multi_fit(pmk_multi* m, int D, int64_t n_leaves, const int64_t* leaf_off, const double* X, const double* y, int kernel_id,
                  const double* kparams, int nparams, double sigma2, int64_t* bad_leaf, int* info) {
  if (!m) return PMK_ERR_ARG;
  m->staged_training = false;
  m->fitted = false;
  if (bad_leaf) *bad_leaf = 0;
  if (info) *info = 0;
  if (int rc = check_training(m, D, n_leaves, leaf_off, X, y)) return rc;
  m->D = D;
  m->n_leaves = n_leaves;
  deal_leaves(m, n_leaves, leaf_off);
  const int rc = fit_job(m, leaf_off, X, y, kernel_id, kparams, nparams, sigma2, bad_leaf, info);
  // the inputs are on the GPUs whenever every rank got past its copies: a not-positive-definite leaf leaves them staged
  if (rc == PMK_OK || rc == PMK_ERR_NOT_POSDEF) m->staged_training = true;
  return rc;
}

int pmk_multi_set_tree(pmk_multi* m, int D, int levels, const double* hp_v, const double* hp_c) {
  if (!m) return PMK_ERR_ARG;
  DeviceGuard guard;
  m->tree_set = false;
  for (Rank& r : m->rk) {
    const int rc = pmk_set_tree(r.h, D, levels, hp_v, hp_c);
    if (rc != PMK_OK) return mfail(m, rc, "%s", pmk_last_error(r.h));
  }
  m->tree_set = true;
  return PMK_OK;
}

// ---------------------------------------------------------------------------------------------
namespace {

int check_queries(pmk_multi* m, int64_t Nq, const double* Xq) {
  if (m->D == 0) return mfail(m, PMK_ERR_STATE, "query before fit");
  if (Nq < m->n) return mfail(m, PMK_ERR_ARG, "Nq=%lld: every rank plans at least one query (the reference asserts !isempty(Xq))", (long long)Nq);
  if (!Xq) return mfail(m, PMK_ERR_ARG, "NULL pointer");
  return PMK_OK;
}

// rank i copies its slice of the queries to its GPU (asynchronously on its stream; `sync`: wait for the copy)
int stage_queries_rank(pmk_multi* m, int i, int64_t Nq, const double* Xq, bool sync) {
  Rank& r = m->rk[i];
  const int D = m->D;
  range_of(m->n, Nq, i, &r.q_first, &r.q_count);
  RC(r, r.dXq.ensure(sizeof(double) * (size_t)r.q_count * D));
  RC(r, r.dYq.ensure(sizeof(double) * (size_t)r.q_count));
  RC(r, r.dVq.ensure(sizeof(double) * (size_t)r.q_count));
  RC(r, cudaMemcpyAsync(r.dXq.p, Xq + r.q_first * D, sizeof(double) * (size_t)r.q_count * D, cudaMemcpyHostToDevice, r.stream));
  if (sync) RC(r, cudaStreamSynchronize(r.stream));
  return PMK_OK;
}

// rank i copies its slice of the results into the caller's arrays (queued on its stream; the caller of this synchronises)
int fetch_results_rank(pmk_multi* m, int i, double* Yq, double* Vq, bool mean_only) {
  Rank& r = m->rk[i];
  RC(r, cudaMemcpyAsync(Yq + r.q_first, r.dYq.p, sizeof(double) * (size_t)r.q_count, cudaMemcpyDeviceToHost, r.stream));
  if (!mean_only) RC(r, cudaMemcpyAsync(Vq + r.q_first, r.dVq.p, sizeof(double) * (size_t)r.q_count, cudaMemcpyDeviceToHost, r.stream));
  return PMK_OK;
}

}  // namespace

int pmk_multi_stage_queries(pmk_multi* m, int64_t Nq, const double* Xq) {
  if (!m) return PMK_ERR_ARG;
  m->staged_queries = false;
  m->results_ready = false;
  if (int rc = check_queries(m, Nq, Xq)) return rc;
  const int rc = run_ranks(m, [&](int i) -> int { return stage_queries_rank(m, i, Nq, Xq, true); });
  if (rc == PMK_OK) {
    m->staged_queries = true;
    m->Nq = Nq;
  }
  return rc;
}

namespace {

// The query of every rank as ONE job per rank thread.  With host arrays (hXq / hYq / hVq) the job also stages the rank's slice of the
// queries in front of the plan and copies its slice of the results out behind the combine, everything queued on the rank's stream
// back to back: no synchronize and no second wake-up of the rank threads between the copies and the stages.
int query_job(pmk_multi* m, double radius, double delta, int wkernel_id, const double* wparams, int nw, int flags, int64_t Nq,
              const double* hXq, double* hYq, double* hVq) {
  m->results_ready = false;
  const int n = m->n, D = m->D;
  const bool mean_only = (flags & 1) != 0;
  if (n == 1) {
    // one rank owns every leaf: the plan's own leaf binning feeds the pair kernel directly (nothing is packed, copied or re-sorted)
    const int rc1 = run_ranks(m, [&](int) -> int {
      Rank& r = m->rk[0];
      if (hXq) {
        if (int src = stage_queries_rank(m, 0, Nq, hXq, false)) return src;
      }
      RC(r, cudaEventRecord(r.ev[EV_START], r.stream));
      RK(r, pmk_query_plan_dev(r.h, r.q_count, r.dXq.as<double>(), radius, delta, wkernel_id, wparams, nw, &r.n_pairs));
      RC(r, r.pu.ensure(sizeof(double) * (size_t)r.n_pairs));
      RC(r, r.pv.ensure(sizeof(double) * (size_t)r.n_pairs));
      RC(r, cudaEventRecord(r.ev[EV_PLAN], r.stream));
      RC(r, cudaEventRecord(r.ev[EV_ROUTE], r.stream));
      RK(r, pmk_query_pairs_dev(r.h, flags, r.pu.as<double>(), r.pv.as<double>()));
      RC(r, cudaEventRecord(r.ev[EV_PAIRS], r.stream));
      RK(r, pmk_query_combine_dev(r.h, r.pu.as<double>(), r.pv.as<double>(), r.dYq.as<double>(), r.dVq.as<double>()));
      RC(r, cudaEventRecord(r.ev[EV_END], r.stream));
      if (hYq) {
        if (int frc = fetch_results_rank(m, 0, hYq, hVq, mean_only)) return frc;
      }
      RC(r, cudaStreamSynchronize(r.stream));
      r.ms[PMK_MT_QUERY] = elapsed(r.ev[EV_START], r.ev[EV_END]);
      r.ms[PMK_MT_Q_PLAN] = elapsed(r.ev[EV_START], r.ev[EV_PLAN]);
      r.ms[PMK_MT_Q_ROUTE] = 0.0;
      r.ms[PMK_MT_Q_PAIRS] = elapsed(r.ev[EV_ROUTE], r.ev[EV_PAIRS]);
      r.ms[PMK_MT_Q_RETURN] = elapsed(r.ev[EV_PAIRS], r.ev[EV_END]);
      return PMK_OK;
    });
    if (rc1 != PMK_OK) return rc1;
    for (int k : {PMK_MT_QUERY, PMK_MT_Q_PLAN, PMK_MT_Q_ROUTE, PMK_MT_Q_PAIRS, PMK_MT_Q_RETURN}) m->ms[k] = m->rk[0].ms[k];
    m->last_flags = flags;
    m->results_ready = true;
    return PMK_OK;
  }
  const bool alias = false;
  std::vector<int64_t> first_leaf(n + 1);
  for (int i = 0; i < n; ++i) first_leaf[i] = m->rk[i].leaf_first;
  first_leaf[n] = m->n_leaves;
  Meet& meet = *m->meet;
  const int rc = run_ranks(m, [&](int i) -> int {
    Rank& r = m->rk[i];
    // every stage: do the work unless something failed, then meet the others (a failed rank keeps meeting)
    auto plan = [&]() -> int {
      if (hXq) {
        if (int src = stage_queries_rank(m, i, Nq, hXq, false)) return src;
      }
      RC(r, cudaEventRecord(r.ev[EV_START], r.stream));
      RK(r, pmk_query_plan_dev(r.h, r.q_count, r.dXq.as<double>(), radius, delta, wkernel_id, wparams, nw, &r.n_pairs));
      r.seg.assign(n + 1, 0);
      RK(r, pmk_query_plan_segments(r.h, n, first_leaf.data(), r.seg.data()));
      RC(r, r.tx_X.ensure(sizeof(double) * (size_t)r.n_pairs * D));
      RC(r, r.tx_leaf.ensure(sizeof(int32_t) * (size_t)r.n_pairs));
      RC(r, r.pu.ensure(sizeof(double) * (size_t)r.n_pairs));
      RC(r, r.pv.ensure(sizeof(double) * (size_t)r.n_pairs));
      if (!alias) {
        RC(r, r.tu.ensure(sizeof(double) * (size_t)r.n_pairs));
        RC(r, r.tv.ensure(sizeof(double) * (size_t)r.n_pairs));
      }
      RK(r, pmk_query_plan_pack_dev(r.h, r.tx_X.as<double>(), r.tx_leaf.as<int32_t>()));
      RC(r, cudaEventRecord(r.ev[EV_PLAN], r.stream));
      RC(r, cudaStreamSynchronize(r.stream));       // the owners pull from these buffers after the meeting
      return PMK_OK;
    };
    const bool ok1 = meet.sync(plan() == PMK_OK);
    if (!ok1) return r.rc;
    auto own = [&]() -> int {
      // what every planner sends me: its segment for owner i, in planner order
      r.rx_off.assign(n + 1, 0);
      for (int s = 0; s < n; ++s) r.rx_off[s + 1] = r.rx_off[s] + (m->rk[s].seg[i + 1] - m->rk[s].seg[i]);
      r.n_rx = r.rx_off[n];
      const double* rxX = nullptr;
      const int32_t* rxL = nullptr;
      if (alias) {
        rxX = r.tx_X.as<double>();
        rxL = r.tx_leaf.as<int32_t>();
      } else {
        RC(r, r.rx_X.ensure(sizeof(double) * (size_t)std::max<int64_t>(r.n_rx, 1) * D));
        RC(r, r.rx_leaf.ensure(sizeof(int32_t) * (size_t)std::max<int64_t>(r.n_rx, 1)));
        for (int s = 0; s < n; ++s) {
          const Rank& src = m->rk[s];
          const int64_t cnt = src.seg[i + 1] - src.seg[i];
          if (cnt == 0) continue;
          RC(r, cudaMemcpyPeerAsync(r.rx_X.as<double>() + r.rx_off[s] * D, r.device, src.tx_X.as<double>() + src.seg[i] * D, src.device,
                                    sizeof(double) * (size_t)cnt * D, r.stream));
          RC(r, cudaMemcpyPeerAsync(r.rx_leaf.as<int32_t>() + r.rx_off[s], r.device, src.tx_leaf.as<int32_t>() + src.seg[i], src.device,
                                    sizeof(int32_t) * (size_t)cnt, r.stream));
        }
        rxX = r.rx_X.as<double>();
        rxL = r.rx_leaf.as<int32_t>();
      }
      RC(r, cudaEventRecord(r.ev[EV_ROUTE], r.stream));
      RC(r, r.rx_u.ensure(sizeof(double) * (size_t)std::max<int64_t>(r.n_rx, 1)));
      RC(r, r.rx_v.ensure(sizeof(double) * (size_t)std::max<int64_t>(r.n_rx, 1)));
      RK(r, pmk_query_pairs_routed_dev(r.h, r.n_rx, rxX, rxL, flags, r.rx_u.as<double>(), r.rx_v.as<double>()));
      RC(r, cudaEventRecord(r.ev[EV_PAIRS], r.stream));
      RC(r, cudaStreamSynchronize(r.stream));       // the planners pull u, v after the meeting
      return PMK_OK;
    };
    const bool ok2 = meet.sync(own() == PMK_OK);
    if (!ok2) return r.rc;
    auto finish = [&]() -> int {
      const double* us = r.rx_u.as<double>();
      const double* vs = r.rx_v.as<double>();
      if (!alias) {
        for (int o = 0; o < n; ++o) {
          const Rank& own_r = m->rk[o];
          const int64_t cnt = r.seg[o + 1] - r.seg[o];
          if (cnt == 0) continue;
          RC(r, cudaMemcpyPeerAsync(r.tu.as<double>() + r.seg[o], r.device, own_r.rx_u.as<double>() + own_r.rx_off[i], own_r.device,
                                    sizeof(double) * (size_t)cnt, r.stream));
          if (!(flags & 1))
            RC(r, cudaMemcpyPeerAsync(r.tv.as<double>() + r.seg[o], r.device, own_r.rx_v.as<double>() + own_r.rx_off[i], own_r.device,
                                      sizeof(double) * (size_t)cnt, r.stream));
        }
        us = r.tu.as<double>();
        vs = r.tv.as<double>();
      }
      RK(r, pmk_query_plan_unpack_dev(r.h, us, (flags & 1) ? nullptr : vs, r.pu.as<double>(), r.pv.as<double>()));
      RK(r, pmk_query_set_flags(r.h, flags));
      RK(r, pmk_query_combine_dev(r.h, r.pu.as<double>(), r.pv.as<double>(), r.dYq.as<double>(), r.dVq.as<double>()));
      RC(r, cudaEventRecord(r.ev[EV_END], r.stream));
      if (hYq) {
        if (int frc = fetch_results_rank(m, i, hYq, hVq, mean_only)) return frc;
      }
      RC(r, cudaStreamSynchronize(r.stream));
      r.ms[PMK_MT_QUERY] = elapsed(r.ev[EV_START], r.ev[EV_END]);
      r.ms[PMK_MT_Q_PLAN] = elapsed(r.ev[EV_START], r.ev[EV_PLAN]);
      r.ms[PMK_MT_Q_ROUTE] = elapsed(r.ev[EV_PLAN], r.ev[EV_ROUTE]);
      r.ms[PMK_MT_Q_PAIRS] = elapsed(r.ev[EV_ROUTE], r.ev[EV_PAIRS]);
      r.ms[PMK_MT_Q_RETURN] = elapsed(r.ev[EV_PAIRS], r.ev[EV_END]);
      return PMK_OK;
    };
    // the owners' answer buffers must stay untouched until every planner has pulled its share
    const bool ok3 = meet.sync(finish() == PMK_OK);
    return ok3 ? PMK_OK : r.rc;
  });
  if (rc != PMK_OK) return rc;
  for (int k : {PMK_MT_QUERY, PMK_MT_Q_PLAN, PMK_MT_Q_ROUTE, PMK_MT_Q_PAIRS, PMK_MT_Q_RETURN}) {
    m->ms[k] = 0.0;
    for (Rank& r : m->rk) m->ms[k] = std::max(m->ms[k], r.ms[k]);
  }
  m->last_flags = flags;
  m->results_ready = true;
  return PMK_OK;
}

}  // namespace

int pmk_multi_query_staged(pmk_multi* m, double radius, double delta, int wkernel_id, const double* wparams, int nw, int flags) {
  if (!m) return PMK_ERR_ARG;
  if (!m->fitted) return mfail(m, PMK_ERR_STATE, "query before fit");
  if (!m->tree_set) return mfail(m, PMK_ERR_STATE, "query before pmk_multi_set_tree");
  if (!m->staged_queries) return mfail(m, PMK_ERR_STATE, "pmk_multi_stage_queries has not been called");
  return query_job(m, radius, delta, wkernel_id, wparams, nw, flags, m->Nq, nullptr, nullptr, nullptr);
}

int pmk_multi_fetch_results(pmk_multi* m, double* Yq, double* Vq) {
  if (!m) return PMK_ERR_ARG;
  if (!m->results_ready) return mfail(m, PMK_ERR_STATE, "no query results (pmk_multi_query_staged)");
  const bool mean_only = (m->last_flags & 1) != 0;
  if (!Yq || (!mean_only && !Vq)) return mfail(m, PMK_ERR_ARG, "NULL pointer");
  // every rank copies its slice straight into the caller's arrays: n concurrent device-to-host streams, no gather on a device
  return run_ranks(m, [&](int i) -> int {
    Rank& r = m->rk[i];
    if (int rc = fetch_results_rank(m, i, Yq, Vq, mean_only)) return rc;
    RC(r, cudaStreamSynchronize(r.stream));
    return PMK_OK;
  });
}

// stage + query + fetch in ONE job per rank (see query_job)
int pmk_multi_query(pmk_multi* m, int64_t Nq, const double* Xq, double radius, double delta, int wkernel_id, const double* wparams, int nw,
                    int flags, double* Yq, double* Vq) {
  if (!m) return PMK_ERR_ARG;
  if (!Yq || (!(flags & 1) && !Vq)) return mfail(m, PMK_ERR_ARG, "NULL pointer");
  m->staged_queries = false;
  m->results_ready = false;
  if (int rc = check_queries(m, Nq, Xq)) return rc;
  if (!m->fitted) return mfail(m, PMK_ERR_STATE, "query before fit");
  if (!m->tree_set) return mfail(m, PMK_ERR_STATE, "query before pmk_multi_set_tree");
  const int rc = query_job(m, radius, delta, wkernel_id, wparams, nw, flags, Nq, Xq, Yq, Vq);
  if (rc == PMK_OK) {
    m->staged_queries = true;
    m->Nq = Nq;
  }
  return rc;
}

int pmk_multi_leaf_pairs(pmk_multi* m, int64_t* pairs_per_leaf) {
  if (!m || !pairs_per_leaf) return PMK_ERR_ARG;
  if (!m->results_ready) return mfail(m, PMK_ERR_STATE, "no query results (pmk_multi_query_staged)");
  DeviceGuard guard;
  std::vector<int64_t> one((size_t)m->n_leaves);
  std::fill(pairs_per_leaf, pairs_per_leaf + m->n_leaves, (int64_t)0);
  for (Rank& r : m->rk) {
    const int rc = pmk_last_query_leaf_pairs(r.h, one.data());
    if (rc != PMK_OK) return mfail(m, rc, "%s", pmk_last_error(r.h));
    for (int64_t g = 0; g < m->n_leaves; ++g) pairs_per_leaf[g] += one[g];
  }
  return PMK_OK;
}

int pmk_multi_get_timings(pmk_multi* m, double* ms, double* per_rank_ms) {
  if (!m || !ms) return PMK_ERR_ARG;
  DeviceGuard guard;
  for (int k = 0; k < PMK_MT_COUNT; ++k) ms[k] = m->ms[k];
  if (per_rank_ms)
    for (int i = 0; i < m->n; ++i) {
      const int rc = pmk_get_timings(m->rk[i].h, per_rank_ms + (size_t)i * PMK_T_COUNT);
      if (rc != PMK_OK) return mfail(m, rc, "%s", pmk_last_error(m->rk[i].h));
    }
  return PMK_OK;
}

int64_t pmk_multi_launch_count(const pmk_multi* m) {
  int64_t t = 0;
  if (m)
    for (const Rank& r : m->rk) t += pmk_launch_count(r.h);
  return t;
}

}  // extern "C"
