/* CPU ORACLE, C restatement (test infrastructure + timed CPU baseline, NOT product code)
 * -- PARITY UNPINNED (the reference is Julia, cannot run here, and ships no golden vectors;
 *    see the header of oracle/pmk_oracle.py for what pins these restatements instead).
 *
 * Plain scalar loops with the loop structure of the reference's Julia code -- this is the
 * "reference CPU path" that bench.py times on the GPU box's host cores (cpu_baseline.kind =
 * "port").  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product never does.
 *
 * One difference from the reference is deliberate and generous to the baseline: the reference is
 * single-threaded outside BLAS; here leaves (fit) and queries (query) are spread over OpenMP
 * threads.  Numerically each leaf / query is processed exactly as one Julia iteration would.
 *
 * build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC oracle/pmk_oracle.c -o oracle/_build/libpmk_oracle.so -lm
 *        (-ffp-contract=off: Julia never fuses a*b+c; comparisons in the tree code depend on it)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { SQEXP = 0, SPLINE34, BB10, BB20, BB1EPS, BB2EPS, SPLINE12, SPLINE32, RQ };

static int is_stationary(int k) { return k == SQEXP || k == SPLINE34 || k == SPLINE12 || k == SPLINE32 || k == RQ; }

/* evalkernel(tau, theta): src/RKHS/kernel.jl:299-374 */
static double k_tau(int kind, double a, double tau) {
  switch (kind) {
    case SQEXP: return exp(-a * (tau * tau));                                  /* kernel.jl:350-357 */
    case SPLINE34: {                                                            /* kernel.jl:299-313 */
      double r = tau * a, t = 1.0 - r;
      if (t < 0.0) return 0.0;
      return ((35.0 * (r * r) + 18.0 * r + 3.0) * pow(t, 6.0)) / 3.0;
    }
    case SPLINE12: {                                                            /* kernel.jl:316-330 */
      double r = tau * a, t = 1.0 - r;
      if (t < 0.0) return 0.0;
      return (3.0 * r + 1.0) * (t * t * t);
    }
    case SPLINE32: {                                                            /* kernel.jl:333-347 */
      double r = tau * a, t = 1.0 - r;
      if (t < 0.0) return 0.0;
      return (4.0 * r + 1.0) * pow(t, 4.0);
    }
    case RQ: {                                                                  /* kernel.jl:360-366 */
      double sa = sqrt(a), sd = sqrt(a + tau * tau);
      return (sa * sa * sa) / (sd * sd * sd);
    }
  }
  return 0.0;
}

/* Brownian-bridge kernels on [0,1]: kernel.jl:156-225 */
static double k_bb(int kind, double e, double x, double z) {
  switch (kind) {
    case BB10: return fmin(x, z) - x * z;
    case BB20:
      if (z < x) return (((-1.0 / 6.0) * z) * (1.0 - x)) * ((x * x + z * z) - 2.0 * x);
      return (((-1.0 / 6.0) * x) * (1.0 - z)) * ((x * x + z * z) - 2.0 * z);
    case BB1EPS: return (sinh(e * fmin(x, z)) * sinh(e * (1.0 - fmax(x, z)))) / (e * sinh(e));
    case BB2EPS: {
      double mn = fmin(x, z), mx = fmax(x, z), ad = fabs(x - z), s = x + z;
      double em1 = exp(2.0 * e) - 1.0;
      double mult = exp(-e * s) / (4.0 * (e * e * e) * (em1 * em1));
      double t1 = exp(2.0 * e) * (2.0 * e - e * s - 1.0);
      double t2 = exp(4.0 * e) * (e * s + 1.0);
      double t3 = exp(2.0 * e * (1.0 + s)) * (2.0 * e - e * s + 1.0);
      double t4 = exp(2.0 * e * s) * (e * s - 1.0);
      double t5 = exp(2.0 * e * (2.0 + mn)) * (-e * ad - 1.0);
      double t6 = exp(2.0 * e * mx) * (-e * ad + 1.0);
      double t7 = exp(2.0 * e * (1.0 + mn)) * (1.0 - 2.0 * e + e * ad);
      double t8 = exp(2.0 * e * (1.0 + mx)) * (1.0 + 2.0 * e - e * ad);
      return mult * (((((((t1 + t2) + t3) + t4) + t5) + t6) + t7) + t8);
    }
  }
  return 0.0;
}

/* evalkernel(x1, x2, theta): kernel.jl:277-287 (tau = norm(x1-x2)), :196-198 (tensor product) */
static double evalkernel(int D, const double* x, const double* z, int kind, double a) {
  if (is_stationary(kind)) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) {
      double dd = x[d] - z[d];
      s = (d == 0) ? dd * dd : s + dd * dd;
    }
    return k_tau(kind, a, sqrt(s));
  }
  double out = k_bb(kind, a, x[0], z[0]);
  for (int d = 1; d < D; ++d) out = out * k_bb(kind, a, x[d], z[d]);
  return out;
}

static double dot_seq(int D, const double* v, const double* x) {
  double s = v[0] * x[0];
  for (int d = 1; d < D; ++d) s = s + v[d] * x[d];
  return s;
}

/* constructkernelmatrix!: RKHS.jl:13-34 (lower triangle, then mirror); K is n x n column-major */
static void gram(int D, int n, const double* X, int kind, double a, double* K) {
  for (int j = 0; j < n; ++j)
    for (int i = j; i < n; ++i) K[i + (size_t)n * j] = evalkernel(D, X + (size_t)i * D, X + (size_t)j * D, kind, a);
  for (int j = 1; j < n; ++j)
    for (int i = 0; i < j; ++i) K[i + (size_t)n * j] = K[j + (size_t)n * i];
}

int pmk_oracle_gram(int D, int64_t n, const double* X, int kind, double a, double sigma2, double* K) {
  gram(D, (int)n, X, kind, a, K);
  for (int64_t i = 0; i < n; ++i) K[i + n * i] += sigma2;
  return 0;
}

/* U\y for a dense square matrix = LU with partial pivoting (dgetrf + dgetrs); A is overwritten */
static int lu_solve(int n, double* A, double* b, int* piv) {
  for (int k = 0; k < n; ++k) {
    int p = k;
    double mx = fabs(A[k + (size_t)n * k]);
    for (int i = k + 1; i < n; ++i) {
      double v = fabs(A[i + (size_t)n * k]);
      if (v > mx) { mx = v; p = i; }
    }
    piv[k] = p;
    if (mx == 0.0) return k + 1;
    if (p != k)
      for (int j = 0; j < n; ++j) {
        double t = A[k + (size_t)n * j];
        A[k + (size_t)n * j] = A[p + (size_t)n * j];
        A[p + (size_t)n * j] = t;
      }
    double inv = 1.0 / A[k + (size_t)n * k];
    for (int i = k + 1; i < n; ++i) A[i + (size_t)n * k] *= inv;
    for (int j = k + 1; j < n; ++j) {
      double akj = A[k + (size_t)n * j];
      double* cj = A + (size_t)n * j;
      const double* ck = A + (size_t)n * k;
      for (int i = k + 1; i < n; ++i) cj[i] -= ck[i] * akj;
    }
  }
  for (int k = 0; k < n; ++k) {
    if (piv[k] != k) { double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }
  }
  for (int k = 0; k < n; ++k) {                 /* L y = Pb (unit lower) */
    double bk = b[k];
    const double* ck = A + (size_t)n * k;
    for (int i = k + 1; i < n; ++i) b[i] -= ck[i] * bk;
  }
  for (int k = n - 1; k >= 0; --k) {            /* U x = y */
    b[k] /= A[k + (size_t)n * k];
    double bk = b[k];
    const double* ck = A + (size_t)n * k;
    for (int i = 0; i < k; ++i) b[i] -= ck[i] * bk;
  }
  return 0;
}

/* cholesky(U).L: lower factor, column-major, upper triangle zeroed; info as LAPACK dpotrf */
static int chol_lower(int n, double* A) {
  for (int j = 0; j < n; ++j) {
    double d = A[j + (size_t)n * j];
    if (!(d > 0.0)) return j + 1;
    d = sqrt(d);
    A[j + (size_t)n * j] = d;
    double inv = 1.0 / d;
    double* cj = A + (size_t)n * j;
    for (int i = j + 1; i < n; ++i) cj[i] *= inv;
    for (int k = j + 1; k < n; ++k) {
      double lkj = cj[k];
      double* ck = A + (size_t)n * k;
      for (int i = k; i < n; ++i) ck[i] -= cj[i] * lkj;
    }
  }
  for (int j = 1; j < n; ++j)
    for (int i = 0; i < j; ++i) A[i + (size_t)n * j] = 0.0;
  return 0;
}

/* fitmixtureGP!: src/RKHS/mixtureGP.jl:92-115, per leaf.  L_out: dense n_p x n_p column-major blocks back
 * to back (offset of leaf p = sum_{q<p} n_q^2).  Returns 0, or -3 with bad_leaf (1-based) / info set. */
int pmk_oracle_fit(int D, int64_t n_leaves, const int64_t* leaf_off, const double* X, const double* y, int kind, double a,
                   double sigma2, double* alpha, double* L_out, int64_t* bad_leaf, int* info, int nthreads) {
  int64_t* loff = (int64_t*)malloc(sizeof(int64_t) * (n_leaves + 1));
  loff[0] = 0;
  for (int64_t p = 0; p < n_leaves; ++p) {
    int64_t n = leaf_off[p + 1] - leaf_off[p];
    loff[p + 1] = loff[p] + n * n;
  }
  int64_t first_bad = 0;
  int first_info = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t p = 0; p < n_leaves; ++p) {
    int n = (int)(leaf_off[p + 1] - leaf_off[p]);
    const double* Xp = X + leaf_off[p] * D;
    double* U = (double*)malloc(sizeof(double) * (size_t)n * n);
    double* W = (double*)malloc(sizeof(double) * (size_t)n * n);
    int* piv = (int*)malloc(sizeof(int) * n);
    gram(D, n, Xp, kind, a, U);                                /* mixtureGP.jl:98  */
    for (int i = 0; i < n; ++i) U[i + (size_t)n * i] += sigma2; /* :102-104         */
    memcpy(W, U, sizeof(double) * (size_t)n * n);
    double* c = alpha + leaf_off[p];
    memcpy(c, y + leaf_off[p], sizeof(double) * n);
    lu_solve(n, W, c, piv);                                     /* :106  c = U\y    */
    int inf = chol_lower(n, U);                                 /* :109  cholesky(U) */
    if (inf != 0) {
#pragma omp critical
      if (first_bad == 0 || p + 1 < first_bad) { first_bad = p + 1; first_info = inf; }
    } else if (L_out) {
      memcpy(L_out + loff[p], U, sizeof(double) * (size_t)n * n);
    }
    free(U); free(W); free(piv);
  }
  free(loff);
  if (bad_leaf) *bad_leaf = first_bad;
  if (info) *info = first_info;
  return first_bad ? -3 : 0;
}

/* findpartition: src/patchwork/partition.jl:248-262 on the flattened (pre-order) tree; hv is n_hp x D row-major */
static int findpartition(int D, int levels, const double* hv, const double* hc, const double* x) {
  int Lv = levels - 1, node = 0, leaf = 0;
  for (int d = 0; d < Lv; ++d) {
    int right = !(dot_seq(D, hv + (size_t)node * D, x) < hc[node]);
    leaf = leaf * 2 + right;
    node += right ? (1 << (Lv - 1 - d)) : 1;
  }
  return leaf + 1;
}

int pmk_oracle_findpartition(int D, int levels, const double* hv, const double* hc, int64_t Nq, const double* Xq, int32_t* out) {
  for (int64_t j = 0; j < Nq; ++j) out[j] = findpartition(D, levels, hv, hc, Xq + j * D);
  return 0;
}

/* queryinner!: mixtureGP.jl:296-316.  L dense column-major; forward substitution = dtrsv('L','N','N') */
static void queryinner(int D, int n, const double* Xp, const double* c, const double* L, const double* xq, int kind, double a,
                       double* kq, double* mu, double* vq) {
  for (int i = 0; i < n; ++i) kq[i] = evalkernel(D, xq, Xp + (size_t)i * D, kind, a);
  double m = 0.0;
  for (int i = 0; i < n; ++i) m += kq[i] * c[i];
  for (int j = 0; j < n; ++j) {
    double xj = kq[j] / L[j + (size_t)n * j];
    kq[j] = xj;
    const double* cj = L + (size_t)n * j;
    for (int i = j + 1; i < n; ++i) kq[i] -= xj * cj[i];
  }
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += kq[i] * kq[i];
  double v = evalkernel(D, xq, xq, kind, a) - s;
  if (v < 1e-12) v = 1e-12;
  *mu = m;
  *vq = v;
}

/* querymixtureGP!: mixtureGP.jl:159-294 with findneighbourpartitions :339-405 (scan over ALL hyperplanes).
 * L: dense blocks as written by pmk_oracle_fit.  Leaves of zero length in leaf_off are "absent" (a bounded
 * CPU sample fits only some leaves of the real tree): a query touching one gets NaN, and with
 * structure_only != 0 nothing is evaluated -- only home_out / npairs_out / absent_out are filled, which is
 * how bench.py picks sample queries whose leaves are all present. */
int pmk_oracle_query(int D, int levels, const double* hv, const double* hc, int64_t n_leaves, const int64_t* leaf_off,
                     const double* X, const double* alpha, const double* L, int kind, double a, int64_t Nq, const double* Xq,
                     double radius, double delta, int wkind, double wa, double* Yq, double* Vq, int32_t* home_out,
                     int32_t* npairs_out, int32_t* absent_out, int structure_only, int nthreads) {
  int n_hp = (1 << (levels - 1)) - 1;
  int64_t* loff = (int64_t*)malloc(sizeof(int64_t) * (n_leaves + 1));
  int maxn = 1;
  loff[0] = 0;
  for (int64_t p = 0; p < n_leaves; ++p) {
    int64_t n = leaf_off[p + 1] - leaf_off[p];
    loff[p + 1] = loff[p] + n * n;
    if (n > maxn) maxn = (int)n;
  }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double* kq = (double*)malloc(sizeof(double) * maxn);
    int cap = 16;
    int* region = (int*)malloc(sizeof(int) * cap);
    double* w = (double*)malloc(sizeof(double) * (cap + 1));
    double* u = (double*)malloc(sizeof(double) * (cap + 1));
    double* v = (double*)malloc(sizeof(double) * (cap + 1));
    double z1[8], z2[8];
#pragma omp for schedule(dynamic, 64)
    for (int64_t j = 0; j < Nq; ++j) {
      const double* p = Xq + j * D;
      int home = findpartition(D, levels, hv, hc, p);                       /* :206 */
      int nr = 0;
      for (int i = 0; i < n_hp; ++i) {                                      /* :354 */
        const double* uu = hv + (size_t)i * D;
        double t = -dot_seq(D, uu, p) + hc[i];                              /* :361 */
        double s = 0.0;
        for (int d = 0; d < D; ++d) {
          double z = p[d] + t * uu[d];                                      /* :362 */
          double dd = z - p[d];
          s = (d == 0) ? dd * dd : s + dd * dd;
        }
        if (sqrt(s) < radius) {                                             /* :367 */
          double tp = t + delta, tm = t - delta;
          for (int d = 0; d < D; ++d) { z1[d] = p[d] + tp * uu[d]; z2[d] = p[d] + tm * uu[d]; }
          int r1 = findpartition(D, levels, hv, hc, z1), r2 = findpartition(D, levels, hv, hc, z2);
          if ((r2 == home) != (r1 == home)) {                               /* :387 */
            if (nr == cap) {
              cap *= 2;
              region = (int*)realloc(region, sizeof(int) * cap);
              w = (double*)realloc(w, sizeof(double) * (cap + 1));
              u = (double*)realloc(u, sizeof(double) * (cap + 1));
              v = (double*)realloc(v, sizeof(double) * (cap + 1));
            }
            region[nr] = (r1 == home) ? r2 : r1;
            w[nr] = k_tau(wkind, wa, fabs(t));                              /* :231 */
            ++nr;
          }
        }
      }
      int absent = 0;
      for (int i = 0; i <= nr; ++i) {
        int r = (i < nr) ? region[i] : home;
        if (leaf_off[r] == leaf_off[r - 1]) absent = 1;
      }
      if (home_out) home_out[j] = home;
      if (npairs_out) npairs_out[j] = nr + 1;
      if (absent_out) absent_out[j] = absent;
      if (structure_only) continue;
      if (absent) { Yq[j] = NAN; Vq[j] = NAN; continue; }
      for (int i = 0; i <= nr; ++i) {
        int r = (i < nr) ? region[i] : home;
        if (i == nr) w[i] = 1.0;                                            /* :237 */
        int n = (int)(leaf_off[r] - leaf_off[r - 1]);
        queryinner(D, n, X + leaf_off[r - 1] * D, alpha + leaf_off[r - 1], L + loff[r - 1], p, kind, a, kq, &u[i], &v[i]);
      }
      double sw = 0.0;
      for (int i = 0; i <= nr; ++i) sw = sw + w[i];                         /* :263 */
      double yy = 0.0, vv = 0.0;
      for (int i = 0; i <= nr; ++i) {
        double wi = w[i] / sw;
        yy = yy + wi * u[i];                                                /* :269 */
        vv = vv + wi * (v[i] * wi);                                         /* :272 */
      }
      Yq[j] = yy;
      Vq[j] = vv;
    }
    free(kq); free(region); free(w); free(u); free(v);
  }
  free(loff);
  return 0;
}

int pmk_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
