/* gv_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's all-pairs co-expression path
 * (nceglia/genevector; citations are into /root/reference):
 *   gvo_discretize_*   genevector/metrics.py:25-69 (discretize_genes), with the arithmetic of
 *                      numpy.percentile(method='linear') + np.linspace + np.searchsorted restated
 *                      from numpy 2.3.5 lib/_function_base_impl.py (_quantile 4788-4879,
 *                      _get_indexes 4753-4786, _get_gamma 4633-4654, _lerp 4657-4676; the
 *                      reference pins numpy==1.26.1, same algorithm).
 *   gvo_mi_pairs       src/lib.rs:30-98 (compute_mi_pairs): pair list, u32 joint histogram that
 *                      skips cells where both bins are 0 (lib.rs:52), total = count (lib.rs:62),
 *                      f64 marginals (65-76) and MI (79-88); OpenMP "parallel for" stands in for
 *                      rayon's par_iter (lib.rs:40-41).  Same body as metrics.py:161-198 (Numba)
 *                      and metrics.py:116-135 + 74-81 (NumPy tier).
 *   gvo_joint          the joint histogram of one pair, exposed raw for bit-exact count checks.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this.  Parity pinned by tests/golden/ (generated from the reference itself by
 * oracle/gen_golden.py); the reference's own tests hold no golden vectors for this path.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int gvo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline sets its thread count itself */
void gvo_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}


/* ------------------------------------------------------------------ discretisation */

static int cmp_f32(const void* a, const void* b) {
  const float x = *(const float*)a, y = *(const float*)b;
  return (x > y) - (x < y);
}
static int cmp_f64(const void* a, const void* b) {
  const double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

/* np.linspace(0, 100, k+1)[t] / 100 : step = 100/k; y = t*step + 0; then percentile divides by 100 */
double gvo_quantile_fraction(int k, int t) {
  volatile double step = 100.0 / (double)k;
  volatile double y = (double)t * step;
  y = y + 0.0;
  return y / 100.0;
}

/* One gene column given its sorted positive values s[0..m) (as doubles converted exactly from
 * the input dtype; `is_f32` says the subtraction b-a has to be rounded to float32).
 * Writes k-1 edges, returns k = min(n_bins, n_unique).  metrics.py:58-65 */
static int gene_edges(const double* s, int64_t m, int is_f32, int n_bins, double* edges) {
  int64_t n_uniq = 1;
  for (int64_t t = 1; t < m && n_uniq < n_bins; ++t) n_uniq += (s[t] != s[t - 1]);
  const int k = (int)(n_uniq < n_bins ? n_uniq : n_bins);
  if (k <= 1) return k;
  for (int t = 1; t < k; ++t) {
    const double q = gvo_quantile_fraction(k, t);
    volatile double vi = (double)(m - 1) * q;         /* 'linear': (n-1)*quantiles */
    double a, b, gamma;
    if (vi >= (double)(m - 1)) {                      /* _get_indexes: above bounds -> last */
      a = b = s[m - 1];
      gamma = vi + 1.0;                               /* previous index is -1 there */
    } else {
      const double p = floor(vi);
      a = s[(int64_t)p];
      b = s[(int64_t)p + 1];
      gamma = vi - p;
    }
    volatile double d;                                /* subtract(b, a) in the array dtype */
    if (is_f32) { volatile float df = (float)b - (float)a; d = (double)df; }
    else d = b - a;
    volatile double e;
    if (gamma >= 0.5) { volatile double w = 1.0 - gamma; volatile double pr = d * w; e = b - pr; }
    else { volatile double pr = d * gamma; e = a + pr; }
    edges[t - 1] = e;
  }
  return k;
}

static int bin_of(double v, const double* edges, int n_edges) {
  int bin = 1;                                        /* searchsorted(edges, v, 'right') + 1 */
  for (int t = 0; t < n_edges; ++t) bin += (edges[t] <= v);
  return bin;
}

/* Dense [n_cells][n_genes] C-order input, dtype 0 = float32, 1 = float64.
 * x_disc int32 [n_cells][n_genes], n_bins_per_gene int32 [n_genes], edges (optional) [n_genes][32]. */
int gvo_discretize_dense(const void* X, int dtype, int64_t n_cells, int64_t n_genes, int n_bins,
                         int32_t* x_disc, int32_t* n_bins_per_gene, double* edges_out) {
  if (n_bins < 1 || n_bins > 31) return -1;
  memset(x_disc, 0, sizeof(int32_t) * (size_t)n_cells * (size_t)n_genes);
  int rc = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t g = 0; g < n_genes; ++g) {
    double* s = (double*)malloc(sizeof(double) * (size_t)(n_cells > 0 ? n_cells : 1));
    if (!s) { rc = -2; continue; }
    int64_t m = 0;
    for (int64_t c = 0; c < n_cells; ++c) {
      const double v = dtype == 0 ? (double)((const float*)X)[c * n_genes + g]
                                  : ((const double*)X)[c * n_genes + g];
      if (v > 0) s[m++] = v;                          /* metrics.py:53-54 */
    }
    double edges[32];
    for (int t = 0; t < 32; ++t) edges[t] = INFINITY;
    if (m == 0) {
      n_bins_per_gene[g] = 1;                         /* metrics.py:55-57 */
    } else {
      qsort(s, (size_t)m, sizeof(double), cmp_f64);
      const int k = gene_edges(s, m, dtype == 0, n_bins, edges);
      n_bins_per_gene[g] = k <= 1 ? 2 : k + 1;        /* metrics.py:62,67 */
      const int ne = k <= 1 ? 0 : k - 1;
      for (int64_t c = 0; c < n_cells; ++c) {
        const double v = dtype == 0 ? (double)((const float*)X)[c * n_genes + g]
                                    : ((const double*)X)[c * n_genes + g];
        if (v > 0) x_disc[c * n_genes + g] = bin_of(v, edges, ne);
      }
    }
    if (edges_out) memcpy(edges_out + g * 32, edges, sizeof(edges));
    free(s);
  }
  (void)cmp_f32;
  return rc;
}

/* CSR input (canonical) -> gene-major bins uint8 [n_genes][n_cells]; avoids the dense matrix the
 * reference builds (metrics.py:45-46) so the oracle can run on larger shapes. */
int gvo_discretize_csr(const void* values, int dtype, const int32_t* col_idx, const int64_t* row_ptr,
                       int64_t n_cells, int64_t n_genes, int n_bins, uint8_t* bins_gm,
                       int32_t* n_bins_per_gene, double* edges_out) {
  if (n_bins < 1 || n_bins > 31) return -1;
  const int64_t nnz = row_ptr[n_cells];
  int64_t* gptr = (int64_t*)calloc((size_t)n_genes + 1, sizeof(int64_t));
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * ((size_t)n_genes + 1));
  double* gv = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
  int32_t* gc = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  if (!gptr || !cur || !gv || !gc) { free(gptr); free(cur); free(gv); free(gc); return -2; }
#define VAL(e) (dtype == 0 ? (double)((const float*)values)[e] : ((const double*)values)[e])
  for (int64_t e = 0; e < nnz; ++e) if (VAL(e) > 0) gptr[col_idx[e] + 1]++;
  for (int64_t g = 0; g < n_genes; ++g) gptr[g + 1] += gptr[g];
  memcpy(cur, gptr, sizeof(int64_t) * ((size_t)n_genes + 1));
  for (int64_t r = 0; r < n_cells; ++r)
    for (int64_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
      const double v = VAL(e);
      if (v > 0) { const int64_t p = cur[col_idx[e]]++; gv[p] = v; gc[p] = (int32_t)r; }
    }
#undef VAL
  memset(bins_gm, 0, (size_t)n_genes * (size_t)n_cells);
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t g = 0; g < n_genes; ++g) {
    const int64_t m = gptr[g + 1] - gptr[g];
    double edges[32];
    for (int t = 0; t < 32; ++t) edges[t] = INFINITY;
    if (m == 0) {
      n_bins_per_gene[g] = 1;
    } else {
      double* s = (double*)malloc(sizeof(double) * (size_t)m);
      memcpy(s, gv + gptr[g], sizeof(double) * (size_t)m);
      qsort(s, (size_t)m, sizeof(double), cmp_f64);
      const int k = gene_edges(s, m, dtype == 0, n_bins, edges);
      free(s);
      n_bins_per_gene[g] = k <= 1 ? 2 : k + 1;
      const int ne = k <= 1 ? 0 : k - 1;
      for (int64_t p = gptr[g]; p < gptr[g + 1]; ++p)
        bins_gm[g * n_cells + gc[p]] = (uint8_t)bin_of(gv[p], edges, ne);
    }
    if (edges_out) memcpy(edges_out + g * 32, edges, sizeof(edges));
  }
  free(gptr); free(cur); free(gv); free(gc);
  return 0;
}

/* ------------------------------------------------------------------ joint histogram + MI */

/* lib.rs:46-56 for one pair.  joint is u32 [na*nb]; returns count. */
uint32_t gvo_joint(const int32_t* x_disc, int64_t n_cells, int64_t n_genes, int64_t i, int64_t j,
                   int na, int nb, uint32_t* joint) {
  memset(joint, 0, sizeof(uint32_t) * (size_t)na * (size_t)nb);
  uint32_t count = 0;
  for (int64_t c = 0; c < n_cells; ++c) {
    const int a = x_disc[c * n_genes + i];
    const int b = x_disc[c * n_genes + j];
    if (a > 0 || b > 0) { joint[a * nb + b] += 1; count += 1; }
  }
  return count;
}

/* lib.rs:62-88 */
double gvo_mi_from_joint_u32(const uint32_t* joint, int na, int nb, uint32_t count) {
  const double total = (double)count;
  double px[64], py[64], jf[64 * 64];
  for (int a = 0; a < na; ++a) px[a] = 0.0;
  for (int b = 0; b < nb; ++b) py[b] = 0.0;
  for (int a = 0; a < na; ++a)
    for (int b = 0; b < nb; ++b) {
      const double p = (double)joint[a * nb + b] / total;
      jf[a * nb + b] = p; px[a] += p; py[b] += p;
    }
  double mi = 0.0;
  for (int a = 0; a < na; ++a)
    for (int b = 0; b < nb; ++b) {
      const double pxy = jf[a * nb + b];
      const double pp = px[a] * py[b];
      if (pxy > 0.0 && pp > 0.0) mi += pxy * log2(pxy / pp);
    }
  return mi;
}

/* All upper-triangle pairs (lib.rs:30-98).  x_disc int32 [n_cells][n_genes] C-order exactly as
 * the reference holds it (strided column walks included: this is also the CPU baseline).
 *   mi_out  : float64 [n_genes][n_genes], both triangles written, skipped pairs stay 0
 *   corr    : optional float64 [n_genes][n_genes]; sign_mode 1 = np.sign(corr) (metrics.py:134,
 *             the documented rule), 2 = lib.rs:91-94 (corr cast to f32, >= 0 -> +1 else -1)
 * Returns the number of pairs computed. */
int64_t gvo_mi_pairs(const int32_t* x_disc, const int32_t* n_bins_per_gene, int64_t n_cells,
                     int64_t n_genes, const double* corr, int sign_mode, double* mi_out) {
  int64_t done = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : done)
  for (int64_t i = 0; i < n_genes; ++i) {
    uint32_t joint[64 * 64];
    for (int64_t j = i + 1; j < n_genes; ++j) {
      const int na = n_bins_per_gene[i], nb = n_bins_per_gene[j];
      if (na <= 1 || nb <= 1 || na > 64 || nb > 64) continue;            /* lib.rs:33 */
      const uint32_t count = gvo_joint(x_disc, n_cells, n_genes, i, j, na, nb, joint);
      if (count == 0) continue;                                          /* lib.rs:58-60 */
      double mi = gvo_mi_from_joint_u32(joint, na, nb, count);
      if (corr && sign_mode == 1) {
        const double c = corr[i * n_genes + j];
        mi *= (c > 0) - (c < 0);
      } else if (corr && sign_mode == 2) {
        const float c = (float)corr[i * n_genes + j];
        mi *= c >= 0.0f ? 1.0 : -1.0;
      }
      mi_out[i * n_genes + j] = mi;
      mi_out[j * n_genes + i] = mi;
      ++done;
    }
  }
  return done;
}

/* Same, on gene-major uint8 bins [n_genes][n_cells] (unit-stride walks); for larger parity
 * cases where the int32 cell-major matrix of the reference would not fit. */
int64_t gvo_mi_pairs_gm(const uint8_t* bins_gm, const int32_t* n_bins_per_gene, int64_t n_cells,
                        int64_t n_genes, const int8_t* sign, double* mi_out) {
  int64_t done = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : done)
  for (int64_t i = 0; i < n_genes; ++i) {
    uint32_t joint[64 * 64];
    const uint8_t* ci = bins_gm + i * n_cells;
    for (int64_t j = i + 1; j < n_genes; ++j) {
      const int na = n_bins_per_gene[i], nb = n_bins_per_gene[j];
      if (na <= 1 || nb <= 1 || na > 64 || nb > 64) continue;
      const uint8_t* cj = bins_gm + j * n_cells;
      memset(joint, 0, sizeof(uint32_t) * (size_t)na * (size_t)nb);
      uint32_t count = 0;
      for (int64_t c = 0; c < n_cells; ++c) {
        const int a = ci[c], b = cj[c];
        if (a > 0 || b > 0) { joint[a * nb + b] += 1; count += 1; }
      }
      if (count == 0) continue;
      double mi = gvo_mi_from_joint_u32(joint, na, nb, count);
      if (sign) mi *= (double)sign[i * n_genes + j];
      mi_out[i * n_genes + j] = mi;
      mi_out[j * n_genes + i] = mi;
      ++done;
    }
  }
  return done;
}

/* Exact integer Pearson sign for count data: sign(N*Sxy - Sx*Sy), 0 if a gene is constant
 * (np.corrcoef gives NaN there, nan_to_num -> 0: metrics.py:111).  vals_gm int32 [G][N]. */
void gvo_sign_int(const int32_t* vals_gm, int64_t n_cells, int64_t n_genes, int8_t* sign) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < n_genes; ++i) {
    const int32_t* xi = vals_gm + i * n_cells;
    __int128 sx = 0, sxx = 0;
    for (int64_t c = 0; c < n_cells; ++c) { sx += xi[c]; sxx += (__int128)xi[c] * xi[c]; }
    const __int128 vi = (__int128)n_cells * sxx - sx * sx;
    for (int64_t j = i; j < n_genes; ++j) {
      const int32_t* xj = vals_gm + j * n_cells;
      __int128 sy = 0, syy = 0, sxy = 0;
      for (int64_t c = 0; c < n_cells; ++c) {
        sy += xj[c]; syy += (__int128)xj[c] * xj[c]; sxy += (__int128)xi[c] * xj[c];
      }
      const __int128 vj = (__int128)n_cells * syy - sy * sy;
      const __int128 cov = (__int128)n_cells * sxy - sx * sy;
      const int8_t s = (i == j || vi == 0 || vj == 0) ? 0 : (cov > 0 ? 1 : (cov < 0 ? -1 : 0));
      sign[i * n_genes + j] = s;
      sign[j * n_genes + i] = s;
    }
  }
}
