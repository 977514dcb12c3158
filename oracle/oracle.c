/*
 * oracle.c -- CPU restatement of flexgsea's sample-permutation enrichment path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under flexgsea-r_b200/ (the product) may
 * include, link, import or execute this file; only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() use it, and
 * only as the checker / the CPU baseline.
 *
 * Every function follows the reference (NKI-CCB/flexgsea-r) line by line and
 * cites the file:line it restates.  Arithmetic that R performs in 80-bit long
 * double (sum, cumsum, mean, colMeans, var) is done in `long double` here too,
 * so on x86-64 the oracle reproduces R's rounding.
 *
 * Parity pins: tests/golden/weighted_ks.json (decoded from the reference's
 * tests/testthat/test_weighted_ks-*.rds) pins rank + weighted KS + max_es_at +
 * le_prop; the verbatim-compiled src/ggsea.cpp (oracle/_ref) pins orc_s2n_C;
 * test_fdr.R / test_s2n.R / test_lm.R properties pin significance and scores.
 * maxmean, mean and lm have no numeric golden in the reference ("parity
 * unpinned" for those three).
 *
 * All matrices are column-major, exactly as R lays them out.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_ERR_ARG 1
#define ORC_ERR_NONFINITE 3

#define ORC_SCORE_S2N 0
#define ORC_SCORE_LM 1
#define ORC_ES_WKS 0
#define ORC_ES_MAXMEAN 1
#define ORC_ES_MEAN 2

#define NA_LOGICAL INT_MIN

void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------ */
/* a4: s2n_C, src/ggsea.cpp:6-76 (restated; oracle/_ref holds the verbatim    */
/* compile of the reference source this is checked against).                 */
/* x is [n_x x G] (a gene's samples contiguous), y is [n x R] int logical.   */
/* ------------------------------------------------------------------------ */
void orc_s2n_C(const double *x, int n_x, int G, const int *y, int n, int R,
               double *coef /* [G x R] */) {
  for (long i = 0; i < (long)G * R; i++) coef[i] = 0.0;
  if (n != n_x) return; /* src/ggsea.cpp:15-17: zeros on row mismatch */
  for (int ir = 0; ir < R; ir++) {
    const int *yc = y + (long)ir * n;
    for (int ig = 0; ig < G; ig++) {
      const double *xc = x + (long)ig * n;
      double sum1 = 0.0, sum2 = 0.0, mean1, mean2, var1, var2, m;
      int n1 = 0, n2 = 0;
      for (int is = 0; is < n; is++) { /* :28-38 */
        if (yc[is] == NA_LOGICAL) {
        } else if (yc[is]) { sum1 += xc[is]; n1++; }
        else { sum2 += xc[is]; n2++; }
      }
      mean1 = sum1 / n1;
      mean2 = sum1 / n2; /* sic, :40 -- corrected by the second pass */
      sum1 = 0.0; sum2 = 0.0;
      for (int is = 0; is < n; is++) { /* :44-52 */
        if (yc[is] == NA_LOGICAL) {
        } else if (yc[is]) sum1 += xc[is] - mean1;
        else sum2 += xc[is] - mean2;
      }
      mean1 += sum1 / n1;
      mean2 += sum2 / n2;
      sum1 = 0.0; sum2 = 0.0;
      for (int is = 0; is < n; is++) { /* :58-68 */
        if (yc[is] == NA_LOGICAL) {
        } else if (yc[is]) { m = xc[is] - mean1; sum1 += m * m; }
        else { m = xc[is] - mean2; sum2 += m * m; }
      }
      var1 = sum1 / n1; /* population variance, :69-70 */
      var2 = sum2 / n2;
      coef[(long)ir * G + ig] = (mean1 - mean2) / (sqrt(var1) + sqrt(var2));
    }
  }
}

/* a3: flexgsea_s2n post-processing, R/flexgsea.R:715-721.  +-Inf replaced by
 * 1.1 * max/min finite over the whole [G x R] matrix; optional abs. */
void orc_s2n_post(double *coef, long len, int abs_flag) {
  double mx = -INFINITY, mn = INFINITY;
  int any_fin = 0, has_pinf = 0, has_ninf = 0;
  for (long i = 0; i < len; i++) {
    if (isfinite(coef[i])) {
      any_fin = 1;
      if (coef[i] > mx) mx = coef[i];
      if (coef[i] < mn) mn = coef[i];
    } else if (isinf(coef[i])) {
      if (coef[i] > 0) has_pinf = 1; else has_ninf = 1;
    }
  }
  /* R: max(numeric(0)) is -Inf (with a warning), times 1.1 stays -Inf */
  if (!any_fin) { mx = -INFINITY; mn = INFINITY; }
  if (has_pinf) { /* :715 evaluated first */
    double v = mx * 1.1;
    for (long i = 0; i < len; i++) if (isinf(coef[i]) && coef[i] > 0) coef[i] = v;
  }
  if (has_ninf) { /* :716 -- min over finite values AFTER the first patch */
    mn = INFINITY; any_fin = 0;
    for (long i = 0; i < len; i++)
      if (isfinite(coef[i])) { any_fin = 1; if (coef[i] < mn) mn = coef[i]; }
    double v = mn * 1.1;
    for (long i = 0; i < len; i++) if (isinf(coef[i]) && coef[i] < 0) coef[i] = v;
  }
  if (abs_flag) for (long i = 0; i < len; i++) coef[i] = fabs(coef[i]);
}

/* ------------------------------------------------------------------------ */
/* a5: flexgsea_lm, R/flexgsea.R:647-676.  design is the model matrix         */
/* [n x q] (intercept first when has_intercept).  lm.fit == LINPACK          */
/* dqrdc2 + dqrsl (Householder QR, tol 1e-7), restated with plain loops;     */
/* sd() == sqrt(var) with R's two-pass long-double mean (cov.c).             */
/* ------------------------------------------------------------------------ */
static double r_sd(const double *v, int n) {
  long double s = 0.0L;
  for (int i = 0; i < n; i++) s += v[i];
  long double mean = s / n;
  if (isfinite((double)mean)) {
    long double t = 0.0L;
    for (int i = 0; i < n; i++) t += (v[i] - mean);
    mean += t / n;
  }
  long double ss = 0.0L;
  for (int i = 0; i < n; i++) { long double d = v[i] - mean; ss += d * d; }
  return sqrt((double)(ss / (n - 1)));
}

/* gene standard deviations (fixed across permutations: apply(x, 2, sd)) */
void orc_gene_sd(const double *x, int n, int G, double *sd_out) {
#pragma omp parallel for schedule(static)
  for (int g = 0; g < G; g++) sd_out[g] = r_sd(x + (long)g * n, n);
}

/* Householder QR of a [n x q] (column-major, overwritten), LINPACK dqrdc2
 * without the pivoting branch taken (full-rank designs); returns rank. */
static int house_qr(double *a, int n, int q, double *qraux) {
  int rank = q;
  double *colnorm0 = (double *)malloc(sizeof(double) * q);
  for (int j = 0; j < q; j++) {
    double s = 0; for (int i = 0; i < n; i++) s += a[(long)j * n + i] * a[(long)j * n + i];
    colnorm0[j] = sqrt(s);
  }
  for (int l = 0; l < q && l < n; l++) {
    double *cl = a + (long)l * n;
    double nrm = 0; for (int i = l; i < n; i++) nrm += cl[i] * cl[i];
    nrm = sqrt(nrm);
    /* dqrdc2 limited pivoting: a column whose remaining norm fell below
     * tol * original norm is rank deficient */
    if (colnorm0[l] == 0.0 || nrm < 1e-7 * colnorm0[l]) { rank = l; break; }
    if (cl[l] != 0.0) nrm = copysign(nrm, cl[l]);
    for (int i = l; i < n; i++) cl[i] /= nrm;
    cl[l] += 1.0;
    for (int j = l + 1; j < q; j++) {
      double *cj = a + (long)j * n;
      double t = 0; for (int i = l; i < n; i++) t -= cl[i] * cj[i];
      t /= cl[l];
      for (int i = l; i < n; i++) cj[i] += t * cl[i];
    }
    qraux[l] = cl[l];
    cl[l] = -nrm;
  }
  free(colnorm0);
  return rank;
}

/* coefficients of lm.fit(design, xg) for one gene column xg[n]; qr from
 * house_qr.  work[n]. */
static void house_solve(const double *qr, const double *qraux, int n, int q,
                        const double *xg, double *work, double *b) {
  memcpy(work, xg, sizeof(double) * n);
  for (int j = 0; j < q; j++) { /* dqrsl: qty */
    const double *cj = qr + (long)j * n;
    double t = -qraux[j] * work[j];
    for (int i = j + 1; i < n; i++) t -= cj[i] * work[i];
    t /= qraux[j];
    work[j] += t * qraux[j];
    for (int i = j + 1; i < n; i++) work[i] += t * cj[i];
  }
  for (int j = 0; j < q; j++) b[j] = work[j];
  for (int j = q - 1; j >= 0; j--) { /* back substitution */
    b[j] /= qr[(long)j * n + j];
    for (int i = 0; i < j; i++) b[i] -= b[j] * qr[(long)j * n + i];
  }
}

/* scores [G x R], R = q - has_intercept.  Returns ORC_ERR_NONFINITE when the
 * design is rank deficient (lm.fit would give NA coefficients). */
int orc_lm_scores(const double *x, int n, int G, const double *design, int q,
                  int has_intercept, const double *gene_sd, int abs_flag,
                  double *out) {
  int R = q - (has_intercept ? 1 : 0);
  double *qr = (double *)malloc(sizeof(double) * (long)n * q);
  double *qraux = (double *)calloc(q, sizeof(double));
  memcpy(qr, design, sizeof(double) * (long)n * q);
  int rank = house_qr(qr, n, q, qraux);
  int rc = ORC_OK;
  if (rank < q) rc = ORC_ERR_NONFINITE;
  else {
    double *work = (double *)malloc(sizeof(double) * n);
    double *b = (double *)malloc(sizeof(double) * q);
    for (int g = 0; g < G; g++) {
      house_solve(qr, qraux, n, q, x + (long)g * n, work, b);
      for (int r = 0; r < R; r++) {
        double t = b[r + (has_intercept ? 1 : 0)] / gene_sd[g]; /* :671 */
        out[(long)r * G + g] = abs_flag ? fabs(t) : t;
      }
    }
    free(work); free(b);
  }
  free(qr); free(qraux);
  return rc;
}

/* ------------------------------------------------------------------------ */
/* a6: flexgsea_weighted_ks$prepare, R/flexgsea.R:927-933.                    */
/* gene.rank = G - rank(s, 'keep', 'first') + 1.  rank(ties='first') orders  */
/* ascending with ties broken by index (-0.0 == 0.0 is a tie).               */
/* ------------------------------------------------------------------------ */
typedef struct { double v; int i; } orc_vi;
static int cmp_vi(const void *a, const void *b) {
  const orc_vi *p = (const orc_vi *)a, *q = (const orc_vi *)b;
  if (p->v < q->v) return -1;
  if (p->v > q->v) return 1;
  return (p->i > q->i) - (p->i < q->i);
}
void orc_rank_desc(const double *s, int G, int *rank_out /* 1-based desc */) {
  orc_vi *t = (orc_vi *)malloc(sizeof(orc_vi) * G);
  for (int i = 0; i < G; i++) { t[i].v = s[i]; t[i].i = i; }
  qsort(t, G, sizeof(orc_vi), cmp_vi);
  for (int q = 0; q < G; q++) rank_out[t[q].i] = G - (q + 1) + 1;
  free(t);
}

/* ------------------------------------------------------------------------ */
/* a7 + a8: flexgsea_weighted_ks_, R/flexgsea.R:781-910, one (response, perm) */
/* column.  set[] holds 1-based gene indices (duplicates allowed).           */
/* Optional outputs (NULL to skip): max_es_at, le_prop (:897-902),           */
/* order (running_es_at, :883-885; 1-based gene ids in rank order),          */
/* run_pos/run_neg (:877-882), le_first/le_count: leading edge is            */
/* order[le_first .. le_first+le_count) (:886-896).                          */
/* ------------------------------------------------------------------------ */
typedef struct { int r; int j; } orc_rj;
static int cmp_rj(const void *a, const void *b) {
  const orc_rj *p = (const orc_rj *)a, *q = (const orc_rj *)b;
  if (p->r != q->r) return (p->r > q->r) - (p->r < q->r);
  return (p->j > q->j) - (p->j < q->j); /* order() is stable */
}
double orc_wks_column(const double *score /* [G] */, const int *rank /* [G] */,
                      int G, const int *set, int m, double p,
                      double *max_es_at, double *le_prop, int *order_out,
                      double *run_pos, double *run_neg, int *le_first,
                      int *le_count) {
  double *w = (double *)malloc(sizeof(double) * m);
  orc_rj *o = (orc_rj *)malloc(sizeof(orc_rj) * m);
  long double sw = 0.0L;
  for (int j = 0; j < m; j++) {
    double a = fabs(score[set[j] - 1]);
    w[j] = (p == 1.0) ? a : pow(a, p); /* :819 */
    sw += w[j];
  }
  double sumw = (double)sw;
  for (int j = 0; j < m; j++) { w[j] = w[j] / sumw; /* :820 */
    o[j].r = rank[set[j] - 1]; o[j].j = j; }
  qsort(o, m, sizeof(orc_rj), cmp_rj); /* :822 */
  long double cum = 0.0L;
  double es_pos = -INFINITY, es_neg = INFINITY;
  int wmax = -1, wmin = -1;
  double denom = (double)(G - m);
  for (int k = 0; k < m; k++) {
    double wo = w[o[k].j];
    cum += wo;                                   /* :824 cumsum in long double */
    double d_hit = (double)cum;
    double d_miss = ((double)o[k].r - (double)(k + 1)) / denom; /* :825-826 */
    double rp = d_hit - d_miss;                  /* :828 */
    double rn = rp - wo;                         /* :829 */
    if (run_pos) run_pos[k] = rp;
    if (run_neg) run_neg[k] = rn;
    if (order_out) order_out[k] = set[o[k].j];
    /* max()/min() with NaN: R returns NaN if any NaN present */
    if (isnan(rp)) { es_pos = NAN; } else if (!isnan(es_pos) && rp > es_pos) { es_pos = rp; wmax = k; }
    if (isnan(rn)) { es_neg = NAN; } else if (!isnan(es_neg) && rn < es_neg) { es_neg = rn; wmin = k; }
  }
  double es;
  int do_pos = es_pos > -es_neg; /* :833 strict */
  if (isnan(es_pos) || isnan(es_neg)) { es = NAN; do_pos = 0; }
  else es = do_pos ? es_pos : es_neg;
  int wmes = do_pos ? wmax : wmin;
  if (wmes < 0) wmes = 0;
  if (max_es_at) *max_es_at = (m > 0) ? (double)o[wmes].r : NAN;   /* :897-899 */
  int first = do_pos ? 0 : wmes, count = do_pos ? wmes + 1 : m - wmes; /* :889,:894 */
  if (le_first) *le_first = first;
  if (le_count) *le_count = count;
  if (le_prop) *le_prop = (double)count / (double)m;               /* :900-902 */
  free(w); free(o);
  return es;
}

/* a9: flexgsea_maxmean_, R/flexgsea.R:725-741, one column */
double orc_maxmean_column(const double *score, const int *set, int m) {
  long double sp = 0.0L, sn = 0.0L;
  for (int j = 0; j < m; j++) {
    double s = score[set[j] - 1], a = fabs(s);
    sp += (s + a) / 2;   /* :734 */
    sn += (-s + a) / 2;  /* :735 */
  }
  double pos = (double)(sp / m), neg = (double)(sn / m);
  double es = pos > neg ? pos : neg; /* pmax :736 */
  if (isnan(pos) || isnan(neg)) return NAN;
  if (neg > pos) es = -1 * es;       /* :737 strict */
  return es;
}
/* a9: flexgsea_mean_, R/flexgsea.R:755-767, one column */
double orc_mean_column(const double *score, const int *set, int m) {
  long double s = 0.0L;
  for (int j = 0; j < m; j++) s += score[set[j] - 1];
  return (double)(s / m); /* colMeans :763 */
}

/* ES for every set of one score column; es_out[S] */
void orc_es_column(int es_kind, double p, const double *score, int G,
                   const int *offsets, const int *index, int S, double *es_out) {
  int *rank = NULL;
  if (es_kind == ORC_ES_WKS) {
    rank = (int *)malloc(sizeof(int) * G);
    orc_rank_desc(score, G, rank);
  }
  for (int s = 0; s < S; s++) {
    const int *set = index + offsets[s];
    int m = offsets[s + 1] - offsets[s];
    if (es_kind == ORC_ES_WKS)
      es_out[s] = orc_wks_column(score, rank, G, set, m, p, 0, 0, 0, 0, 0, 0, 0);
    else if (es_kind == ORC_ES_MAXMEAN) es_out[s] = orc_maxmean_column(score, set, m);
    else es_out[s] = orc_mean_column(score, set, m);
  }
  free(rank);
}

/* ES from a ready score array [G, R, P] (user-defined gene.score.fn path):
 * es_out [S, R, P].  R/flexgsea.R:410-415. */
int orc_es_from_scores(int es_kind, double p, const double *scores, int G, int R,
                       int P, const int *offsets, const int *index, int S,
                       double *es_out) {
  long C = (long)R * P;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 4)
  for (long c = 0; c < C; c++) {
    const double *col = scores + c * G;
    for (int g = 0; g < G; g++) if (!isfinite(col[g])) { bad = 1; }
    orc_es_column(es_kind, p, col, G, offsets, index, S, es_out + c * S);
  }
  return bad ? ORC_ERR_NONFINITE : ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* K1 definition shared with the device: Philox4x32-10 Fisher-Yates.         */
/* perm_out[i] (1-based) for i in [0,n).  Counter = (chunk, 0, id_lo, id_hi),*/
/* key = (seed_lo, seed_hi); draw j in [0,i] as mulhi(u32, i+1).             */
/* ------------------------------------------------------------------------ */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
void orc_philox_perm(uint64_t seed, uint64_t perm_id, int n, int *perm_out) {
  for (int i = 0; i < n; i++) perm_out[i] = i + 1;
  uint32_t buf[4]; int have = 0; uint32_t chunk = 0;
  for (int i = n - 1; i >= 1; i--) {
    if (!have) {
      buf[0] = chunk++; buf[1] = 0; buf[2] = (uint32_t)perm_id; buf[3] = (uint32_t)(perm_id >> 32);
      philox4x32_10(buf, (uint32_t)seed, (uint32_t)(seed >> 32));
      have = 4;
    }
    uint32_t u = buf[4 - have]; have--;
    uint32_t j = (uint32_t)(((uint64_t)u * (uint64_t)(i + 1)) >> 32);
    int t = perm_out[i]; perm_out[i] = perm_out[j]; perm_out[j] = t;
  }
}

/* ------------------------------------------------------------------------ */
/* a1 + a2: flexgsea_perm_sequential, R/flexgsea.R:356-420.                   */
/* perm is [n x P] int32, 1-based: y.perm[i,] = y[perm[i,p],] (:384).         */
/* score_kind S2N: resp = int y_bin [n x ncol]; LM: resp = double design     */
/* [n x ncol] (has_intercept => first column dropped from the scores).       */
/* es_null [S, R, P].  Returns ORC_ERR_NONFINITE as the guard at :400-402.   */
/* ------------------------------------------------------------------------ */
int orc_perm_null(const double *x, int n, int G, int score_kind, const void *resp,
                  int ncol, int has_intercept, const int *perm, int P,
                  const int *offsets, const int *index, int S, int es_kind,
                  double p, int abs_flag, double *es_null, double *scores_out) {
  int R = (score_kind == ORC_SCORE_S2N) ? ncol : ncol - (has_intercept ? 1 : 0);
  double *gsd = NULL;
  if (score_kind == ORC_SCORE_LM) {
    gsd = (double *)malloc(sizeof(double) * G);
    orc_gene_sd(x, n, G, gsd); /* recomputed per permutation in R (:671); same value */
  }
  int bad = 0;
#pragma omp parallel
  {
    double *sc = (double *)malloc(sizeof(double) * (long)G * R);
    void *yp = malloc((score_kind == ORC_SCORE_S2N ? sizeof(int) : sizeof(double)) * (long)n * ncol);
#pragma omp for schedule(dynamic, 1)
    for (int ip = 0; ip < P; ip++) {
      const int *pi = perm + (long)ip * n;
      int rc = ORC_OK;
      if (score_kind == ORC_SCORE_S2N) {
        const int *y = (const int *)resp; int *y2 = (int *)yp;
        for (int c = 0; c < ncol; c++)
          for (int i = 0; i < n; i++) y2[(long)c * n + i] = y[(long)c * n + pi[i] - 1];
        orc_s2n_C(x, n, G, y2, n, ncol, sc);
        orc_s2n_post(sc, (long)G * R, abs_flag);
      } else {
        const double *d = (const double *)resp; double *d2 = (double *)yp;
        for (int c = 0; c < ncol; c++)
          for (int i = 0; i < n; i++) d2[(long)c * n + i] = d[(long)c * n + pi[i] - 1];
        rc = orc_lm_scores(x, n, G, d2, ncol, has_intercept, gsd, abs_flag, sc);
      }
      if (rc != ORC_OK) bad = 1;
      for (long i = 0; i < (long)G * R; i++) if (!isfinite(sc[i])) bad = 1; /* :400-402 */
      if (scores_out) memcpy(scores_out + (long)ip * G * R, sc, sizeof(double) * (long)G * R);
      for (int r = 0; r < R; r++)
        orc_es_column(es_kind, p, sc + (long)r * G, G, offsets, index, S,
                      es_null + ((long)ip * R + r) * S);
    }
    free(sc); free(yp);
  }
  free(gsd);
  return bad ? ORC_ERR_NONFINITE : ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* a10-a12: significance, R/flexgsea.R:474-629.  es[S], es_null[S x P].      */
/* ------------------------------------------------------------------------ */
static double r_mean_ld(const double *v, long n) { /* R's mean(): summary.c */
  if (n == 0) return NAN;
  long double s = 0.0L;
  for (long i = 0; i < n; i++) s += v[i];
  s /= n;
  if (isfinite((double)s)) {
    long double t = 0.0L;
    for (long i = 0; i < n; i++) t += (v[i] - s);
    s += t / n;
  }
  return (double)s;
}

/* adj_fdr_nes :474-498 */
static void adj_fdr_nes(double *fdr, const double *nes, int S) {
  orc_vi *t = (orc_vi *)malloc(sizeof(orc_vi) * S);
  int np = 0;
  for (int i = 0; i < S; i++) if (nes[i] >= 0) { t[np].v = nes[i]; t[np].i = i; np++; }
  qsort(t, np, sizeof(orc_vi), cmp_vi); /* order(decreasing=F), stable by index */
  double mn = 1.0;
  for (int k = 0; k < np; k++) { int i = t[k].i;
    if (fdr[i] <= mn) mn = fdr[i]; else fdr[i] = mn; }
  int nn = 0;
  for (int i = 0; i < S; i++) if (nes[i] < 0) { t[nn].v = -nes[i]; t[nn].i = i; nn++; }
  /* order(nes, decreasing=T): descending value, ties keep index order
   * (R's order is stable for decreasing too) == ascending of -nes, index asc */
  qsort(t, nn, sizeof(orc_vi), cmp_vi);
  mn = 1.0;
  for (int k = 0; k < nn; k++) { int i = t[k].i;
    if (fdr[i] <= mn) mn = fdr[i]; else fdr[i] = mn; }
  free(t);
}

/* calc_fdr_nes :500-533; counts_out (optional) [S x 2] = per-set
 * (#{nes cmp nes_i}, #{nes_null cmp nes_i}); glob_out[4] = the four global
 * counts of :501-504. */
static void calc_fdr_nes(const double *nes, const double *nes_null, int S, long P,
                         int abs_flag, double *fdr, long long *counts_out,
                         long long *glob_out) {
  long long cp = 0, cn = 0, ncp = 0, ncn = 0;
  for (int i = 0; i < S; i++) { if (nes[i] > 0) cp++; if (nes[i] < 0) cn++; }
  long SP = (long)S * P;
  for (long i = 0; i < SP; i++) { if (nes_null[i] > 0) ncp++; if (nes_null[i] < 0) ncn++; }
  if (glob_out) { glob_out[0] = cp; glob_out[1] = cn; glob_out[2] = ncp; glob_out[3] = ncn; }
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < S; i++) {
    double v = nes[i];
    long long a = 0, b = 0;
    double rel, rel_null;
    if ((v >= 0) || abs_flag) {
      for (int j = 0; j < S; j++) a += nes[j] > v;
      for (long j = 0; j < SP; j++) b += nes_null[j] > v;
      rel = (double)(a + 1) / (double)cp;
      rel_null = (double)(b + 1) / (double)ncp;
    } else {
      for (int j = 0; j < S; j++) a += nes[j] < v;
      for (long j = 0; j < SP; j++) b += nes_null[j] < v;
      rel = (double)(a + 1) / (double)cn;
      rel_null = (double)(b + 1) / (double)ncn;
    }
    if (isnan(v)) { rel = NAN; }
    fdr[i] = rel_null / rel;
    if (counts_out) { counts_out[2 * i] = a; counts_out[2 * i + 1] = b; }
  }
  adj_fdr_nes(fdr, nes, S);
}

/* p.adjust(p, 'BH') (stats::p.adjust) */
static int cmp_vi_pdesc(const void *a, const void *b) {
  const orc_vi *p = (const orc_vi *)a, *q = (const orc_vi *)b;
  if (p->v > q->v) return -1;
  if (p->v < q->v) return 1;
  return (p->i > q->i) - (p->i < q->i);
}
static void p_adjust_bh(const double *p, int n, double *out) {
  orc_vi *t = (orc_vi *)malloc(sizeof(orc_vi) * n);
  for (int i = 0; i < n; i++) { t[i].v = p[i]; t[i].i = i; }
  qsort(t, n, sizeof(orc_vi), cmp_vi_pdesc); /* o = order(p, decreasing=TRUE) */
  double cm = INFINITY;
  for (int k = 0; k < n; k++) {            /* i = n:1 */
    double v = ((double)n / (double)(n - k)) * t[k].v;
    if (v < cm) cm = v;
    out[t[k].i] = cm < 1.0 ? cm : 1.0;
  }
  free(t);
}

/*
 * flexgsea_calc_sig :564-629 (split_p=1, calc_nes=1) and
 * flexgsea_calc_sig_simple :544-547 (split_p=0, calc_nes=0).
 * Outputs (each [S], NULL to skip): p, p_high, p_low, nes, fdr, fwer.
 * pcounts_out (optional) [S x 2]: numerator-1 and denominator-1 of p (or of
 * p.high / p.low in the simple variant); fdr_counts_out [S x 2],
 * glob_counts_out[4]; nes_null_out [S x P] optional.
 */
int orc_calc_sig(const double *es, const double *es_null, int S, long P,
                 int split_p, int calc_nes, int abs_flag, double *p_out,
                 double *p_high, double *p_low, double *nes_out, double *fdr_out,
                 double *fwer_out, long long *pcounts_out,
                 long long *fdr_counts_out, long long *glob_counts_out,
                 double *nes_null_out) {
  if (P == 0) return ORC_OK; /* :575-577 */
  double *pv = (double *)malloc(sizeof(double) * S);
  if (split_p) {
    for (int s = 0; s < S; s++) { /* :579-587 */
      long long num = 0, den = 0;
      if ((es[s] >= 0.0) || abs_flag) {
        for (long j = 0; j < P; j++) { double v = es_null[s + j * S];
          if (v >= 0.0) { den++; if (es[s] <= v) num++; } }
      } else {
        for (long j = 0; j < P; j++) { double v = es_null[s + j * S];
          if (v <= 0.0) { den++; if (es[s] >= v) num++; } }
      }
      pv[s] = (double)(num + 1) / (double)(den + 1);
      if (pcounts_out) { pcounts_out[2 * s] = num; pcounts_out[2 * s + 1] = den; }
    }
  } else {
    for (int s = 0; s < S; s++) { /* :589-599 */
      long long hi = 0, lo = 0;
      for (long j = 0; j < P; j++) { double v = es_null[s + j * S];
        hi += es[s] <= v; lo += es[s] >= v; }
      double ph = (double)(hi + 1) / (double)(P + 1), pl = (double)(lo + 1) / (double)(P + 1);
      if (p_high) p_high[s] = ph;
      if (abs_flag) pv[s] = ph;
      else { if (p_low) p_low[s] = pl;
        double mn = pl < ph ? pl : ph; mn *= 2; pv[s] = mn < 1.0 ? mn : 1.0; }
      if (pcounts_out) { pcounts_out[2 * s] = hi; pcounts_out[2 * s + 1] = lo; }
    }
  }
  if (p_out) memcpy(p_out, pv, sizeof(double) * S);
  if (calc_nes) {
    double *mpos = (double *)malloc(sizeof(double) * S);
    double *mneg = (double *)malloc(sizeof(double) * S);
    double *tmp = (double *)malloc(sizeof(double) * P);
    double *nes = (double *)malloc(sizeof(double) * S);
    for (int s = 0; s < S; s++) {
      long k = 0;
      for (long j = 0; j < P; j++) { double v = es_null[s + j * S]; if (v >= 0) tmp[k++] = v; }
      mpos[s] = r_mean_ld(tmp, k);                         /* :602 */
      if (isnan(mpos[s]) && es[s] >= 0.0) mpos[s] = es[s]; /* :603-604 */
      k = 0;
      for (long j = 0; j < P; j++) { double v = es_null[s + j * S]; if (v < 0) tmp[k++] = v; }
      mneg[s] = r_mean_ld(tmp, k);                         /* :612-614 */
      if (isnan(mneg[s]) && es[s] < 0.0) mneg[s] = es[s];  /* :615-616 */
    }
    double *nn = nes_null_out ? nes_null_out : (double *)malloc(sizeof(double) * (long)S * P);
    for (int s = 0; s < S; s++) {
      if (abs_flag) nes[s] = es[s] / mpos[s];              /* :607 */
      else nes[s] = es[s] / (es[s] >= 0 ? mpos[s] : -mneg[s]); /* :617-618 */
    }
    for (long j = 0; j < P; j++)
      for (int s = 0; s < S; s++) {
        double v = es_null[s + j * S];
        if (abs_flag) nn[s + j * S] = v / mpos[s];         /* :608-610 */
        else nn[s + j * S] = v / (v >= 0 ? mpos[s] : -mneg[s]); /* :619-621 */
      }
    double *fdr = (double *)malloc(sizeof(double) * S);
    calc_fdr_nes(nes, nn, S, P, abs_flag, fdr, fdr_counts_out, glob_counts_out); /* :623 */
    if (nes_out) memcpy(nes_out, nes, sizeof(double) * S);
    if (fdr_out) memcpy(fdr_out, fdr, sizeof(double) * S);
    if (!nes_null_out) free(nn);
    free(fdr); free(mpos); free(mneg); free(tmp); free(nes);
  } else if (fdr_out) {
    p_adjust_bh(pv, S, fdr_out); /* :625 */
  }
  if (fwer_out) for (int s = 0; s < S; s++) { /* :627 bonferroni */
    double v = pv[s] * S; fwer_out[s] = v < 1.0 ? v : 1.0; }
  free(pv);
  return ORC_OK;
}
