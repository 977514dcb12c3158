// C ABI of the flexgsea sm_100a library: context management, uploads and the
// orchestration of the per-block pipeline
//   permute labels -> gene scores -> rank -> ES        (fgs_perm_null)
//   per-set null reductions -> NES -> pooled FDR        (fgs_calc_sig / staged)
// See include/flexgsea_b200.h for the contract and the reference citations.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <climits>
#include <thread>
#include <unordered_map>

#include "fgs_common.cuh"

namespace fgs {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int launch_sig_means(fgs_ctx *c, const double *d_sums_parts, int nparts, size_t stride, const long long *d_counts,
                     const double *d_es, int S, long long P_total, int split_p, int calc_nes,
                     int abs_flag, double *d_p, double *d_ph, double *d_pl, double *d_mpos,
                     double *d_mneg, double *d_nes, double *d_fwer);
int launch_sig_sum_counts(fgs_ctx *c, const double *d_parts, int nparts, size_t stride, int S, long long *d_counts);
int launch_bh(fgs_ctx *c, const double *d_p, int S, double *d_fdr);

}  // namespace fgs (reopened below)
// a tile still in flight (fgs_es_from_scores_tile) finishes before anything else touches the context
int fgs::settle_tile(fgs_ctx *c) {
  if (c->tile_job.joinable()) c->tile_job.join();
  if (c->tile_rc != FGS_OK) {
    const int rc = c->tile_rc;
    c->tile_rc = FGS_OK;
    set_error("%s", c->tile_err.c_str());
    return rc;
  }
  return FGS_OK;
}
namespace fgs {
static int use_device(fgs_ctx *c) {
  FGS_TRY(settle_tile(c));
  FGS_CUDA(cudaSetDevice(c->device));
  return FGS_OK;
}

template <typename T>
static int h2d(fgs_ctx *c, T *dst, const T *src, size_t n) {
  if (n == 0) return FGS_OK;
  FGS_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  return FGS_OK;
}
template <typename T>
static int d2h(fgs_ctx *c, T *dst, const T *src, size_t n) {
  if (n == 0 || dst == nullptr) return FGS_OK;
  FGS_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
  return FGS_OK;
}
static int sync(fgs_ctx *c) {
  FGS_CUDA(cudaStreamSynchronize(c->stream));
  return FGS_OK;
}

// Host -> device upload of a large pageable array (user score columns).  A plain copy from
// pageable memory runs at the pace of the driver's single staging thread (about 10 GB/s
// measured); here `T` host threads each copy 8 MB pieces into their own two pinned staging
// slots and queue the DMA of a piece as soon as it is staged, so the staging memcpy of one
// thread overlaps the others' and the link.  Pieces land in any order; the caller records its
// event on `st` after this returns (every DMA is queued by then).
static const size_t STAGE_PIECE = 8u << 20;
static bool stage_ring(fgs_ctx *c) {
  if (c->stage) return true;
  const size_t need = (size_t)16 * STAGE_PIECE;
  if (cudaHostAlloc((void **)&c->stage, need, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    c->stage = nullptr;
    return false;
  }
  c->stage_bytes = need;
  for (int i = 0; i < 16; i++)
    if (cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      cudaFreeHost(c->stage);
      c->stage = nullptr;
      return false;
    }
  return true;
}
static int stage_threads(fgs_ctx *c) {
  int T = std::max(1, std::min(c->opt_h2d_threads, 8));
  const unsigned hc = std::thread::hardware_concurrency();
  if (hc && (unsigned)T > hc) T = (int)hc;
  return T;
}
static int staged_upload(fgs_ctx *c, void *dst, const void *src, size_t bytes, cudaStream_t st) {
  const int T = stage_threads(c);
  if (bytes < 4 * STAGE_PIECE || c->opt_h2d_threads <= 0) {
    FGS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return FGS_OK;
  }
  if (!stage_ring(c)) {  // no pinned memory to be had: the plain copy still works
    FGS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return FGS_OK;
  }
  const size_t n_piece = (bytes + STAGE_PIECE - 1) / STAGE_PIECE;
  std::vector<int> status(T, 0);
  auto worker = [&](int t) {
    if (cudaSetDevice(c->device) != cudaSuccess) { status[t] = 1; return; }
    int turn = 0;
    for (size_t i = (size_t)t; i < n_piece; i += (size_t)T, turn ^= 1) {
      const int slot = 2 * t + turn;
      char *buf = c->stage + (size_t)slot * STAGE_PIECE;
      const size_t o = i * STAGE_PIECE, len = std::min(STAGE_PIECE, bytes - o);
      // the slot's previous piece (of this call or an earlier one) must have left; an event never
      // recorded counts as complete
      if (cudaEventSynchronize(c->stage_ev[slot]) != cudaSuccess) { status[t] = 1; return; }
      memcpy(buf, (const char *)src + o, len);
      if (cudaMemcpyAsync((char *)dst + o, buf, len, cudaMemcpyHostToDevice, st) != cudaSuccess ||
          cudaEventRecord(c->stage_ev[slot], st) != cudaSuccess) { status[t] = 1; return; }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back(worker, t);
  worker(0);
  for (auto &x : th) x.join();
  for (int t = 0; t < T; t++)
    if (status[t]) { cudaStreamSynchronize(st); set_error("staged upload failed: %s", cudaGetErrorString(cudaGetLastError())); return FGS_ERR_CUDA; }
  return FGS_OK;
}

// The way back (a block of es.null into the caller's pageable array): the mirror image.  Every
// thread keeps two pieces in flight -- DMA into one pinned slot while it copies the other out --
// and the call returns when the data is in `dst`.  `st` must already be ordered after the
// kernels that produce `src`.
static int staged_download(fgs_ctx *c, void *dst, const void *src, size_t bytes, cudaStream_t st) {
  if (bytes == 0 || !dst) return FGS_OK;
  const int T = stage_threads(c);
  if (bytes < 4 * STAGE_PIECE || c->opt_h2d_threads <= 0 || !stage_ring(c)) {
    FGS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    FGS_CUDA(cudaStreamSynchronize(st));
    return FGS_OK;
  }
  const size_t n_piece = (bytes + STAGE_PIECE - 1) / STAGE_PIECE;
  std::vector<int> status(T, 0);
  auto worker = [&](int t) {
    if (cudaSetDevice(c->device) != cudaSuccess) { status[t] = 1; return; }
    size_t pend_o[2] = {0, 0}, pend_len[2] = {0, 0};
    auto finish = [&](int turn) -> bool {  // wait for the slot's DMA, copy the piece out
      const int slot = 2 * t + turn;
      if (cudaEventSynchronize(c->stage_ev[slot]) != cudaSuccess) return false;
      if (pend_len[turn]) memcpy((char *)dst + pend_o[turn], c->stage + (size_t)slot * STAGE_PIECE, pend_len[turn]);
      pend_len[turn] = 0;
      return true;
    };
    int turn = 0;
    for (size_t i = (size_t)t; i < n_piece; i += (size_t)T, turn ^= 1) {
      const int slot = 2 * t + turn;
      if (!finish(turn)) { status[t] = 1; return; }
      const size_t o = i * STAGE_PIECE, len = std::min(STAGE_PIECE, bytes - o);
      if (cudaMemcpyAsync(c->stage + (size_t)slot * STAGE_PIECE, (const char *)src + o, len, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaEventRecord(c->stage_ev[slot], st) != cudaSuccess) { status[t] = 1; return; }
      pend_o[turn] = o; pend_len[turn] = len;
    }
    if (!finish(turn) || !finish(turn ^ 1)) status[t] = 1;   // oldest first
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back(worker, t);
  worker(0);
  for (auto &x : th) x.join();
  for (int t = 0; t < T; t++)
    if (status[t]) { cudaStreamSynchronize(st); set_error("staged download failed: %s", cudaGetErrorString(cudaGetLastError())); return FGS_ERR_CUDA; }
  return FGS_OK;
}

// es.null of block `k` (recorded as finished in blk_out[k & 1]) into the caller's array
static int download_block(fgs_ctx *c, int slot, double *dst, const double *src, size_t n) {
  FGS_CUDA(cudaStreamWaitEvent(c->copy_stream, c->blk_out[slot], 0));
  return staged_download(c, dst, src, n * sizeof(double), c->copy_stream);
}

// Householder QR of the [n x q] design (column-major), then H = R^-1 Q^T.
// Returns false when rank deficient (tolerance as LINPACK dqrdc2 in lm.fit).
static bool design_pinv(const std::vector<double> &D, int n, int q, std::vector<double> &H) {
  std::vector<double> a(D), qraux(q, 0.0), norm0(q, 0.0);
  for (int j = 0; j < q; j++) {
    double s = 0;
    for (int i = 0; i < n; i++) s += a[(size_t)j * n + i] * a[(size_t)j * n + i];
    norm0[j] = std::sqrt(s);
  }
  if (q > n) return false;
  for (int l = 0; l < q; l++) {
    double *cl = &a[(size_t)l * n];
    double nrm = 0;
    for (int i = l; i < n; i++) nrm += cl[i] * cl[i];
    nrm = std::sqrt(nrm);
    if (norm0[l] == 0.0 || nrm < 1e-7 * norm0[l]) return false;
    if (cl[l] != 0.0) nrm = std::copysign(nrm, cl[l]);
    for (int i = l; i < n; i++) cl[i] /= nrm;
    cl[l] += 1.0;
    for (int j = l + 1; j < q; j++) {
      double *cj = &a[(size_t)j * n];
      double t = 0;
      for (int i = l; i < n; i++) t -= cl[i] * cj[i];
      t /= cl[l];
      for (int i = l; i < n; i++) cj[i] += t * cl[i];
    }
    qraux[l] = cl[l];
    cl[l] = -nrm;
  }
  H.assign((size_t)q * n, 0.0);  // H[c * n + i]
  std::vector<double> w(n), b(q);
  for (int e = 0; e < n; e++) {  // column e of H = coefficients of lm.fit(D, unit_e)
    std::fill(w.begin(), w.end(), 0.0);
    w[e] = 1.0;
    for (int j = 0; j < q; j++) {
      const double *cj = &a[(size_t)j * n];
      double t = -qraux[j] * w[j];
      for (int i = j + 1; i < n; i++) t -= cj[i] * w[i];
      t /= qraux[j];
      w[j] += t * qraux[j];
      for (int i = j + 1; i < n; i++) w[i] += t * cj[i];
    }
    for (int j = 0; j < q; j++) b[j] = w[j];
    for (int j = q - 1; j >= 0; j--) {
      b[j] /= a[(size_t)j * n + j];
      for (int i = 0; i < j; i++) b[i] -= b[j] * a[(size_t)j * n + i];
    }
    for (int j = 0; j < q; j++) H[(size_t)j * n + e] = b[j];
  }
  return true;
}

static int n_response_of(fgs_ctx *c, int score_kind) {
  if (score_kind == FGS_SCORE_S2N) return c->s2n_R;
  if (score_kind == FGS_SCORE_LM) return c->lm_R;
  return 0;
}

static int check_problem(fgs_ctx *c, bool need_sets) {
  if (c->G <= 0 || c->n <= 0) { set_error("fgs_set_x has not been called"); return FGS_ERR_STATE; }
  if (need_sets) {
    if (c->S <= 0) { set_error("fgs_set_gene_sets has not been called"); return FGS_ERR_STATE; }
    if (c->max_set > c->G) { set_error("gene set index %d exceeds the %d genes of x", c->max_set, c->G); return FGS_ERR_ARG; }
  }
  return FGS_OK;
}

static constexpr int CERT_OVERFLOW = 1000;  // internal
// flags buffer: [0, P) per-permutation bits, [n-1] global bits
static int flags_reset(fgs_ctx *c, long long P) {
  FGS_TRY(c->flags.ensure((size_t)P + 1));
  FGS_CUDA(cudaMemsetAsync(c->flags.p, 0, c->flags.n * sizeof(int), c->stream));
  return FGS_OK;
}
static int flags_check(fgs_ctx *c, long long P) {
  std::vector<int> h(c->flags.n);
  FGS_CUDA(cudaMemcpyAsync(h.data(), c->flags.p, c->flags.n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  FGS_TRY(sync(c));
  if (h[c->flags.n - 1] & 8) { set_error("permutation matrix holds an index outside [1, n_sample]"); return FGS_ERR_ARG; }
  if (h[c->flags.n - 1] & 16) return CERT_OVERFLOW;  // certification lists full: caller reruns bit-faithfully
  for (long long p = 0; p < P; p++)
    if (h[p]) {
      set_error("Gene scoring function returned NA or Inf.");  // R/flexgsea.R:400-402
      return FGS_ERR_NONFINITE;
    }
  return FGS_OK;
}

struct StageTimer {
  fgs_ctx *c;
  std::vector<cudaEvent_t> evs;  // 4 per block: start, scored, ranked, es
  explicit StageTimer(fgs_ctx *ctx) : c(ctx) {}
  void mark() {
    if (!c->profiling) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    evs.push_back(e);
  }
  void drop() {
    for (auto e : evs) cudaEventDestroy(e);
    evs.clear();
  }
  void finish(float prologue_ms) {
    for (int i = 0; i < 5; i++) c->last_ms[i] = 0;
    if (!c->profiling) return;
    c->last_ms[0] = prologue_ms;
    for (size_t b = 0; b + 3 < evs.size(); b += 4) {
      float t;
      cudaEventElapsedTime(&t, evs[b], evs[b + 1]); c->last_ms[1] += t;
      cudaEventElapsedTime(&t, evs[b + 1], evs[b + 2]); c->last_ms[2] += t;
      cudaEventElapsedTime(&t, evs[b + 2], evs[b + 3]); c->last_ms[3] += t;
    }
    for (auto e : evs) cudaEventDestroy(e);
    evs.clear();
  }
};

// columns (score vectors) processed per block: bounded so the scratch
// (scores 8G + rank 2G + weights 8G bytes per column) stays ~<= 1.5 GB
static long long block_columns(fgs_ctx *c, long long C_total) {
  long long cap = (long long)(3.0e9 / (18.0 * (double)c->G + 64));
  if (cap < 1) cap = 1;
  long long cb = c->opt_perm_block;
  if (cb <= 0) {
    // the rank and ES kernels are persistent (a CTA per SM takes a column at a time): a
    // block that is a whole number of rounds over the SMs has no idle tail
    cb = 4096;
    if (cb > cap) cb = cap;
    if (cb > c->num_sms) cb = ((cb + c->num_sms / 2) / c->num_sms) * c->num_sms;
    if (cb > cap) cb -= c->num_sms;
  }
  if (cb > cap) cb = cap;
  if (cb > C_total) cb = C_total;
  return cb < 1 ? 1 : cb;
}

// A problem small enough that computing EVERY weighted-KS unit in the reference's own
// operation order (es_wks_tie_kernel, O(m^2 / 32) per unit) costs next to nothing.  Worth
// doing because small problems are where it matters: with a dozen genes or four samples the
// statistic takes a handful of values, NES of different sets tie up to rounding, and the
// pooled-FDR counts then hang on the last bit of every null value, not only of those near
// the observed ES.
static bool small_wks_problem(const fgs_ctx *c, long long C) {
  return c->opt_exact_small && c->sum_m2 * (double)C * (25.0 / 32.0) <= 5e8;
}
// the same for maxmean / mean (launch_es_mean_exact: double-double sums, ~10 instructions per member)
static bool small_mean_problem(const fgs_ctx *c, long long C) {
  return c->opt_exact_small && (double)c->nnz * (double)C * (10.0 / 32.0) <= 5e8;
}
static bool small_problem(const fgs_ctx *c, int es_kind, long long C) {
  return es_kind == FGS_ES_WEIGHTED_KS ? small_wks_problem(c, C) : small_mean_problem(c, C);
}

// Few samples: a permutation only matters through the class labels it gives every response, and
// there are at most prod_r n! / (n1_r! n2_r! nNA_r!) distinct labellings -- 6 for 2 + 2 samples,
// 20 for 3 + 3, 924 for 6 + 6.  When that is well below the number of permutations, each distinct
// labelling is computed ONCE and its column copied to the permutations that share it
// (fgs_perm_null).  Faster, and permutations that repeat a labelling get identical values by
// construction -- what the `<=` / `>=` / `>` counts of flexgsea_calc_sig and calc_fdr_nes
// (R/flexgsea.R:578-587, :500-533) see in the reference, which runs the same code on the same data.
static double distinct_labellings(const fgs_ctx *c) {
  if (c->s2n_n <= 0 || c->s2n_n > 32 || c->h_n1n2.empty()) return INFINITY;
  double prod = 1.0;
  for (int r = 0; r < c->s2n_R; r++) {
    const int n1 = c->h_n1n2[2 * r], n2 = c->h_n1n2[2 * r + 1], n = c->s2n_n;
    double m = 1.0;  // n! / (n1! n2! (n - n1 - n2)!) = C(n, n1) * C(n - n1, n2)
    for (int i = 1; i <= n1; i++) m = m * (double)(n - n1 + i) / (double)i;
    for (int i = 1; i <= n2; i++) m = m * (double)(n - n1 - n2 + i) / (double)i;
    prod *= m;
    if (prod > 1e15) return INFINITY;
  }
  return prod;
}
// With the labellings deduplicated the all-exact path (every unit in the reference's operation
// order) is affordable for far larger problems: budget ~2e10 warp instructions (tens of ms).
static bool small_problem_dedup(const fgs_ctx *c, int es_kind, double columns) {
  if (!c->opt_exact_small) return false;
  const double cost = es_kind == FGS_ES_WEIGHTED_KS ? c->sum_m2 * columns * (25.0 / 32.0)
                                                    : (double)c->nnz * columns * (10.0 / 32.0);
  return cost <= 2e10;
}

// rank + ES of the score columns currently in c->scores
static int rank_and_es(fgs_ctx *c, int es_kind, double p_exp, int abs_flag, long long Cb,
                       double *d_es, StageTimer &tm) {
  const int G = c->G;
  if (es_kind == FGS_ES_WEIGHTED_KS) {
    const size_t Gp = pitch_of(G);
    FGS_TRY(c->rank16.ensure(Gp * Cb));
    FGS_TRY(c->sabs.ensure(Gp * Cb));
    long long l0 = c->launches;
    if (c->cur_cert) {
      // ranks from tensor-core scores: adjacent sorted scores closer than the
      // contraction's error bound are recomputed exactly and their columns ranked again
      TieCert tc;
      tc.list = c->cert_list.p + c->opt_cert_cap; tc.count = c->cert_ctr.p + 1; tc.cap = c->opt_cert_cap;
      FGS_TRY(c->cert_redo.ensure((size_t)Cb + 8));
      FGS_CUDA(cudaMemsetAsync(c->cert_redo.p, 0, (size_t)Cb * sizeof(int), c->stream));
      tc.redo = c->cert_redo.p; tc.overflow = c->cur_ovf + 1;
      tc.rel_tol = 1e-10; tc.abs_tol = 1e-12;  // >= 50x the contraction-vs-reference error bound (k_s2n_mma.cu)
      tc.dup = c->has_dups ? c->dupclass.p : nullptr;
      FGS_TRY(launch_sort(c, c->scores.p, G, Cb, abs_flag, p_exp, c->rank16.p, c->sabs.p, nullptr, &tc));
      FGS_TRY(launch_s2n_exact_entries(c, c->x.p, c->n, G, c->cur_labels, c->n1n2.p, c->s2n_R, tc.list,
                                       tc.count, c->opt_cert_cap, c->scores.p, c->flags.p + c->cur_flag0,
                                       /*after_patch=*/true, c->cur_ovf + 1));
      // More near-tied pairs than the list holds (thousands of duplicated genes, heavily tied
      // data): THIS BLOCK's scores are recomputed with the bit-faithful kernel and every column
      // of it ranked again -- queued behind the device flag, no host round trip, and the other
      // blocks keep the tensor-core path.  (Not needed when the epilogue list already overflowed:
      // the block was recomputed before the first ranking.)
      FGS_TRY(launch_s2n(c, c->x.p, c->n, G, c->cur_labels, c->n1n2.p, c->s2n_R, Cb, c->scores.p,
                         c->flags.p + c->cur_flag0, c->cur_ovf + 1, c->cur_ovf));
      FGS_TRY(launch_s2n_patch(c, c->scores.p, G, c->s2n_R, c->cur_P, c->flags.p + c->cur_flag0));
      FGS_TRY(launch_mark_all_if(c, c->cert_redo.p, Cb, c->cur_ovf + 1, c->cur_ovf));
      FGS_TRY(launch_sort(c, c->scores.p, G, Cb, abs_flag, p_exp, c->rank16.p, c->sabs.p, c->cert_redo.p,
                          nullptr));
    } else {
      FGS_TRY(launch_sort(c, c->scores.p, G, Cb, abs_flag, p_exp, c->rank16.p, c->sabs.p));
    }
    c->last_launches[2] += c->launches - l0;
    tm.mark();
    l0 = c->launches;
    FGS_TRY(launch_es_wks(c, c->rank16.p, c->sabs.p, G, Cb, d_es, c->cur_cert));
    if (c->cur_cert) {
      // structural ties between es_pos and -es_neg hang on the last bit of the weights:
      // the tied units' member scores are recomputed in the reference's operation order
      // before the exact redo (list region 0 of the certification buffers is free again)
      int2 *list = c->cert_list.p;
      int *cnt = c->cert_ctr.p + 2;
      FGS_TRY(launch_tie_members(c, list, cnt, c->opt_cert_cap, c->flags.p + (c->flags.n - 1)));
      FGS_TRY(launch_s2n_exact_entries(c, c->x.p, c->n, G, c->cur_labels, c->n1n2.p, c->s2n_R, list, cnt,
                                       c->opt_cert_cap, c->scores.p, c->flags.p + c->cur_flag0,
                                       /*after_patch=*/true));
      FGS_TRY(launch_tie_refresh_weights(c, list, cnt, c->opt_cert_cap, c->scores.p, c->rank16.p, c->sabs.p,
                                         G, p_exp));
      FGS_TRY(launch_es_wks_ties(c, c->rank16.p, c->sabs.p, G, Cb, d_es));
    }
    c->last_launches[3] += c->launches - l0;
    tm.mark();
  } else if (es_kind == FGS_ES_MAXMEAN || es_kind == FGS_ES_MEAN) {
    tm.mark();
    long long l0 = c->launches;
    FGS_TRY(launch_es_mean(c, c->scores.p, G, Cb, es_kind, abs_flag, d_es));
    if (c->cur_cert && !c->cur_exact_all) {  // listed units: exact member scores first (tensor-core scores)
      int2 *list = c->cert_list.p;
      int *cnt = c->cert_ctr.p + 2;
      FGS_TRY(launch_tie_members(c, list, cnt, c->opt_cert_cap, c->flags.p + (c->flags.n - 1)));
      FGS_TRY(launch_s2n_exact_entries(c, c->x.p, c->n, G, c->cur_labels, c->n1n2.p, c->s2n_R, list, cnt,
                                       c->opt_cert_cap, c->scores.p, c->flags.p + c->cur_flag0,
                                       /*after_patch=*/true));
    }
    FGS_TRY(launch_es_mean_exact(c, c->scores.p, G, Cb, es_kind, abs_flag, d_es));
    c->last_launches[3] += c->launches - l0;
    tm.mark();
  } else {
    set_error("unknown es_kind %d", es_kind);
    return FGS_ERR_ARG;
  }
  c->last_cols = Cb;
  c->last_G = G;
  return FGS_OK;
}

}  // namespace fgs

using namespace fgs;

extern "C" {

int fgs_abi_version(void) { return FGS_ABI_VERSION; }
const char *fgs_last_error(void) { return g_err; }

int fgs_device_count(int *count) {
  if (!count) return FGS_ERR_ARG;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *count = 0;
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return FGS_ERR_CUDA;
  }
  *count = n;
  return FGS_OK;
}

int fgs_create(int device, fgs_ctx **out) {
  if (!out) return FGS_ERR_ARG;
  *out = nullptr;
  int n = 0;
  FGS_TRY(fgs_device_count(&n));
  if (device < 0 || device >= n) { set_error("device %d out of range (%d visible)", device, n); return FGS_ERR_ARG; }
  FGS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FGS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return FGS_ERR_UNSUPPORTED;
  }
  fgs_ctx *c = new fgs_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->smem_optin = (int)prop.sharedMemPerBlockOptin;
  FGS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  FGS_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) FGS_CUDA(cudaEventCreate(&c->ev[i]));
  for (int i = 0; i < 2; i++) {
    FGS_CUDA(cudaEventCreateWithFlags(&c->blk_copy[i], cudaEventDisableTiming));
    FGS_CUDA(cudaEventCreateWithFlags(&c->blk_done[i], cudaEventDisableTiming));
    FGS_CUDA(cudaEventCreateWithFlags(&c->blk_out[i], cudaEventDisableTiming));
  }
  *out = c;
  return FGS_OK;
}

int fgs_destroy(fgs_ctx *c) {
  if (!c) return FGS_OK;
  if (c->tile_job.joinable()) c->tile_job.join();
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  fgs_comm_destroy(c);
  c->set_groups.release(); c->x.release(); c->xc.release(); c->xtot.release(); c->cert_list.release(); c->cert_ctr.release();
  c->cert_redo.release(); c->sort_redo.release(); c->sort_glist.release(); c->gene_sd.release(); c->set_off.release(); c->set_idx.release(); c->set_idx16.release(); c->sig_lut.release(); c->set_slow.release(); c->tie_list.release(); c->tie_ctr.release(); c->obs_es.release(); c->ident.release(); c->obs_labels.release();
  for (auto &b : c->api_d) b.release();
  for (auto &b : c->api_i) b.release();
  c->api_l.release();
  c->set_info.release(); c->set_invgm.release(); c->slow_ids.release(); c->ybin.release();
  c->n1n2.release(); c->lm_H.release(); c->perm.release(); c->labels.release(); c->lm_B.release();
  c->scores.release(); c->scores2.release(); c->rank16.release(); c->sabs.release(); c->sort_keys.release();
  c->sort_idx.release(); c->flags.release(); c->counters.release(); c->es_null.release();
  c->sig_d.release(); c->sig_i.release();
  for (int i = 0; i < 2; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 2; i++) {
    if (c->blk_copy[i]) cudaEventDestroy(c->blk_copy[i]);
    if (c->blk_done[i]) cudaEventDestroy(c->blk_done[i]);
    if (c->blk_out[i]) cudaEventDestroy(c->blk_out[i]);
  }
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->stage) {
    cudaFreeHost(c->stage);
    for (int i = 0; i < 16; i++) if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
  }
  for (int i = 0; i < 8; i++) if (c->sig_ev[i]) cudaEventDestroy(c->sig_ev[i]);
  c->null_full.release(); c->class_lists.release(); c->blk_ovf.release(); c->perm_map.release(); c->es_uniq.release(); c->dupclass.release(); c->sig_t.release(); c->inv_n1n2.release();
  cudaStreamDestroy(c->stream);
  delete c;
  return FGS_OK;
}

int fgs_set_progress(fgs_ctx *c, fgs_progress_fn cb, void *user) {
  if (!c) return FGS_ERR_ARG;
  c->progress = cb;
  c->progress_user = user;
  return FGS_OK;
}

// Block-loop hook: with a progress callback set, wait until the block two back is finished
// (so at most two blocks are queued), report, and stop when asked to.  `k` = index of the block
// just enqueued; its blk_done event is recorded here.
static int block_progress(fgs_ctx *c, long long k, long long cols_per_block, long long total) {
  if (!c->progress) return FGS_OK;
  const int slot = (int)(k & 1);
  if (k >= 2) FGS_CUDA(cudaEventSynchronize(c->blk_done[slot]));
  FGS_CUDA(cudaEventRecord(c->blk_done[slot], c->stream));
  const long long done = std::min(total, std::max<long long>(0, k - 1) * cols_per_block);
  if (c->progress(c->progress_user, done, total) != 0) {
    cudaStreamSynchronize(c->stream);
    set_error("stopped by the progress callback after %lld of %lld score columns", done, total);
    return FGS_ERR_CANCELLED;
  }
  return FGS_OK;
}

int fgs_launch_count(fgs_ctx *c, long long *count) {
  if (!c || !count) return FGS_ERR_ARG;
  *count = c->launches;
  return FGS_OK;
}
int fgs_set_profiling(fgs_ctx *c, int enabled) {
  if (!c) return FGS_ERR_ARG;
  c->profiling = enabled != 0;
  return FGS_OK;
}
int fgs_last_timings(fgs_ctx *c, float *ms, long long *launches) {
  if (!c) return FGS_ERR_ARG;
  if (ms) for (int i = 0; i < 5; i++) ms[i] = c->last_ms[i];
  if (launches) for (int i = 0; i < 5; i++) launches[i] = c->last_launches[i];
  return FGS_OK;
}
int fgs_set_option(fgs_ctx *c, const char *name, long long value) {
  if (!c || !name) return FGS_ERR_ARG;
  if (!strcmp(name, "perm_block")) c->opt_perm_block = (int)value;
  else if (!strcmp(name, "h2d_threads")) c->opt_h2d_threads = (int)value;
  else if (!strcmp(name, "force_generic_es")) c->opt_force_generic = (int)value;
  else if (!strcmp(name, "force_sort64")) c->opt_sort64 = (int)value;
  else if (!strcmp(name, "s2n_mode")) c->opt_s2n_mode = (int)value;
  else if (!strcmp(name, "es_exact_small")) c->opt_exact_small = (int)value;
  else if (!strcmp(name, "tail_split")) c->opt_tail_split = (int)value;
  else if (!strcmp(name, "dedup_labels")) c->opt_dedup = (int)value;
  else if (!strcmp(name, "bulk_stage")) c->opt_bulk_stage = (int)value;
  else if (!strcmp(name, "arrange_members")) { c->opt_arrange = (int)value; c->invgm_G = -1; }
  else if (!strcmp(name, "nperm_total")) c->opt_nperm_total = value > 0 ? value : 0;
  else if (!strcmp(name, "cert_cap")) c->opt_cert_cap = value > 2 ? (int)value : 2;
  else { set_error("unknown option '%s'", name); return FGS_ERR_ARG; }
  return FGS_OK;
}

int fgs_get_stat(fgs_ctx *c, const char *name, long long *value) {
  if (!c || !name || !value) return FGS_ERR_ARG;
  if (!strcmp(name, "dedup_labellings")) *value = c->last_dedup;
  else if (!strcmp(name, "fallback_blocks")) *value = c->last_fallback_blocks;
  else if (!strcmp(name, "duplicated_rows")) *value = c->has_dups ? 1 : 0;
  else { set_error("unknown stat '%s'", name); return FGS_ERR_ARG; }
  return FGS_OK;
}

// ---------------------------------------------------------------------------
int fgs_set_x(fgs_ctx *c, const double *x, int n_sample, int n_gene) {
  if (!c || !x || n_sample <= 0 || n_gene <= 0) { set_error("fgs_set_x: bad argument"); return FGS_ERR_ARG; }
  FGS_TRY(use_device(c));
  FGS_TRY(c->x.ensure((size_t)n_sample * n_gene));
  FGS_TRY(h2d(c, c->x.p, x, (size_t)n_sample * n_gene));
  c->n = n_sample; c->G = n_gene;
  c->obs_valid = false;
  c->gene_sd_valid = false;
  c->xc_valid = false;
  // genes whose expression rows are bit-identical (duplicated probes): hashed on the device,
  // confirmed here on the caller's array; see TieCert::dup
  c->has_dups = false;
  {
    DevBuf<unsigned long long> &dh = c->sort_keys;   // scratch
    FGS_TRY(dh.ensure((size_t)n_gene));
    FGS_TRY(launch_row_hash(c, c->x.p, n_sample, n_gene, dh.p));
    std::vector<unsigned long long> h((size_t)n_gene);
    FGS_CUDA(cudaMemcpyAsync(h.data(), dh.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    FGS_TRY(sync(c));
    std::vector<int> order((size_t)n_gene), dup((size_t)n_gene);
    for (int g = 0; g < n_gene; g++) { order[g] = g; dup[g] = g; }
    std::sort(order.begin(), order.end(), [&](int a, int b) { return h[a] != h[b] ? h[a] < h[b] : a < b; });
    const size_t row = (size_t)n_sample * sizeof(double);
    for (int i = 0; i < n_gene;) {
      int j = i + 1;
      while (j < n_gene && h[order[j]] == h[order[i]]) j++;
      for (int a = i; a < j; a++) {            // (groups are almost always of size one)
        if (dup[order[a]] != order[a]) continue;
        for (int b = a + 1; b < j; b++)
          if (dup[order[b]] == order[b] &&
              !memcmp(x + (size_t)order[a] * n_sample, x + (size_t)order[b] * n_sample, row)) {
            dup[order[b]] = order[a];
            c->has_dups = true;
          }
      }
      i = j;
    }
    if (c->has_dups) {
      FGS_TRY(c->dupclass.ensure((size_t)n_gene));
      FGS_TRY(h2d(c, c->dupclass.p, dup.data(), (size_t)n_gene));
      FGS_TRY(sync(c));
    }
  }
  return FGS_OK;
}

int fgs_set_gene_sets(fgs_ctx *c, const int *offsets, const int *index, int n_sets) {
  if (!c || !offsets || n_sets <= 0) { set_error("fgs_set_gene_sets: bad argument"); return FGS_ERR_ARG; }
  FGS_TRY(use_device(c));
  if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return FGS_ERR_ARG; }
  for (int s = 0; s < n_sets; s++)
    if (offsets[s + 1] < offsets[s]) { set_error("offsets not monotone at set %d", s); return FGS_ERR_ARG; }
  const long long nnz = offsets[n_sets];
  if (nnz > 0 && !index) { set_error("fgs_set_gene_sets: index is NULL"); return FGS_ERR_ARG; }
  // The member lists are checked and made 0-based on the device (launch_prep_sets):
  // the host only orders the sets by size.
  FGS_TRY(c->set_off.ensure((size_t)n_sets + 1));
  FGS_TRY(c->set_idx.ensure((size_t)std::max<long long>(nnz, 1)));
  FGS_TRY(c->set_info.ensure(n_sets));
  FGS_TRY(c->set_slow.ensure((size_t)n_sets + 16));
  FGS_TRY(h2d(c, c->set_off.p, offsets, (size_t)n_sets + 1));
  FGS_TRY(h2d(c, c->set_idx.p, index, (size_t)nnz));  // raw 1-based ids, rewritten in place
  int status[2];
  std::vector<unsigned char> slow(n_sets);
  FGS_TRY(launch_prep_sets(c, n_sets, status, slow.data()));  // synchronises
  if (status[0] < n_sets) {
    const int s = status[0];
    int bad = 0;
    for (int j = offsets[s]; j < offsets[s + 1]; j++) if (index[j] < 1) { bad = index[j]; break; }
    set_error("gene index %d < 1 in set %d (NA from match()?)", bad, s);
    c->S = 0;
    return FGS_ERR_ARG;
  }
  std::vector<int> order(n_sets), slow_ids;
  for (int s = 0; s < n_sets; s++) { order[s] = s; if (slow[s]) slow_ids.push_back(s); }
  std::stable_sort(order.begin(), order.end(), [&](int p, int q) {
    return (offsets[p + 1] - offsets[p]) > (offsets[q + 1] - offsets[q]);
  });
  std::vector<int4> info(n_sets);
  long long blocks = 0;  // members as uint16 in 16-byte blocks, processing order (launch_es_wks fills them)
  for (int k = 0; k < n_sets; k++) {
    const int sid = order[k], m = offsets[sid + 1] - offsets[sid];
    info[k] = make_int4(offsets[sid], slow[sid] ? (m | (int)0x80000000) : m, (int)blocks, sid);
    blocks += (m + 7) / 8;
  }
  c->idx16_blocks = blocks;
  // Work list of the weighted-KS kernel.  A set of m members is sorted with SL = 4 .. 64 members per
  // lane (the smallest with 8 * SL >= m) on L = ceil(m / SL) lanes; sets of the same (SL, L) sit back
  // to back in one warp, floor(31 / L) of them (one lane stays idle as the all-pad partner) or four
  // of eight lanes.  Sets above 512 members take a warp each.
  std::vector<int2> groups;
  {
    int cur_sl = -1, cur_L = -1, cur_cap = 0;
    for (int k = 0; k < n_sets; k++) {
      const int m = info[k].y & 0x7fffffff;
      int sl, L, cap;
      if (m > 512) { sl = m <= 1024 ? 5 : 6; L = 32; cap = 1; }
      else {
        sl = 2;
        while ((8 << sl) < m) sl++;
        L = std::max(3, (m + (1 << sl) - 1) >> sl);
        if (L >= 7) L = 8;            // seven lanes also make four sets per warp
        cap = L == 8 ? 4 : 31 / L;
      }
      if (groups.empty() || sl != cur_sl || L != cur_L || (groups.back().y & 0xff) >= cur_cap) {
        groups.push_back(make_int2(k, (L << 8) | (sl << 16)));
        cur_sl = sl; cur_L = L; cur_cap = cap;
      }
      groups.back().y++;
    }
  }
  c->n_groups = (int)groups.size();
  c->n_wide_groups = 0;
  while (c->n_wide_groups < c->n_groups && ((groups[c->n_wide_groups].y >> 8) & 0xff) == 32) c->n_wide_groups++;
  FGS_TRY(c->set_groups.ensure(std::max<size_t>(groups.size(), 1)));
  FGS_TRY(h2d(c, c->set_groups.p, groups.data(), groups.size()));
  c->sum_m2 = 0.0;
  for (int s2 = 0; s2 < n_sets; s2++) { const double m = offsets[s2 + 1] - offsets[s2]; c->sum_m2 += m * m; }
  FGS_TRY(c->slow_ids.ensure(std::max<size_t>(slow_ids.size(), 1)));
  FGS_TRY(h2d(c, c->set_info.p, info.data(), (size_t)n_sets));
  FGS_TRY(h2d(c, c->slow_ids.p, slow_ids.data(), slow_ids.size()));
  FGS_TRY(sync(c));
  c->S = n_sets; c->nnz = nnz; c->max_set = status[1]; c->n_slow = (int)slow_ids.size();
  c->invgm_G = -1;
  c->obs_valid = false;
  c->h_set_off.assign(offsets, offsets + n_sets + 1);
  return FGS_OK;
}

int fgs_set_response_s2n(fgs_ctx *c, const int *y_bin, int n_sample, int n_response) {
  if (!c || !y_bin || n_sample <= 0 || n_response <= 0) { set_error("fgs_set_response_s2n: bad argument"); return FGS_ERR_ARG; }
  FGS_TRY(use_device(c));
  std::vector<int> nn(2 * (size_t)n_response, 0);
  c->s2n_has_na_pending = false;
  for (int r = 0; r < n_response; r++)
    for (int i = 0; i < n_sample; i++) {
      int v = y_bin[(size_t)r * n_sample + i];
      if (v == INT_MIN) { c->s2n_has_na_pending = true; continue; }
      nn[2 * r + (v ? 0 : 1)]++;
    }
  FGS_TRY(c->ybin.ensure((size_t)n_sample * n_response));
  FGS_TRY(c->n1n2.ensure(2 * (size_t)n_response));
  FGS_TRY(h2d(c, c->ybin.p, y_bin, (size_t)n_sample * n_response));
  FGS_TRY(h2d(c, c->n1n2.p, nn.data(), nn.size()));
  std::vector<double> inv(nn.size());
  for (size_t i = 0; i < nn.size(); i++) inv[i] = 1.0 / (double)nn[i];
  FGS_TRY(c->inv_n1n2.ensure(inv.size()));
  FGS_TRY(h2d(c, c->inv_n1n2.p, inv.data(), inv.size()));
  FGS_TRY(sync(c));
  c->h_n1n2 = nn;
  c->s2n_n = n_sample;
  c->s2n_R = n_response;
  c->obs_valid = false;
  c->s2n_has_na = c->s2n_has_na_pending;
  if (c->n && c->n != n_sample) { /* checked again at run time */ }
  return FGS_OK;
}

int fgs_set_response_lm(fgs_ctx *c, const double *design, int n_sample, int n_col, int has_intercept) {
  if (!c || !design || n_sample <= 0 || n_col <= 0) { set_error("fgs_set_response_lm: bad argument"); return FGS_ERR_ARG; }
  FGS_TRY(use_device(c));
  const int R = n_col - (has_intercept ? 1 : 0);
  if (R <= 0) { set_error("design has no non-intercept column"); return FGS_ERR_ARG; }
  std::vector<double> D(design, design + (size_t)n_sample * n_col), H;
  if (!design_pinv(D, n_sample, n_col, H)) {
    set_error("Gene scoring function returned NA or Inf. (rank-deficient design)");
    return FGS_ERR_NONFINITE;
  }
  FGS_TRY(c->lm_H.ensure((size_t)R * n_sample));
  FGS_TRY(h2d(c, c->lm_H.p, H.data() + (has_intercept ? n_sample : 0), (size_t)R * n_sample));
  FGS_TRY(sync(c));
  c->lm_q = n_col; c->lm_R = R; c->lm_icpt = has_intercept ? 1 : 0;
  c->obs_valid = false;
  c->h_design.swap(D);
  return FGS_OK;
}

int fgs_n_response(fgs_ctx *c, int score_kind, int *n_response) {
  if (!c || !n_response) return FGS_ERR_ARG;
  *n_response = n_response_of(c, score_kind);
  return FGS_OK;
}

// ---------------------------------------------------------------------------
// gene scores of `P` permutations starting at device perm column p0 (d_perm may
// be NULL = identity) into c->scores [P*R][G]; flags at c->flags + flag0
static int score_block(fgs_ctx *c, int score_kind, const int *d_perm, const uint32_t *d_labels,
                       int P, long long flag0, bool force_exact = false) {
  const int n = c->n, G = c->G;
  const int R = n_response_of(c, score_kind);
  const long long Cb = (long long)P * R;
  FGS_TRY(c->scores.ensure((size_t)G * Cb));
  c->cur_cert = false;
  if (score_kind == FGS_SCORE_S2N) {
    const bool tensor = c->opt_s2n_mode == 1 && !c->s2n_has_na && !force_exact;
    if (tensor) {
      // K2: FP64 tensor-core contraction; entries whose variance cancelled (or are
      // not finite) are recomputed with the bit-faithful arithmetic right away
      if (!c->xc_valid) {
        FGS_TRY(c->xc.ensure((size_t)n * G));
        FGS_TRY(c->xtot.ensure((size_t)2 * G));
        FGS_TRY(launch_center(c, c->x.p, n, G, c->xc.p, c->xtot.p));
        c->xc_valid = true;
      }
      FGS_TRY(c->cert_list.ensure((size_t)2 * c->opt_cert_cap));
      FGS_TRY(c->cert_ctr.ensure(4));
      FGS_CUDA(cudaMemsetAsync(c->cert_ctr.p, 0, 4 * sizeof(int), c->stream));
      FGS_TRY(launch_s2n_mma(c, c->xc.p, c->xtot.p, n, G, d_labels, c->inv_n1n2.p, R, Cb, c->scores.p,
                             c->cert_list.p, c->cert_ctr.p, c->opt_cert_cap, c->cur_ovf));
      FGS_TRY(launch_s2n_exact_entries(c, c->x.p, n, G, d_labels, c->n1n2.p, R, c->cert_list.p,
                                       c->cert_ctr.p, c->opt_cert_cap, c->scores.p, c->flags.p + flag0, false,
                                       c->cur_ovf));
      // epilogue list full (most entries cancelled: near-constant genes): the block is rescored bit-faithfully
      FGS_TRY(launch_s2n(c, c->x.p, n, G, d_labels, c->n1n2.p, R, Cb, c->scores.p, c->flags.p + flag0, c->cur_ovf,
                         nullptr));
      c->cur_cert = true;
      c->cur_labels = d_labels;
      c->cur_flag0 = flag0;
      c->cur_P = P;
    } else {
      FGS_TRY(launch_s2n(c, c->x.p, n, G, d_labels, c->n1n2.p, R, Cb, c->scores.p, c->flags.p + flag0));
    }
    FGS_TRY(launch_s2n_patch(c, c->scores.p, G, R, P, c->flags.p + flag0));
  } else {
    if (!c->gene_sd_valid) {
      FGS_TRY(c->gene_sd.ensure(G));
      FGS_TRY(launch_gene_sd(c, c->x.p, n, G, c->gene_sd.p));
      c->gene_sd_valid = true;
    }
    FGS_TRY(c->lm_B.ensure((size_t)n * Cb));
    FGS_TRY(launch_lm_build_B(c, c->lm_H.p, d_perm, n, R, P, c->lm_B.p));
    FGS_TRY(launch_lm_gemm(c, c->x.p, n, G, c->lm_B.p, Cb, c->gene_sd.p, c->scores.p,
                           c->flags.p + flag0, R));
  }
  return FGS_OK;
}

static int check_response(fgs_ctx *c, int score_kind) {
  if (score_kind == FGS_SCORE_S2N) {
    if (c->s2n_R <= 0) { set_error("fgs_set_response_s2n has not been called"); return FGS_ERR_STATE; }
  } else if (score_kind == FGS_SCORE_LM) {
    if (c->lm_R <= 0) { set_error("fgs_set_response_lm has not been called"); return FGS_ERR_STATE; }
  } else { set_error("unknown score_kind %d", score_kind); return FGS_ERR_ARG; }
  return FGS_OK;
}

int fgs_s2n(fgs_ctx *c, const double *x, int n_sample_x, int n_gene, const int *y, int n_sample_y,
            int n_response, double *coef) {
  if (!c || !x || !y || !coef || n_gene <= 0 || n_response <= 0) { set_error("fgs_s2n: bad argument"); return FGS_ERR_ARG; }
  if (n_sample_x != n_sample_y) {  // src/ggsea.cpp:15-17
    std::memset(coef, 0, sizeof(double) * (size_t)n_gene * n_response);
    return FGS_OK;
  }
  FGS_TRY(use_device(c));
  // independent of the resident problem: temporary buffers
  DevBuf<double> dx, dsc; DevBuf<int> dy, dnn, dfl; DevBuf<uint32_t> dlab;
  const int n = n_sample_x, W = (n + 31) / 32;
  int rc = FGS_OK;
  std::vector<int> nn(2 * (size_t)n_response, 0);
  for (int r = 0; r < n_response; r++)
    for (int i = 0; i < n; i++) { int v = y[(size_t)r * n + i]; if (v != INT_MIN) nn[2 * r + (v ? 0 : 1)]++; }
  auto body = [&]() -> int {
    FGS_TRY(dx.ensure((size_t)n * n_gene)); FGS_TRY(dsc.ensure((size_t)n_gene * n_response));
    FGS_TRY(dy.ensure((size_t)n * n_response)); FGS_TRY(dnn.ensure(nn.size()));
    FGS_TRY(dfl.ensure(2)); FGS_TRY(dlab.ensure((size_t)n_response * 2 * W));
    FGS_TRY(h2d(c, dx.p, x, (size_t)n * n_gene)); FGS_TRY(h2d(c, dy.p, y, (size_t)n * n_response));
    FGS_TRY(h2d(c, dnn.p, nn.data(), nn.size()));
    FGS_CUDA(cudaMemsetAsync(dfl.p, 0, 2 * sizeof(int), c->stream));
    FGS_TRY(flags_reset(c, 1));
    FGS_TRY(launch_labels(c, nullptr, dy.p, n, n_response, 1, dlab.p));
    FGS_TRY(launch_s2n(c, dx.p, n, n_gene, dlab.p, dnn.p, n_response, n_response, dsc.p, dfl.p));
    FGS_TRY(d2h(c, coef, dsc.p, (size_t)n_gene * n_response));
    FGS_TRY(sync(c));
    return FGS_OK;
  };
  rc = body();
  dx.release(); dsc.release(); dy.release(); dnn.release(); dfl.release(); dlab.release();
  return rc;
}

int fgs_score_observed(fgs_ctx *c, int score_kind, int abs_flag, double *scores) {
  if (!c || !scores) return FGS_ERR_ARG;
  FGS_TRY(use_device(c));
  FGS_TRY(check_problem(c, false));
  FGS_TRY(check_response(c, score_kind));
  const int R = n_response_of(c, score_kind), n = c->n, G = c->G;
  FGS_TRY(flags_reset(c, 1));
  if (score_kind == FGS_SCORE_S2N) {
    const int W = (n + 31) / 32;
    FGS_TRY(c->labels.ensure((size_t)R * 2 * W));
    FGS_TRY(launch_labels(c, nullptr, c->ybin.p, n, R, 1, c->labels.p));
  }
  FGS_TRY(score_block(c, score_kind, nullptr, c->labels.p, 1, 0, /*force_exact=*/true));
  if (abs_flag) FGS_TRY(launch_abs_inplace(c, c->scores.p, (long long)G * R));
  FGS_TRY(d2h(c, scores, c->scores.p, (size_t)G * R));
  FGS_TRY(flags_check(c, 1));
  c->pend_score_kind = score_kind; c->pend_abs = abs_flag ? 1 : 0;
  return FGS_OK;
}

int fgs_perm_null(fgs_ctx *c, int score_kind, int es_kind, double p_exp, int abs_flag,
                  const int *perm, unsigned long long seed, long long perm_id0, int n_perm,
                  double *es_null_out) {
  if (!c || n_perm < 0) { set_error("fgs_perm_null: bad argument"); return FGS_ERR_ARG; }
  FGS_TRY(use_device(c));
  FGS_TRY(check_problem(c, true));
  FGS_TRY(check_response(c, score_kind));
  const int n = c->n, G = c->G, S = c->S, P = n_perm;
  const int R = n_response_of(c, score_kind);
  for (int i = 0; i < 5; i++) c->last_launches[i] = 0;
  FGS_TRY(c->es_null.ensure(std::max<size_t>((size_t)S * R * P, 1)));
  c->null_S = S; c->null_R = R; c->null_P = P;
  // the observed ES of this problem, if fgs_observed_es made one: null units that come out
  // within rounding of it are redone in the reference's arithmetic (see es_wks_tie_kernel)
  c->cur_obs = (c->obs_valid && c->obs_kind == es_kind && c->obs_S == S && c->obs_R == R) ? c->obs_es.p : nullptr;
  c->cur_obs_R = R;
  c->null_exact_near_obs = c->cur_obs != nullptr;
  // (routing by the TOTAL permutation count when the caller shards: the same units take the same
  // path on one GPU and on eight)
  const long long P_total = std::max<long long>(P, c->opt_nperm_total);
  c->cur_exact_all = score_kind == FGS_SCORE_S2N && small_problem(c, es_kind, (long long)R * P_total);
  // tiny label space (see distinct_labellings): decided from the class sizes and the TOTAL
  // permutation count, so every shard of a sharded null takes the same route
  const double n_lab = (score_kind == FGS_SCORE_S2N && c->opt_dedup) ? distinct_labellings(c) : INFINITY;
  const bool try_dedup = n_lab * 2.0 <= (double)P_total && P >= 2;
  if (try_dedup && small_problem_dedup(c, es_kind, n_lab * R)) c->cur_exact_all = true;
  if (c->cur_exact_all) c->null_exact_near_obs = true;
  if (P == 0) return FGS_OK;
  FGS_TRY(flags_reset(c, P));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  FGS_CUDA(cudaEventRecord(c->ev[0], c->stream));  // always-on device time of the whole null loop
  if (c->profiling) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, c->stream); }
  long long l0 = c->launches;
  // K1: permutations (uploaded or drawn) and, for s2n, permuted class masks
  FGS_TRY(c->perm.ensure((size_t)n * P));
  if (perm) FGS_TRY(h2d(c, c->perm.p, perm, (size_t)n * P));
  else FGS_TRY(launch_philox_perm(c, seed, perm_id0, n, P, c->perm.p));
  const int W = (n + 31) / 32;
  c->cur_ident = nullptr;
  // Pe "effective permutations" computed below: all of them, or the distinct labellings only
  int Pe = P;
  double *es_e = c->es_null.p;
  bool dedup = false;
  if (score_kind == FGS_SCORE_S2N) {
    FGS_TRY(c->labels.ensure((size_t)P * R * 2 * W));
    FGS_TRY(launch_labels(c, c->perm.p, c->ybin.p, n, R, P, c->labels.p));
    if (try_dedup && W == 1) {
      // the label masks are tiny (2 R words per permutation): group them on the host
      const size_t key_words = (size_t)R * 2;
      std::vector<uint32_t> h((size_t)P * key_words), hu;
      FGS_CUDA(cudaMemcpyAsync(h.data(), c->labels.p, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
      FGS_TRY(sync(c));
      std::unordered_map<std::string, int> seen;
      std::vector<int> map(P);
      for (int p = 0; p < P; p++) {
        std::string key((const char *)&h[(size_t)p * key_words], key_words * sizeof(uint32_t));
        auto it = seen.find(key);
        if (it == seen.end()) {
          it = seen.emplace(std::move(key), (int)seen.size()).first;
          hu.insert(hu.end(), h.begin() + (size_t)p * key_words, h.begin() + (size_t)(p + 1) * key_words);
        }
        map[p] = it->second;
      }
      const int U = (int)seen.size();
      if ((long long)U * 2 <= P) {
        dedup = true;
        Pe = U;
        FGS_TRY(c->perm_map.ensure((size_t)P));
        FGS_TRY(h2d(c, c->perm_map.p, map.data(), (size_t)P));
        FGS_TRY(h2d(c, c->labels.p, hu.data(), hu.size()));   // the distinct labellings, in order of first use
        FGS_TRY(c->es_uniq.ensure((size_t)S * R * U));
        es_e = c->es_uniq.p;
        FGS_TRY(sync(c));  // (map / hu are host temporaries)
      }
    }
    if (c->cur_obs) {  // label-preserving permutations get the observed ES as it is
      FGS_TRY(c->obs_labels.ensure((size_t)R * 2 * W));
      FGS_TRY(c->ident.ensure((size_t)Pe));
      FGS_TRY(launch_labels(c, nullptr, c->ybin.p, n, R, 1, c->obs_labels.p));
      FGS_TRY(launch_label_identity(c, c->labels.p, c->obs_labels.p, R, W, Pe, c->ident.p));
    }
  }
  c->last_dedup = dedup ? Pe : 0;
  c->last_launches[0] = c->launches - l0;
  float pro_ms = 0;
  if (c->profiling) { cudaEventRecord(e1, c->stream); }
  StageTimer tm(c);
  const long long Cb_max = block_columns(c, (long long)Pe * R);
  const int Pb_max = (int)std::max<long long>(1, Cb_max / R);
  const int n_blk = (Pe + Pb_max - 1) / Pb_max;
  FGS_TRY(c->blk_ovf.ensure((size_t)2 * n_blk + 2));
  FGS_CUDA(cudaMemsetAsync(c->blk_ovf.p, 0, ((size_t)2 * n_blk + 2) * sizeof(int), c->stream));
  for (int p0 = 0; p0 < Pe; p0 += Pb_max) {
    const int Pb = std::min(Pb_max, Pe - p0);
    c->cur_ovf = c->blk_ovf.p + 2 * (size_t)(p0 / Pb_max);
    tm.mark();
    l0 = c->launches;
    FGS_TRY(score_block(c, score_kind, c->perm.p + (size_t)p0 * n,
                        c->labels.p + (size_t)p0 * R * 2 * W, Pb, p0, /*force_exact=*/c->cur_exact_all));
    c->last_launches[1] += c->launches - l0;
    tm.mark();
    // (the observed ES is copied to label-preserving permutations only when it is known to be the
    // ES of THESE scores: made by fgs_score_observed with this score kind and abs flag, same exponent)
    c->cur_ident = (c->cur_obs && score_kind == FGS_SCORE_S2N && c->obs_score_kind == score_kind &&
                    c->obs_abs == (abs_flag ? 1 : 0) && (es_kind != FGS_ES_WEIGHTED_KS || c->obs_p == p_exp))
                       ? c->ident.p + p0 : nullptr;
    FGS_TRY(rank_and_es(c, es_kind, p_exp, abs_flag, (long long)Pb * R, es_e + (size_t)p0 * R * S, tm));
    if (c->cur_ident)
      FGS_TRY(launch_copy_observed(c, c->cur_obs, c->cur_ident, R, (long long)Pb * R, es_e + (size_t)p0 * R * S));
    if (c->progress) {
      const int prc = block_progress(c, p0 / Pb_max, (long long)Pb_max * R, (long long)Pe * R);
      if (prc != FGS_OK) { c->cur_ident = nullptr; c->cur_exact_all = false; return prc; }
    }
    if (es_null_out && !dedup) {
      // es.null goes back block by block: block k-1 leaves (pinned staging ring, several host
      // threads) while block k, already queued, is computed
      const int k = p0 / Pb_max;
      FGS_CUDA(cudaEventRecord(c->blk_out[k & 1], c->stream));
      if (k > 0) {
        const size_t o = (size_t)(p0 - Pb_max) * R * S;
        FGS_TRY(download_block(c, (k - 1) & 1, es_null_out + o, c->es_null.p + o, (size_t)Pb_max * R * S));
      }
      if (p0 + Pb >= P) {
        const size_t o = (size_t)p0 * R * S;
        FGS_TRY(download_block(c, k & 1, es_null_out + o, c->es_null.p + o, (size_t)Pb * R * S));
      }
    }
  }
  c->cur_ident = nullptr;
  c->cur_exact_all = false;
  if (dedup) {  // every permutation takes the column of its labelling
    FGS_TRY(launch_scatter_null(c, c->es_uniq.p, c->perm_map.p, (long long)S * R, P, c->es_null.p));
    if (es_null_out) {
      FGS_CUDA(cudaEventRecord(c->blk_out[0], c->stream));
      FGS_TRY(download_block(c, 0, es_null_out, c->es_null.p, (size_t)S * R * P));
    }
  }
  FGS_CUDA(cudaEventRecord(c->ev[1], c->stream));
  std::vector<int> h_ovf((size_t)2 * n_blk);
  FGS_CUDA(cudaMemcpyAsync(h_ovf.data(), c->blk_ovf.p, h_ovf.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  int rc = flags_check(c, Pe);  // (synchronises)
  c->last_fallback_blocks = 0;
  for (int k = 0; k < n_blk; k++) c->last_fallback_blocks += (h_ovf[2 * k] || h_ovf[2 * k + 1]) ? 1 : 0;
  if (c->profiling) {
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&pro_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  tm.finish(pro_ms);
  if (cudaEventSynchronize(c->ev[1]) == cudaSuccess) cudaEventElapsedTime(&c->last_ms[4], c->ev[0], c->ev[1]);
  c->last_launches[4] = c->last_launches[0] + c->last_launches[1] + c->last_launches[2] + c->last_launches[3];
  if (rc == CERT_OVERFLOW) {
    // more near-tied scores than the certification lists hold (e.g. thousands of
    // duplicated genes): run this call with the bit-faithful score kernel instead
    const int saved = c->opt_s2n_mode;
    c->opt_s2n_mode = 0;
    rc = fgs_perm_null(c, score_kind, es_kind, p_exp, abs_flag, perm, seed, perm_id0, n_perm, es_null_out);
    c->opt_s2n_mode = saved;
    return rc;
  }
  if (rc != FGS_OK) return rc;
  return FGS_OK;   // es_null_out was filled block by block inside the loop
}

// scores of permutations [perm_off, perm_off + n_perm) of a null of P_total permutations; `async`:
// queue the work and return (fgs_null_finish waits and reports)
static int es_from_scores_impl(fgs_ctx *c, int es_kind, double p_exp, const double *scores, int n_gene,
                               int n_response, int n_perm, long long perm_off, long long P_total, double *es_out,
                               bool async) {
  if (!c || !scores || n_gene <= 0 || n_response <= 0 || n_perm < 0 || perm_off < 0 || perm_off + n_perm > P_total) {
    set_error("fgs_es_from_scores: bad argument");
    return FGS_ERR_ARG;
  }
  FGS_CUDA(cudaSetDevice(c->device));   // (not use_device: this may BE the tile in flight)
  if (c->S <= 0) { set_error("fgs_set_gene_sets has not been called"); return FGS_ERR_STATE; }
  if (c->G != 0 && c->G != n_gene && c->x.p) { /* x not needed on this path; sets decide */ }
  if (c->max_set > n_gene) { set_error("gene set index %d exceeds %d genes", c->max_set, n_gene); return FGS_ERR_ARG; }
  const int saveG = c->G;
  c->G = n_gene;
  c->cur_cert = false;  // user scores: nothing to certify
  const int S = c->S, R = n_response, P = n_perm;
  const long long C = (long long)R * P, C_total = (long long)R * P_total, col0 = (long long)R * perm_off;
  for (int i = 0; i < 5; i++) c->last_launches[i] = 0;
  int rc = FGS_OK;
  auto body = [&]() -> int {
    if (perm_off == 0) {
      FGS_TRY(c->es_null.ensure(std::max<size_t>((size_t)S * C_total, 1)));
    } else if (c->es_null.n < (size_t)S * C_total || c->null_S != S || c->null_R != R || c->null_P != P_total) {
      set_error("fgs_es_from_scores_tile: tiles must start at permutation 0 and keep the gene sets, the number of "
                "responses and the total");
      return FGS_ERR_STATE;
    }
    c->null_S = S; c->null_R = R; c->null_P = (int)P_total;
    double *const d_null = c->es_null.p + (size_t)col0 * S;   // this call's columns of the resident null
    c->cur_obs = (c->obs_valid && c->obs_kind == es_kind && c->obs_S == S && c->obs_R == R) ? c->obs_es.p : nullptr;
    c->cur_obs_R = R;
    c->cur_ident = nullptr;
    c->null_exact_near_obs = c->cur_obs != nullptr;
    c->cur_exact_all = small_problem(c, es_kind, std::max<long long>(C_total, (long long)R * c->opt_nperm_total));
    if (c->cur_exact_all) c->null_exact_near_obs = true;
    if (C == 0) return FGS_OK;
    if (perm_off == 0) FGS_TRY(flags_reset(c, C_total));
    int *const d_flags = c->flags.p + col0;
    FGS_CUDA(cudaEventRecord(c->ev[0], c->stream));
    StageTimer tm(c);
    // Blocks of score columns stream in: block k+1 is uploaded on copy_stream into the second
    // score buffer while block k is checked, ranked and scored on the compute stream (the step a
    // user-defined gene.score.fn feeds, R/flexgsea.R:393-394).  The compute work
    // of a block is enqueued BEFORE the next upload is issued: an upload from pageable host memory
    // keeps the calling thread busy, and the device must already have its work by then.
    const long long Cb_max = block_columns(c, C);
    const long long n_blk = (C + Cb_max - 1) / Cb_max;
    const size_t blk_elems = (size_t)n_gene * std::min(Cb_max, C);
    FGS_TRY(c->scores.ensure(blk_elems));
    if (n_blk > 1) FGS_TRY(c->scores2.ensure(blk_elems));
    auto upload = [&](long long k, double *dst) -> int {
      const long long c0 = k * Cb_max, Cb = std::min(Cb_max, C - c0);
      const int slot = (int)(k & 1);
      if (k >= 2) FGS_CUDA(cudaStreamWaitEvent(c->copy_stream, c->blk_done[slot], 0));
      FGS_TRY(staged_upload(c, dst, scores + (size_t)c0 * n_gene, (size_t)n_gene * Cb * sizeof(double),
                            c->copy_stream));
      FGS_CUDA(cudaEventRecord(c->blk_copy[slot], c->copy_stream));
      return FGS_OK;
    };
    // nothing queued earlier on the compute stream may still read the score buffers
    FGS_CUDA(cudaEventRecord(c->blk_done[0], c->stream));
    FGS_CUDA(cudaStreamWaitEvent(c->copy_stream, c->blk_done[0], 0));
    FGS_TRY(upload(0, c->scores.p));
    int brc = FGS_OK;
    for (long long k = 0; k < n_blk && brc == FGS_OK; k++) {
      const long long c0 = k * Cb_max, Cb = std::min(Cb_max, C - c0);
      const int slot = (int)(k & 1);
      tm.mark();
      FGS_CUDA(cudaStreamWaitEvent(c->stream, c->blk_copy[slot], 0));
      FGS_TRY(launch_check_finite(c, c->scores.p, (long long)n_gene * Cb, n_gene, d_flags + c0));
      tm.mark();
      FGS_TRY(rank_and_es(c, es_kind, p_exp, 0, Cb, d_null + (size_t)c0 * S, tm));
      // (no progress hook from a tile's helper thread: an R binding's hook must run on R's own thread)
      const bool hook = c->progress && !async;
      if (hook && k >= 2) FGS_CUDA(cudaEventSynchronize(c->blk_done[slot]));
      FGS_CUDA(cudaEventRecord(c->blk_done[slot], c->stream));
      if (es_out) FGS_CUDA(cudaEventRecord(c->blk_out[slot], c->stream));
      if (k + 1 < n_blk) FGS_TRY(upload(k + 1, c->scores2.p));
      if (es_out && k > 0) {  // results of block k-1 go back while block k is computed
        const size_t o = (size_t)(c0 - Cb_max) * S;
        FGS_TRY(download_block(c, slot ^ 1, es_out + o, d_null + o, (size_t)Cb_max * S));
      }
      if (es_out && k + 1 == n_blk)
        FGS_TRY(download_block(c, slot, es_out + (size_t)c0 * S, d_null + (size_t)c0 * S, (size_t)Cb * S));
      if (hook && c->progress(c->progress_user, std::min(C, std::max<long long>(0, k - 1) * Cb_max), C) != 0) {
        set_error("stopped by the progress callback");
        brc = FGS_ERR_CANCELLED;
      }
      if (k + 1 < n_blk) std::swap(c->scores, c->scores2);
    }
    c->cur_exact_all = false;
    if (brc != FGS_OK) { cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(c->stream); return brc; }
    FGS_CUDA(cudaEventRecord(c->ev[1], c->stream));
    if (async) { tm.drop(); return FGS_OK; }   // fgs_null_finish waits and checks the flags
    int r2 = flags_check(c, C_total);
    tm.finish(0.f);
    if (cudaEventSynchronize(c->ev[1]) == cudaSuccess) cudaEventElapsedTime(&c->last_ms[4], c->ev[0], c->ev[1]);
    if (r2 != FGS_OK) return r2;
    return FGS_OK;   // es_out was filled block by block inside the loop
  };
  rc = body();
  c->G = saveG;
  return rc;
}

int fgs_es_from_scores(fgs_ctx *c, int es_kind, double p_exp, const double *scores, int n_gene,
                       int n_response, int n_perm, double *es_out) {
  if (!c) return FGS_ERR_ARG;
  FGS_TRY(settle_tile(c));
  return es_from_scores_impl(c, es_kind, p_exp, scores, n_gene, n_response, n_perm, 0, n_perm, es_out, false);
}

int fgs_es_from_scores_tile(fgs_ctx *c, int es_kind, double p_exp, const double *scores, int n_gene,
                            int n_response, int n_perm_tile, long long perm_offset, long long n_perm_total) {
  if (!c) return FGS_ERR_ARG;
  if (n_perm_total > INT_MAX) { set_error("too many permutations"); return FGS_ERR_ARG; }
  // the previous tile must be through (its buffer is the caller's again from here on) ...
  FGS_TRY(settle_tile(c));
  // ... and this one is uploaded, ranked and scored by a helper thread while the caller scores the next:
  // even the staging copy out of the caller's pageable array is off the caller's critical path
  c->tile_rc = FGS_OK;
  c->tile_job = std::thread([=]() {
    const int rc = es_from_scores_impl(c, es_kind, p_exp, scores, n_gene, n_response, n_perm_tile, perm_offset,
                                       n_perm_total, nullptr, true);
    if (rc != FGS_OK) { c->tile_rc = rc; c->tile_err = fgs_last_error(); }
  });
  return FGS_OK;
}

int fgs_null_finish(fgs_ctx *c, double *es_null_out) {
  if (!c) return FGS_ERR_ARG;
  FGS_TRY(use_device(c));   // (waits for the tile in flight)
  if (c->null_S <= 0) { set_error("no resident null"); return FGS_ERR_STATE; }
  const long long C_total = (long long)c->null_R * c->null_P;
  FGS_CUDA(cudaStreamSynchronize(c->copy_stream));
  FGS_TRY(flags_check(c, C_total));   // (synchronises the compute stream)
  if (es_null_out) {
    FGS_CUDA(cudaEventRecord(c->blk_out[0], c->stream));
    FGS_TRY(download_block(c, 0, es_null_out, c->es_null.p, (size_t)c->null_S * C_total));
  }
  return FGS_OK;
}

int fgs_debug_last_scores(fgs_ctx *c, double *scores, int *rank, long long n_col) {
  if (!c) return FGS_ERR_ARG;
  FGS_TRY(use_device(c));
  if (n_col > c->last_cols || c->last_G <= 0) { set_error("debug tap: only %lld columns resident", c->last_cols); return FGS_ERR_STATE; }
  const int G = c->last_G;
  if (scores) FGS_TRY(d2h(c, scores, c->scores.p, (size_t)G * n_col));
  if (rank && c->rank16.n < pitch_of(G) * (size_t)n_col) {
    // the last call did not rank (maxmean / mean need no ranks): nothing to report
    std::fill(rank, rank + (size_t)G * n_col, 0);
  } else if (rank) {
    const size_t Gp = pitch_of(G);
    std::vector<uint16_t> h(Gp * n_col);
    FGS_TRY(d2h(c, h.data(), c->rank16.p, Gp * n_col));
    FGS_TRY(sync(c));
    for (long long col = 0; col < n_col; col++)
      for (int g = 0; g < G; g++) rank[(size_t)col * G + g] = h[(size_t)col * Gp + g];
  }
  FGS_TRY(sync(c));
  return FGS_OK;
}

int fgs_null_device_ptr(fgs_ctx *c, void **dptr, int *n_sets, int *n_response, int *n_perm) {
  if (!c) return FGS_ERR_ARG;
  if (dptr) *dptr = c->es_null.p;
  if (n_sets) *n_sets = c->null_S;
  if (n_response) *n_response = c->null_R;
  if (n_perm) *n_perm = c->null_P;
  return FGS_OK;
}

int fgs_set_null(fgs_ctx *c, const double *es_null, int n_sets, int n_response, int n_perm) {
  if (!c || (!es_null && n_perm > 0) || n_sets <= 0 || n_response <= 0 || n_perm < 0) { set_error("fgs_set_null: bad argument"); return FGS_ERR_ARG; }
  FGS_TRY(use_device(c));
  size_t len = (size_t)n_sets * n_response * n_perm;
  FGS_TRY(c->es_null.ensure(std::max<size_t>(len, 1)));
  FGS_TRY(h2d(c, c->es_null.p, es_null, len));
  FGS_TRY(sync(c));
  c->null_S = n_sets; c->null_R = n_response; c->null_P = n_perm;
  c->null_exact_near_obs = false;  // a null made elsewhere: values within rounding of the observed ES count as equal
  return FGS_OK;
}

// ---------------------------------------------------------------------------
int fgs_observed_es(fgs_ctx *c, int es_kind, double p_exp, const double *scores, int n_gene,
                    int n_response, double *es, double *max_es_at, double *le_prop, int *running_es_at,
                    double *running_es_pos, double *running_es_neg, int *le_first, int *le_count) {
  if (!c || !scores || n_response <= 0 || n_gene <= 0) { set_error("fgs_observed_es: bad argument"); return FGS_ERR_ARG; }
  FGS_TRY(use_device(c));
  if (c->S <= 0) { set_error("fgs_set_gene_sets has not been called"); return FGS_ERR_STATE; }
  if (c->max_set > n_gene) { set_error("gene set index %d exceeds %d genes", c->max_set, n_gene); return FGS_ERR_ARG; }
  const int G = n_gene, S = c->S, R = n_response;
  const size_t SR = (size_t)S * R, NR = (size_t)std::max<long long>(c->nnz, 1) * R;
  FGS_TRY(c->scores.ensure((size_t)G * R));
  FGS_TRY(flags_reset(c, R));
  FGS_TRY(h2d(c, c->scores.p, scores, (size_t)G * R));
  FGS_TRY(launch_check_finite(c, c->scores.p, (long long)G * R, G, c->flags.p));
  // grow-only scratch of the context (cudaMalloc / cudaFree per call cost more than the kernels)
  DevBuf<double> &d_es = c->api_d[0], &d_mea = c->api_d[1], &d_lep = c->api_d[2], &d_pos = c->api_d[3], &d_neg = c->api_d[4];
  DevBuf<int> &d_at = c->api_i[0], &d_lf = c->api_i[1], &d_lc = c->api_i[2];
  auto body = [&]() -> int {
    FGS_TRY(d_es.ensure(SR));
    if (es_kind == FGS_ES_WEIGHTED_KS) {
      const size_t Gp = pitch_of(G);
      FGS_TRY(c->rank16.ensure(Gp * R)); FGS_TRY(c->sabs.ensure(Gp * R));
      FGS_TRY(d_mea.ensure(SR)); FGS_TRY(d_lep.ensure(SR)); FGS_TRY(d_lf.ensure(SR)); FGS_TRY(d_lc.ensure(SR));
      FGS_TRY(d_pos.ensure(NR)); FGS_TRY(d_neg.ensure(NR)); FGS_TRY(d_at.ensure(NR));
      FGS_TRY(launch_sort(c, c->scores.p, G, R, 0, p_exp, c->rank16.p, c->sabs.p));
      FGS_TRY(launch_observed_wks(c, c->rank16.p, c->sabs.p, G, R, d_es.p, d_mea.p, d_lep.p, d_at.p,
                                  d_pos.p, d_neg.p, d_lf.p, d_lc.p));
      FGS_TRY(c->obs_es.ensure(SR));
      FGS_CUDA(cudaMemcpyAsync(c->obs_es.p, d_es.p, SR * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      c->obs_S = S; c->obs_R = R; c->obs_valid = true; c->obs_kind = FGS_ES_WEIGHTED_KS;
      c->obs_p = p_exp; c->obs_score_kind = c->pend_score_kind; c->obs_abs = c->pend_abs; c->pend_score_kind = -1;
      c->last_cols = R; c->last_G = G;
      FGS_TRY(d2h(c, max_es_at, d_mea.p, SR)); FGS_TRY(d2h(c, le_prop, d_lep.p, SR));
      FGS_TRY(d2h(c, le_first, d_lf.p, SR)); FGS_TRY(d2h(c, le_count, d_lc.p, SR));
      FGS_TRY(d2h(c, running_es_at, d_at.p, (size_t)c->nnz * R));
      FGS_TRY(d2h(c, running_es_pos, d_pos.p, (size_t)c->nnz * R));
      FGS_TRY(d2h(c, running_es_neg, d_neg.p, (size_t)c->nnz * R));
    } else if (es_kind == FGS_ES_MAXMEAN || es_kind == FGS_ES_MEAN) {
      // the observed pass sums like the reference (double-double for R's long double): every unit
      const bool saved = c->cur_exact_all;
      c->cur_exact_all = true;
      int r2 = launch_es_mean(c, c->scores.p, G, R, es_kind, 0, d_es.p);
      c->cur_exact_all = saved;
      FGS_TRY(r2);
      FGS_TRY(launch_es_mean_exact(c, c->scores.p, G, R, es_kind, 0, d_es.p));
      FGS_TRY(c->obs_es.ensure(SR));
      FGS_CUDA(cudaMemcpyAsync(c->obs_es.p, d_es.p, SR * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      c->obs_S = S; c->obs_R = R; c->obs_valid = true; c->obs_kind = es_kind;
      c->obs_p = p_exp; c->obs_score_kind = c->pend_score_kind; c->obs_abs = c->pend_abs; c->pend_score_kind = -1;
    } else { set_error("unknown es_kind %d", es_kind); return FGS_ERR_ARG; }
    FGS_TRY(d2h(c, es, d_es.p, SR));
    return flags_check(c, R);
  };
  int rc = body();
  cudaStreamSynchronize(c->stream);
  return rc;
}

int fgs_set_observed_es(fgs_ctx *c, int es_kind, const double *es, int n_sets, int n_response) {
  if (!c || !es || n_sets <= 0 || n_response <= 0) { set_error("fgs_set_observed_es: bad argument"); return FGS_ERR_ARG; }
  FGS_TRY(use_device(c));
  if (c->S != n_sets) { set_error("fgs_set_observed_es: %d sets given, %d resident", n_sets, c->S); return FGS_ERR_ARG; }
  const size_t SR = (size_t)n_sets * n_response;
  FGS_TRY(c->obs_es.ensure(SR));
  FGS_TRY(h2d(c, c->obs_es.p, es, SR));
  FGS_TRY(sync(c));
  c->obs_S = n_sets; c->obs_R = n_response; c->obs_kind = es_kind; c->obs_valid = true;
  c->obs_score_kind = -1;   // made elsewhere: never copied into the null, only used to spot near-ties
  return FGS_OK;
}

// ---------------------------------------------------------------------------
// significance
static int sig_check(fgs_ctx *c, int response) {
  if (c->null_S <= 0 || c->null_R <= 0) { set_error("no resident null: run fgs_perm_null / fgs_es_from_scores / fgs_set_null first"); return FGS_ERR_STATE; }
  if (response < 0 || response >= c->null_R) { set_error("response %d out of range (%d)", response, c->null_R); return FGS_ERR_ARG; }
  return FGS_OK;
}

int fgs_sig_stage1(fgs_ctx *c, const double *es, int response, int split_p, int abs_flag,
                   double *sums, long long *counts) {
  if (!c || !es) return FGS_ERR_ARG;
  FGS_TRY(use_device(c));
  FGS_TRY(sig_check(c, response));
  const int S = c->null_S;
  DevBuf<double> &d_es = c->api_d[0], &d_sums = c->api_d[1]; DevBuf<long long> &d_counts = c->api_l;
  auto body = [&]() -> int {
    FGS_TRY(d_es.ensure(S)); FGS_TRY(d_sums.ensure((size_t)S * 4)); FGS_TRY(d_counts.ensure((size_t)S * 6));
    FGS_TRY(h2d(c, d_es.p, es, S));
    FGS_TRY(launch_sig_stage1(c, c->es_null.p, S, c->null_R, c->null_P, response, d_es.p, split_p,
                              abs_flag, d_sums.p, d_counts.p));
    FGS_TRY(d2h(c, sums, d_sums.p, (size_t)S * 4)); FGS_TRY(d2h(c, counts, d_counts.p, (size_t)S * 6));
    return sync(c);
  };
  int rc = body();
  return rc;
}

int fgs_sig_stage2(fgs_ctx *c, const double *es, int response, int split_p, int calc_nes,
                   int abs_flag, const double *sums_parts, int n_parts, const long long *counts_total,
                   long long n_perm_total, double *p, double *p_high, double *p_low, double *nes,
                   double *fwer, double *fdr_bh, long long *null_counts, long long *glob) {
  if (!c || !es || !sums_parts || !counts_total || n_parts <= 0) return FGS_ERR_ARG;
  FGS_TRY(use_device(c));
  FGS_TRY(sig_check(c, response));
  const int S = c->null_S;
  DevBuf<double> &d = c->api_d[0]; DevBuf<long long> &di = c->api_l;
  auto body = [&]() -> int {
    // doubles: es[S] parts[n_parts*S*4] p ph pl mpos mneg nes fwer fdr [8S]
    FGS_TRY(d.ensure((size_t)S * (9 + 4 * (size_t)n_parts)));
    FGS_TRY(di.ensure((size_t)S * 7 + 4));
    double *d_es = d.p, *d_parts = d_es + S, *d_p = d_parts + (size_t)n_parts * S * 4, *d_ph = d_p + S,
           *d_pl = d_ph + S, *d_mp = d_pl + S, *d_mn = d_mp + S, *d_nes = d_mn + S, *d_fw = d_nes + S,
           *d_bh = d_fw + S;
    long long *d_cnt = di.p, *d_nc = d_cnt + (size_t)S * 6, *d_gl = d_nc + S;
    FGS_TRY(h2d(c, d_es, es, S));
    FGS_TRY(h2d(c, d_parts, sums_parts, (size_t)n_parts * S * 4));
    FGS_TRY(h2d(c, d_cnt, counts_total, (size_t)S * 6));
    FGS_TRY(launch_sig_means(c, d_parts, n_parts, (size_t)S * 4, d_cnt, d_es, S, n_perm_total, split_p, calc_nes,
                             abs_flag, d_p, d_ph, d_pl, d_mp, d_mn, d_nes, d_fw));
    if (calc_nes)
      FGS_TRY(launch_sig_stage2(c, c->es_null.p, S, c->null_R, c->null_P, response, d_es, abs_flag,
                                d_mp, d_mn, d_nes, d_nc, d_gl));
    FGS_TRY(d2h(c, p, d_p, S)); FGS_TRY(d2h(c, p_high, d_ph, S)); FGS_TRY(d2h(c, p_low, d_pl, S));
    FGS_TRY(d2h(c, nes, d_nes, S)); FGS_TRY(d2h(c, fwer, d_fw, S));
    if (calc_nes) { FGS_TRY(d2h(c, null_counts, d_nc, S)); FGS_TRY(d2h(c, glob, d_gl, 2)); }
    else { FGS_TRY(launch_bh(c, d_p, S, d_bh)); FGS_TRY(d2h(c, fdr_bh, d_bh, S)); }
    return sync(c);
  };
  int rc = body();
  return rc;
}

int fgs_sig_finish(fgs_ctx *c, const double *nes, const long long *null_counts,
                   const long long *glob, int n_sets, int abs_flag, double *fdr,
                   long long *fdr_counts, long long *glob_counts) {
  if (!c || !nes || !null_counts || !glob || n_sets <= 0) return FGS_ERR_ARG;
  FGS_TRY(use_device(c));
  const int S = n_sets;
  DevBuf<double> &d = c->api_d[0]; DevBuf<long long> &di = c->api_l;
  auto body = [&]() -> int {
    FGS_TRY(d.ensure((size_t)S * 2)); FGS_TRY(di.ensure((size_t)S * 3 + 8));
    double *d_nes = d.p, *d_fdr = d_nes + S;
    long long *d_nc = di.p, *d_gl = d_nc + S, *d_fc = d_gl + 2, *d_gc = d_fc + (size_t)S * 2;
    FGS_TRY(h2d(c, d_nes, nes, S)); FGS_TRY(h2d(c, d_nc, null_counts, S)); FGS_TRY(h2d(c, d_gl, glob, 2));
    FGS_TRY(launch_sig_finish(c, d_nes, d_nc, d_gl, S, abs_flag, d_fdr, d_fc, d_gc));
    FGS_TRY(d2h(c, fdr, d_fdr, S)); FGS_TRY(d2h(c, fdr_counts, d_fc, (size_t)S * 2));
    FGS_TRY(d2h(c, glob_counts, d_gc, 4));
    return sync(c);
  };
  int rc = body();
  return rc;
}

}  // extern "C" (reopened below)

namespace fgs {
// flexgsea_calc_sig on the resident null of this context.  With a communicator (sharded = true)
// the null holds this rank's span of the permutations and the three small exchanges of
// include/flexgsea_b200.h happen here, on the context's stream, between the kernels: nothing
// goes through the host.
int calc_sig_core(fgs_ctx *c, const double *es, int response, int split_p, int calc_nes, int abs_flag,
                  long long P_total, bool sharded, double *p, double *p_high, double *p_low, double *nes,
                  double *fdr, double *fwer, long long *p_counts, long long *fdr_counts,
                  long long *glob_counts) {
  const int S = c->null_S, P = c->null_P;
  const int nparts = sharded ? c->comm_size : 1, part = sharded ? c->comm_rank : 0;
  for (int i = 0; i < 8; i++)
    if (!c->sig_ev[i]) FGS_CUDA(cudaEventCreate(&c->sig_ev[i]));
  DevBuf<double> &d = c->api_d[0]; DevBuf<long long> &di = c->api_l;
  int n_coll = 0;
  auto coll_begin = [&]() { if (sharded && n_coll < 3) cudaEventRecord(c->sig_ev[2 + 2 * n_coll], c->stream); };
  auto coll_end = [&]() { if (sharded && n_coll < 3) cudaEventRecord(c->sig_ev[3 + 2 * n_coll], c->stream); n_coll++; };
  auto body = [&]() -> int {
    // doubles: es parts[nparts * 10 S] p ph pl mpos mneg nes fwer fdr.  A rank's part is its stage-1
    // sums [S x 4 doubles] with its counts [S x 6 int64] right behind: ONE all-gather carries both
    // (the counts are summed over the parts here; the sums are combined in rank order).
    const size_t part_dbl = (size_t)S * 10;
    FGS_TRY(d.ensure((size_t)S * 9 + part_dbl * (size_t)nparts)); FGS_TRY(di.ensure((size_t)S * 9 + 8));
    double *d_es = d.p, *d_parts = d_es + S, *d_p = d_parts + part_dbl * (size_t)nparts, *d_ph = d_p + S, *d_pl = d_ph + S,
           *d_mp = d_pl + S, *d_mn = d_mp + S, *d_nes = d_mn + S, *d_fw = d_nes + S, *d_fdr = d_fw + S;
    long long *d_cnt = di.p, *d_nc = d_cnt + (size_t)S * 6, *d_gl = d_nc + S, *d_fc = d_gl + 2, *d_gc = d_fc + (size_t)S * 2;
    FGS_CUDA(cudaEventRecord(c->sig_ev[0], c->stream));
    FGS_TRY(h2d(c, d_es, es, S));
    double *d_sums = d_parts + (size_t)part * part_dbl;  // this rank's slot: the all-gather is in place
    FGS_TRY(launch_sig_stage1(c, c->es_null.p, S, c->null_R, P, response, d_es, split_p, abs_flag, d_sums,
                              (long long *)(d_sums + (size_t)S * 4)));
    if (sharded) {
      coll_begin();
      FGS_TRY(comm_allgather_f64(c, d_sums, d_parts, part_dbl));
      coll_end();
    }
    FGS_TRY(launch_sig_sum_counts(c, d_parts, nparts, part_dbl, S, d_cnt));
    FGS_TRY(launch_sig_means(c, d_parts, nparts, part_dbl, d_cnt, d_es, S, P_total, split_p, calc_nes, abs_flag, d_p, d_ph,
                             d_pl, d_mp, d_mn, d_nes, d_fw));
    if (calc_nes) {
      FGS_TRY(launch_sig_stage2(c, c->es_null.p, S, c->null_R, P, response, d_es, abs_flag, d_mp, d_mn,
                                d_nes, d_nc, d_gl));
      if (sharded) {  // [S] exceedance counts + the two global counts sit back to back
        coll_begin();
        FGS_TRY(comm_allreduce_i64(c, d_nc, (size_t)S + 2));
        coll_end();
      }
      FGS_TRY(launch_sig_finish(c, d_nes, d_nc, d_gl, S, abs_flag, d_fdr, d_fc, d_gc));
    } else {
      FGS_TRY(launch_bh(c, d_p, S, d_fdr));
    }
    FGS_CUDA(cudaEventRecord(c->sig_ev[1], c->stream));
    FGS_TRY(d2h(c, p, d_p, S)); FGS_TRY(d2h(c, p_high, d_ph, S)); FGS_TRY(d2h(c, p_low, d_pl, S));
    FGS_TRY(d2h(c, nes, d_nes, S)); FGS_TRY(d2h(c, fdr, d_fdr, S)); FGS_TRY(d2h(c, fwer, d_fw, S));
    if (p_counts) {  // (numerator-1, denominator-1) of p, or (high, low) counts
      std::vector<long long> h((size_t)S * 6);
      FGS_TRY(d2h(c, h.data(), d_cnt, (size_t)S * 6));
      FGS_TRY(sync(c));
      for (int s = 0; s < S; s++) {
        if (split_p) {
          bool pos_side = (es[s] >= 0.0) || abs_flag;
          p_counts[2 * s] = pos_side ? h[(size_t)s * 6 + 3] : h[(size_t)s * 6 + 4];
          p_counts[2 * s + 1] = pos_side ? h[(size_t)s * 6 + 0] : h[(size_t)s * 6 + 2];
        } else { p_counts[2 * s] = h[(size_t)s * 6 + 3]; p_counts[2 * s + 1] = h[(size_t)s * 6 + 4]; }
      }
    }
    if (calc_nes) { FGS_TRY(d2h(c, fdr_counts, d_fc, (size_t)S * 2)); FGS_TRY(d2h(c, glob_counts, d_gc, 4)); }
    return sync(c);
  };
  const int rc = body();
  if (rc == FGS_OK) {
    float t = 0.f;
    c->last_sig_ms[1] = 0.f;
    cudaEventElapsedTime(&c->last_sig_ms[0], c->sig_ev[0], c->sig_ev[1]);
    for (int k = 0; k < std::min(n_coll, 3); k++)
      if (cudaEventElapsedTime(&t, c->sig_ev[2 + 2 * k], c->sig_ev[3 + 2 * k]) == cudaSuccess) c->last_sig_ms[1] += t;
    // one step on the device timeline: start of the most recent null loop -> end of this call
    if (cudaEventElapsedTime(&t, c->ev[0], c->sig_ev[1]) == cudaSuccess) c->last_sig_ms[2] = t;
    else { cudaGetLastError(); c->last_sig_ms[2] = 0.f; }
    c->last_sig_ms[3] = (float)n_coll;
  }
  return rc;
}
}  // namespace fgs

extern "C" {

int fgs_calc_sig(fgs_ctx *c, const double *es, int response, int split_p, int calc_nes, int abs_flag,
                 double *p, double *p_high, double *p_low, double *nes, double *fdr, double *fwer,
                 long long *p_counts, long long *fdr_counts, long long *glob_counts) {
  if (!c || !es) return FGS_ERR_ARG;
  FGS_TRY(use_device(c));
  FGS_TRY(sig_check(c, response));
  if (c->null_P == 0) return FGS_OK;  // R/flexgsea.R:575-577
  return calc_sig_core(c, es, response, split_p, calc_nes, abs_flag, c->null_P, false, p, p_high, p_low, nes,
                       fdr, fwer, p_counts, fdr_counts, glob_counts);
}

int fgs_last_sig_timings(fgs_ctx *c, float *ms) {
  if (!c || !ms) return FGS_ERR_ARG;
  for (int i = 0; i < 4; i++) ms[i] = c->last_sig_ms[i];
  return FGS_OK;
}

}  // extern "C"
