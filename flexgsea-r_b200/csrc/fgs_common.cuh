// Shared declarations of the flexgsea sm_100a library (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <thread>
#include <vector>

#include "flexgsea_b200.h"

namespace fgs {

void set_error(const char *fmt, ...);

#define FGS_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      fgs::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
                     cudaGetErrorString(e__));                                 \
      return e__ == cudaErrorMemoryAllocation ? FGS_ERR_OOM : FGS_ERR_CUDA;    \
    }                                                                          \
  } while (0)

#define FGS_TRY(call)                  \
  do {                                 \
    int rc__ = (call);                 \
    if (rc__ != FGS_OK) return rc__;   \
  } while (0)

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;  // capacity in elements
  int ensure(size_t want) {
    if (want <= n) return FGS_OK;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
    if (e != cudaSuccess) {
      set_error("cudaMalloc of %zu bytes failed: %s", want * sizeof(T), cudaGetErrorString(e));
      cudaGetLastError();
      return FGS_ERR_OOM;
    }
    n = want;
    return FGS_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// near-tie certification of the ranks derived from tensor-core scores
struct TieCert {
  int2 *list = nullptr;   // (column, gene) entries to recompute exactly
  int *count = nullptr;   // device counter
  int cap = 0;
  int *redo = nullptr;    // [C] columns to rank again
  int *overflow = nullptr;  // bit 16 raised when the list is full
  double rel_tol = 0, abs_tol = 0;
  // dup[g] = first gene whose expression row is bit-identical to gene g's (NULL: no duplicated
  // rows).  Such genes score identically in the reference AND here, whatever the arithmetic, so
  // an exact tie between them needs no recomputation: index order decides in both.
  const int *dup = nullptr;
};

// error bits accumulated on the device during a run
enum : int { FLAG_POS_INF = 1, FLAG_NEG_INF = 2, FLAG_NAN = 4 };

}  // namespace fgs

struct fgs_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 0;
  int smem_optin = 0;  // max dynamic shared memory per block
  long long launches = 0;
  bool profiling = false;
  float last_ms[5] = {0, 0, 0, 0, 0};
  long long last_launches[5] = {0, 0, 0, 0, 0};
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // user-score streaming (fgs_es_from_scores): upload of block k+1 on copy_stream while
  // block k is ranked; blk_copy / blk_done order the two score buffers
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t blk_copy[2] = {nullptr, nullptr}, blk_done[2] = {nullptr, nullptr}, blk_out[2] = {nullptr, nullptr};
  // pinned staging ring of the threaded host->device upload (fgs_api.cu: staged_upload)
  char *stage = nullptr;
  size_t stage_bytes = 0;
  cudaEvent_t stage_ev[16] = {};
  int opt_h2d_threads = 8;
  int (*progress)(void *, long long, long long) = nullptr;
  void *progress_user = nullptr;

  // fgs_es_from_scores_tile: the tile handed over last is uploaded, ranked and scored by this
  // helper thread while the caller scores the next one; joined by the next call on the context
  std::thread tile_job;
  int tile_rc = 0;
  std::string tile_err;

  // options
  int opt_perm_block = 0;      // 0 = auto
  int opt_force_generic = 0;
  int opt_exact_small = 1;     // small problems: every ES unit in the reference's operation order
  double sum_m2 = 0.0;         // sum of squared set sizes (cost of that path)
  bool cur_exact_all = false;
  int opt_sort64 = 0;          // force the 64-bit global-memory sort (tests)
  int opt_s2n_mode = 1;        // 1 = tensor-core contraction + rank certification, 0 = bit-faithful kernel
  int opt_cert_cap = 4 << 20;  // entries per certification list (epilogue list, tie list)
  int opt_score_ppt = 0;
  long long opt_nperm_total = 0;  // permutations over ALL ranks when the caller shards them (0: this call's)
  int opt_bulk_stage = 1;      // weight / score columns staged by cp.async.bulk (0: the ld.global / st.shared loop)
  int opt_arrange = 1;         // member blocks dealt out by bank for the maxmean / mean gathers (arrange_idx16_kernel)
  int opt_tail_split = 1;      // split the columns of a partial last round over the idle SMs (ES kernels)

  // expression matrix x [n x G] (gene's samples contiguous)
  int n = 0, G = 0;
  fgs::DevBuf<double> x;
  fgs::DevBuf<double> xc;       // centred expression (tensor-core s2n), [G][n]
  fgs::DevBuf<double> xtot;     // per gene {sum xc, sum xc^2}
  bool xc_valid = false;
  fgs::DevBuf<int2> cert_list;  // (column, gene) entries for exact recomputation
  fgs::DevBuf<int> cert_ctr;    // [0] epilogue count, [1] tie count
  fgs::DevBuf<int> cert_redo;   // [Cb]
  fgs::DevBuf<uint16_t> sort_glist;  // [grid][Gp] gene lists of the score-range buckets (long columns)
  fgs::DevBuf<int> sort_redo;   // [Cb] columns the 32-bit sort hands to the 64-bit sort
  fgs::DevBuf<int> dupclass;    // [G] first gene with a bit-identical expression row (see TieCert::dup)
  bool has_dups = false;
  fgs::DevBuf<double> gene_sd;  // sd(x[,g]) for lm
  bool gene_sd_valid = false;

  // gene sets
  int S = 0;
  long long nnz = 0;
  int max_set = 0;
  fgs::DevBuf<int> set_off;        // [S+1]
  fgs::DevBuf<int> set_idx;        // [nnz] 0-based gene ids
  // [S] in processing order (descending size):
  //   {member offset, m | slow << 31, offset in set_idx16 in blocks of 8, set id};
  // "slow" = duplicated members or more than ES_MAX_SET members (order-free kernel)
  fgs::DevBuf<int4> set_info;
  // es_wks_kernel work list: {first set in processing order, count | lanes per set << 8 |
  // log2(members per lane) << 16}: the sets one warp takes at once (fgs_set_gene_sets)
  fgs::DevBuf<int2> set_groups;
  int n_groups = 0, n_wide_groups = 0;  // the wide ones (a set per warp) come first
  fgs::DevBuf<double> set_invgm;   // [S] 1 / (G - m), processing order
  // observed ES of the last fgs_observed_es (weighted KS): null units that land within
  // rounding of it are recomputed in the reference's operation order (es_wks_tie_kernel)
  fgs::DevBuf<double> obs_es;      // [S x R]
  int obs_S = 0, obs_R = 0, obs_kind = -1;
  bool obs_valid = false;
  // what the observed ES was made from: weight exponent, and -- when its scores came from
  // fgs_score_observed of this context -- the score kind and abs flag (-1: unknown provenance)
  double obs_p = 1.0;
  int obs_score_kind = -1, obs_abs = 0;
  int pend_score_kind = -1, pend_abs = 0;   // the last fgs_score_observed, until an observed pass takes it
  const double *cur_obs = nullptr;   // set by rank_and_es for the ES launches of the current call
  int cur_obs_R = 1;
  const int *cur_ident = nullptr;    // [Pb] label-preserving permutations of the current block (or NULL)
  fgs::DevBuf<int> ident;            // [P]
  fgs::DevBuf<uint32_t> obs_labels;  // [R][2][W] class masks of the unpermuted response
  bool null_exact_near_obs = false;  // the resident null was made with that treatment
  fgs::DevBuf<int2> tie_list;      // (column, set) units whose es_pos and -es_neg tied: redone exactly
  fgs::DevBuf<int> tie_ctr;
  int tie_cap = 0;
  fgs::DevBuf<unsigned char> set_slow;  // [S] flags + [16] status bytes of launch_prep_sets
  fgs::DevBuf<uint16_t> set_idx16; // members as uint16, processing order, 8-blocks padded with gene G
  long long idx16_blocks = 0;
  int invgm_G = -1;                // G set_invgm / set_idx16 were filled for (-1: stale)
  std::vector<int> h_set_off;
  int n_slow = 0;
  fgs::DevBuf<int> slow_ids;       // [n_slow]

  // s2n response: y_bin [n x R]
  int s2n_R = 0;
  bool s2n_has_na = false, s2n_has_na_pending = false;  // NA_LOGICAL present in y_bin
  bool cur_cert = false;            // scores of the current block came from the tensor-core path
  int *cur_ovf = nullptr;           // this block's two overflow flags (epilogue list, tie list) in blk_ovf
  int cur_P = 0;                    // permutations of the current block
  const uint32_t *cur_labels = nullptr;
  long long cur_flag0 = 0;
  fgs::DevBuf<int> ybin;
  fgs::DevBuf<int> n1n2;  // [R][2]
  fgs::DevBuf<double> inv_n1n2;  // [R][2] reciprocals (tensor-core score epilogue)
  // lm response
  int lm_q = 0, lm_R = 0, lm_icpt = 0;
  fgs::DevBuf<double> lm_H;  // [R x n] rows of R^-1 Q^T for the scored columns
  std::vector<double> h_design;

  // permutation machinery
  fgs::DevBuf<int> perm;          // [n x Pb] 1-based
  fgs::DevBuf<uint32_t> labels;   // [Pb*R][2][W]
  fgs::DevBuf<uint16_t> class_lists;  // [Cb][n] samples of class 1 then class 2 per column (bit-faithful s2n)
  fgs::DevBuf<int> perm_map;      // [P] permutation -> distinct labelling (few samples: fgs_perm_null)
  fgs::DevBuf<double> es_uniq;    // [S, R, U] ES of the distinct labellings
  int opt_dedup = 1, last_dedup = 0, last_fallback_blocks = 0;
  std::vector<int> h_n1n2;        // host copy of the class sizes [R][2]
  int s2n_n = 0;
  fgs::DevBuf<int> blk_ovf;       // [2 * blocks] certification-list overflow of a block (epilogue list, tie list)
  fgs::DevBuf<double> lm_B;       // [n x Cb] permuted design block

  // per-block scratch
  fgs::DevBuf<double> scores;     // [Cb][G]
  fgs::DevBuf<double> scores2;    // second block of user scores being uploaded (fgs_es_from_scores)
  fgs::DevBuf<uint16_t> rank16;   // [Cb][G] 1-based descending rank
  fgs::DevBuf<double> sabs;       // [Cb][G] |score|^p in rank order
  fgs::DevBuf<unsigned long long> sort_keys;  // ping-pong scratch
  fgs::DevBuf<uint32_t> sort_idx;
  fgs::DevBuf<int> flags;         // per-perm non-finite flags + [last] global
  fgs::DevBuf<int> counters;      // work counters
  long long last_cols = 0;        // columns held by scores/rank16 (debug taps)
  int last_G = 0;

  // resident null distribution es_null [S, R, P]
  fgs::DevBuf<double> es_null;
  int null_S = 0, null_R = 0, null_P = 0;

  // significance scratch
  fgs::DevBuf<double> sig_d;
  fgs::DevBuf<long long> sig_i;
  fgs::DevBuf<double> api_d[5];    // grow-only scratch of the observed / significance entry points
  fgs::DevBuf<int> api_i[3];
  fgs::DevBuf<long long> api_l;
  fgs::DevBuf<double> sig_t;       // [n1 + n2] the two threshold lists as one ascending list
  fgs::DevBuf<int> sig_lut;        // coarse cell index over the sorted NES thresholds (pooled FDR)

  // permutations sharded over GPUs (fgs_multi.cu): NCCL communicator of this rank
  void *comm = nullptr;            // ncclComm_t
  int comm_rank = 0, comm_size = 1;
  bool comm_owned = true;          // false: the communicator belongs to a fgs_group
  fgs::DevBuf<double> null_full;   // all-gathered null (fgs_comm_allgather_null), swapped with es_null
  cudaEvent_t sig_ev[8] = {};      // [0],[1] significance call; [2..7] around its collectives
  float last_sig_ms[4] = {0, 0, 0, 0};
};

namespace fgs {

// ---- kernel launch wrappers (each returns FGS_OK / error) ------------------
// K1
int launch_philox_perm(fgs_ctx *c, unsigned long long seed, long long perm_id0, int n, int P,
                       int *d_perm);
int launch_labels(fgs_ctx *c, const int *d_perm, const int *d_ybin, int n, int R, int P,
                  uint32_t *d_labels);
// K2'
// only_if / unless: device flags that make the launch conditional (block-level fallback of the tensor path)
int launch_s2n(fgs_ctx *c, const double *d_x, int n, int G, const uint32_t *d_labels,
               const int *d_n1n2, int R, long long C, double *d_scores, int *d_flags,
               const int *only_if = nullptr, const int *unless = nullptr);
int launch_s2n_patch(fgs_ctx *c, double *d_scores, int G, int R, int P, int *d_flags);
// K2 (lm)
int launch_lm_build_B(fgs_ctx *c, const double *d_H, const int *d_perm, int n, int R, int P,
                      double *d_B);
int launch_lm_gemm(fgs_ctx *c, const double *d_x, int n, int G, const double *d_B, long long C,
                   const double *d_gene_sd, double *d_scores, int *d_flags, int R);
int launch_gene_sd(fgs_ctx *c, const double *d_x, int n, int G, double *d_sd);
int launch_row_hash(fgs_ctx *c, const double *d_x, int n, int G, unsigned long long *d_hash);
// generic helpers
int launch_check_finite(fgs_ctx *c, const double *d_v, long long len, long long per_flag,
                        int *d_flags);
int launch_abs_inplace(fgs_ctx *c, double *d_v, long long len);
// K3
int launch_sort(fgs_ctx *c, const double *d_scores, int G, long long C, int abs_flag, double p_exp,
                uint16_t *d_rank16, double *d_sabs, const int *only_cols = nullptr,
                const TieCert *cert = nullptr);
int launch_center(fgs_ctx *c, const double *d_x, int n, int G, double *d_xc, double *d_tot);
int launch_s2n_mma(fgs_ctx *c, const double *d_xc, const double *d_tot, int n, int G,
                   const uint32_t *d_labels, const double *d_inv_n, int R, long long C, double *d_scores,
                   int2 *d_list, int *d_count, int cap, int *d_overflow);
int launch_mark_all_if(fgs_ctx *c, int *d_redo, long long C, const int *only_if, const int *unless);
int launch_s2n_exact_entries(fgs_ctx *c, const double *d_x, int n, int G, const uint32_t *d_labels,
                             const int *d_n1n2, int R, const int2 *d_list, const int *d_count, int cap,
                             double *d_scores, int *d_flags, bool after_patch = false,
                             const int *unless = nullptr);
// K4 / K5
int launch_prep_sets(fgs_ctx *c, int S, int *h_status, unsigned char *h_slow);
int launch_es_wks(fgs_ctx *c, const uint16_t *d_rank16, const double *d_sabs, int G, long long C,
                  double *d_es, bool defer_ties = false);
int launch_scatter_null(fgs_ctx *c, const double *d_uniq, const int *d_map, long long unit, int P, double *d_null);
int launch_copy_observed(fgs_ctx *c, const double *d_obs, const int *d_ident, int R, long long C,
                         double *d_es);
int launch_label_identity(fgs_ctx *c, const uint32_t *d_labels, const uint32_t *d_obs, int R, int W, int P,
                          int *d_ident);
int ensure_tie_list(fgs_ctx *c);
int launch_es_mean_exact(fgs_ctx *c, const double *d_scores, int G, long long C, int es_kind,
                         int abs_flag, double *d_es);
int launch_es_wks_ties(fgs_ctx *c, const uint16_t *d_rank16, const double *d_sabs, int G, long long C,
                       double *d_es);
int launch_tie_members(fgs_ctx *c, int2 *d_out, int *d_out_ctr, int out_cap, int *d_overflow);
int launch_tie_refresh_weights(fgs_ctx *c, const int2 *d_list, const int *d_ctr, int cap,
                               const double *d_scores, const uint16_t *d_rank16, double *d_sabs, int G,
                               double p_exp);
int launch_es_wks_generic(fgs_ctx *c, const uint16_t *d_rank16, const double *d_sabs, int G,
                          long long C, const int *d_set_ids, int n_ids, double *d_es);
int launch_es_mean(fgs_ctx *c, const double *d_scores, int G, long long C, int es_kind,
                   int abs_flag, double *d_es);
int launch_observed_wks(fgs_ctx *c, const uint16_t *d_rank16, const double *d_sabs, int G, int R,
                        double *d_es, double *d_max_es_at, double *d_le_prop, int *d_at,
                        double *d_pos, double *d_neg, int *d_le_first, int *d_le_count);
// K6 / K7
int launch_sig_stage1(fgs_ctx *c, const double *d_null, int S, int R, int P, int response,
                      const double *d_es, int split_p, int abs_flag, double *d_sums,
                      long long *d_counts);
int launch_sig_stage2(fgs_ctx *c, const double *d_null, int S, int R, int P, int response,
                      const double *d_es, int abs_flag, const double *d_mpos, const double *d_mneg,
                      double *d_nes, long long *d_null_counts, long long *d_glob);
int launch_sig_finish(fgs_ctx *c, const double *d_nes, const long long *d_null_counts,
                      const long long *d_glob, int S, int abs_flag, double *d_fdr,
                      long long *d_fdr_counts, long long *d_glob_counts);

int settle_tile(fgs_ctx *c);  // joins the tile in flight (fgs_es_from_scores_tile), reports its failure
// collectives of the significance exchange on c->stream (fgs_multi.cu); no-ops without a communicator
int comm_allgather_f64(fgs_ctx *c, const double *send, double *recv, size_t count);
int comm_allreduce_i64(fgs_ctx *c, long long *buf, size_t count);

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
// row pitch (elements) of the per-column rank16 / sabs arrays: 16-byte aligned rows
// hit weight |s|^p (R/flexgsea.R:819).  R's `^` returns x * x for p == 2 and calls the C
// library's pow otherwise; CUDA's pow is good to 2 ulp, glibc's to ~0.5, so only the
// exponents with an exact form (0, 1, 2) give bit-identical weights -- enough for the
// result everywhere (ES tolerance 1e-10) except the sign at an exact structural tie.
__device__ __forceinline__ double hit_weight(double a, double p_exp) {
  if (p_exp == 1.0) return a;
  if (p_exp == 2.0) return __dmul_rn(a, a);
  return pow(a, p_exp);
}

// largest set the in-register weighted-KS kernel takes (64 ranks per lane on a whole warp)
constexpr int ES_MAX_SET = 2048;
__host__ __device__ inline size_t pitch_of(int G) { return ((size_t)G + 7) & ~(size_t)7; }

#ifdef __CUDACC__
// Bulk copies global -> shared by the copy engine (cp.async.bulk, SASS UBLKCP) with an mbarrier that
// counts the bytes: ONE thread issues the copy of a whole column and the CTA's threads do other
// staging work until they wait for it, instead of every thread looping over ld.global / st.shared.
// Source, destination and size are multiples of 16 bytes.
__device__ __forceinline__ void bulk_bar_init(uint64_t *bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// (called by one thread, after a __syncthreads() behind the last generic-proxy access of `dst`)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void bulk_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(b), "r"(parity) : "memory");
}
#endif

}  // namespace fgs
