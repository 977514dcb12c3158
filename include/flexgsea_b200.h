/*
 * flexgsea_b200.h -- C ABI of the sm_100a library that replaces flexgsea's
 * sample-permutation enrichment loop.
 *
 * Drop-in boundary: the reference package reaches native
 * code through R's .Call into its DLL -- one registered routine,
 * `_flexgsea_s2n_C(SEXP x, SEXP y)` (reference src/RcppExports.cpp:10-24),
 * called from R/RcppExports.R:4-6.  Everything else on the hot path is
 * interpreted R (R/flexgsea.R:356-472, :474-629, :725-937).  The entry points
 * below are what an Rcpp shim (r-package/src/flexgsea_native.cpp, see
 * INTEGRATION.md) binds: plain pointers and sizes, no R, torch or CUDA types.
 *
 * Conventions
 *  - every matrix is column-major, exactly as R holds it; inputs are read-only
 *    host memory unless a parameter says "device";
 *  - gene-set members are 1-based gene indices as produced by
 *    match(set, gene.names) (R/flexgsea.R:279, :412), duplicates kept;
 *  - every function returns FGS_OK or an error code; fgs_last_error() gives the
 *    message (thread-local).  Nothing here longjmps or throws across the ABI.
 *  - there is no CPU fallback: without a CUDA device fgs_create fails.
 */
#ifndef FLEXGSEA_B200_H
#define FLEXGSEA_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGS_ABI_VERSION 2

#define FGS_OK 0
#define FGS_ERR_ARG 1
#define FGS_ERR_CUDA 2
/* "Gene scoring function returned NA or Inf." R/flexgsea.R:251-253, :400-402 */
#define FGS_ERR_NONFINITE 3
#define FGS_ERR_OOM 4
#define FGS_ERR_STATE 5
#define FGS_ERR_UNSUPPORTED 6
#define FGS_ERR_CANCELLED 7 /* the progress callback asked to stop */
#define FGS_ERR_COMM 8      /* NCCL missing or a collective failed */

/* gene.score.fn built-ins: flexgsea_s2n (R/flexgsea.R:689-723),
 * flexgsea_lm (:647-676) */
#define FGS_SCORE_S2N 0
#define FGS_SCORE_LM 1
/* es.fn built-ins: flexgsea_weighted_ks (:925-937), flexgsea_maxmean
 * (:749-753), flexgsea_mean (:775-779) */
#define FGS_ES_WEIGHTED_KS 0
#define FGS_ES_MAXMEAN 1
#define FGS_ES_MEAN 2

typedef struct fgs_ctx fgs_ctx;

int fgs_abi_version(void);
const char *fgs_last_error(void);
int fgs_device_count(int *count);

/* One context per GPU (device ordinal); owns the stream and all device state.
 * The R side keeps it in an external pointer with a finalizer. */
int fgs_create(int device, fgs_ctx **ctx);
int fgs_destroy(fgs_ctx *ctx);

/* ---- s2n_C: reference src/ggsea.cpp:6-76, .Call entry
 * src/RcppExports.cpp:10-19.  x [n_sample_x x n_gene] double, y
 * [n_sample_y x n_response] int logical (NA_LOGICAL = INT_MIN is skipped),
 * coef [n_gene x n_response].  Bit-identical to the reference: sequential
 * sample order, its three passes, population variance, no FMA.  Row mismatch
 * returns an all-zero matrix like src/ggsea.cpp:15-17. */
int fgs_s2n(fgs_ctx *ctx, const double *x, int n_sample_x, int n_gene, const int *y,
            int n_sample_y, int n_response, double *coef);

/* ---- problem upload (once per flexgsea() call) ------------------------- */
/* x [n_sample x n_gene] (R/flexgsea.R:159-173, t.x == FALSE layout) */
int fgs_set_x(fgs_ctx *ctx, const double *x, int n_sample, int n_gene);
/* filtered gene sets as CSR (replaces the per-set match() of :279 and :412) */
int fgs_set_gene_sets(fgs_ctx *ctx, const int *offsets, const int *index, int n_sets);
/* flexgsea_s2n's y_bin (:704-710): [n_sample x n_response] int logical */
int fgs_set_response_s2n(fgs_ctx *ctx, const int *y_bin, int n_sample, int n_response);
/* flexgsea_lm's model matrix (:648-663): [n_sample x n_col] double;
 * has_intercept != 0 drops the first column from the scores (:667-669).
 * FGS_ERR_NONFINITE when the design is rank deficient (lm.fit gives NA). */
int fgs_set_response_lm(fgs_ctx *ctx, const double *design, int n_sample, int n_col,
                        int has_intercept);
/* number of score columns R the current response yields for score_kind */
int fgs_n_response(fgs_ctx *ctx, int score_kind, int *n_response);

/* ---- observed pass (R/flexgsea.R:246-292) ------------------------------ */
/* gene.score.fn(x, y, abs): scores [n_gene x R], +-Inf patch of :715-716 and
 * abs applied.  FGS_ERR_NONFINITE mirrors the stop() of :251-253. */
int fgs_score_observed(fgs_ctx *ctx, int score_kind, int abs_flag, double *scores);
/* es.fn$run on the observed scores with all extras
 * (flexgsea_weighted_ks_ :841-908).  scores [n_gene x R] host.  Outputs, each
 * optional (NULL): es [S x R]; max_es_at, le_prop [S x R] (weighted KS only);
 * ragged arrays sharing the CSR offsets, [nnz x R]: running_es_at (1-based gene
 * ids in rank order), running_es_pos, running_es_neg; le_first/le_count
 * [S x R]: the leading edge of set s is running_es_at[offsets[s] + le_first ..
 * + le_count).
 * The weighted-KS observed ES stays in the context (until x, the gene sets or
 * the response change): a later fgs_perm_null / fgs_es_from_scores recomputes
 * every null unit that lands within rounding of it in the reference's own
 * operation order, so the `<=` / `>=` comparisons of flexgsea_calc_sig see the
 * ties the reference sees (label-preserving permutations, small universes).
 * Call it before the null loop, as flexgsea() does (:266-305).  When the scores
 * are the ones fgs_score_observed just returned, fgs_perm_null with the same
 * score kind, abs flag and exponent also copies the observed ES verbatim to
 * the permutations that leave every class label in place; an observed ES of
 * any other provenance (other scores, fgs_set_observed_es) is only used to
 * spot the near-ties, never copied. */
int fgs_observed_es(fgs_ctx *ctx, int es_kind, double p_exponent, const double *scores,
                    int n_gene, int n_response, double *es, double *max_es_at, double *le_prop,
                    int *running_es_at, double *running_es_pos, double *running_es_neg,
                    int *le_first, int *le_count);
/* The same hand-over for an observed ES that was computed elsewhere (e.g. by the R
 * closures of a caller that keeps the observed pass in R): es [S x R] for the
 * ES kind the null loop will use. */
int fgs_set_observed_es(fgs_ctx *ctx, int es_kind, const double *es, int n_sets, int n_response);


/* ---- null loop: flexgsea_perm_sequential / _parallel (:356-472) -------- */
/* Computes es_null [S, R, n_perm] for permutations perm_id0 .. perm_id0 +
 * n_perm - 1.  perm != NULL: [n_sample x n_perm] int, 1-based,
 * y.perm[i,] = y[perm[i,p],] (:384) -- an R-made matrix reproduces R's RNG
 * stream.  perm == NULL: permutations are drawn on the device, Philox4x32-10
 * keyed by (seed, global permutation id), so the result does not depend on how
 * permutations are sharded over GPUs.  The null stays resident on the device
 * for fgs_calc_sig; es_null_out (host, optional) receives a copy. */
int fgs_perm_null(fgs_ctx *ctx, int score_kind, int es_kind, double p_exponent,
                  int abs_flag, const int *perm, unsigned long long seed,
                  long long perm_id0, int n_perm, double *es_null_out);
/* Same, for a user-defined gene.score.fn: the R closure made
 * scores [n_gene, n_response, n_perm] (:393-394); ranking onward runs here. */
int fgs_es_from_scores(fgs_ctx *ctx, int es_kind, double p_exponent, const double *scores,
                       int n_gene, int n_response, int n_perm, double *es_out);
/* The same in tiles, for a gene.score.fn that is slow on the host (limma's
 * moderated t, R/flexgsea_limma.R:15-31): the caller scores block.size
 * permutations, hands the tile over and goes on scoring the next one while the
 * device uploads, ranks and scores this one -- the call only queues the work.
 * Tiles must be given in order from permutation 0 and cover [0, n_perm_total).
 * A helper thread of the context takes the tile -- staging copy, upload and
 * kernel launches -- so the call returns at once; `scores` [n_gene, n_response,
 * n_perm_tile] must stay valid and unmodified until the NEXT call on this
 * context returns (the next tile, fgs_null_finish, or any other entry point:
 * each waits for the tile in flight first).  A failed tile is reported by that
 * next call.  fgs_null_finish waits for the device, reports what the tiles found
 * (FGS_ERR_NONFINITE like :400-402) and optionally copies es_null
 * [S, R, n_perm_total] to the host; the null then is the resident one. */
int fgs_es_from_scores_tile(fgs_ctx *ctx, int es_kind, double p_exponent, const double *scores,
                            int n_gene, int n_response, int n_perm_tile, long long perm_offset,
                            long long n_perm_total);
int fgs_null_finish(fgs_ctx *ctx, double *es_null_out);
/* debug / parity taps for the block most recently processed by fgs_perm_null
 * or fgs_es_from_scores when n_perm * R <= the internal block (tests only):
 * scores [G x C], rank (1-based descending, R/flexgsea.R:929-931) [G x C] */
int fgs_debug_last_scores(fgs_ctx *ctx, double *scores, int *rank, long long n_col);
/* device pointer + shape of the resident null [S, R, P_local] (for NCCL
 * all-gather by the host when es_null is a requested return value) */
int fgs_null_device_ptr(fgs_ctx *ctx, void **dptr, int *n_sets, int *n_response, int *n_perm);
/* load a null made elsewhere (user es.fn, or tests): es_null [S, R, P] host */
int fgs_set_null(fgs_ctx *ctx, const double *es_null, int n_sets, int n_response, int n_perm);

/* ---- significance: flexgsea_calc_sig (:564-629), _simple (:544-547) ---- */
/* One response at a time, on the resident null.  Single-GPU convenience:
 * outputs [S] each (NULL to skip): p, p_high, p_low (simple variant), nes, fdr,
 * fwer; counts for parity checks: p_counts [S x 2] (numerator-1, denominator-1;
 * or the high/low counts), fdr_counts [S x 2] (#{nes cmp nes_i},
 * #{nes_null cmp nes_i}), glob_counts [4] (:501-504). */
int fgs_calc_sig(fgs_ctx *ctx, const double *es, int response, int split_p, int calc_nes,
                 int abs_flag, double *p, double *p_high, double *p_low, double *nes,
                 double *fdr, double *fwer, long long *p_counts, long long *fdr_counts,
                 long long *glob_counts);
/* Staged form for permutations sharded over GPUs (one context per rank).
 * stage1: local per-set partials over this rank's shard: sums [S x 4] =
 * (sum of null >= 0 as hi,lo double-double; sum of null < 0 as hi,lo),
 * counts [S x 6] = (n(null>=0), n(null<0), n(null<=0), #{es<=null & null>=0},
 * #{es>=null & null<=0}, spare) or, simple variant, (#{es<=null}, #{es>=null}).
 * The host all-gathers sums (combined in rank order) and all-reduces counts.
 * stage2: given the combined means, local pooled-FDR counts: null_counts [S]
 * (#{nes_null cmp nes_i} over the shard) and glob [2] (#{nes_null>0},
 * #{nes_null<0}); all-reduced by the host.  finish: S-length epilogue on the
 * device from combined counts. */
int fgs_sig_stage1(fgs_ctx *ctx, const double *es, int response, int split_p, int abs_flag,
                   double *sums, long long *counts);
/* sums_parts [n_parts x S x 4]: every rank's stage-1 sums, in rank order;
 * counts_total [S x 6]: the all-reduced counts; n_perm_total: P over all ranks.
 * Outputs [S]: p, p_high, p_low, nes, fwer; null_counts [S], glob [2] are this
 * rank's pooled-FDR counts (only when calc_nes); fdr_bh [S] is p.adjust(p, 'BH')
 * (:625, only when !calc_nes -- then there is no stage 3). */
int fgs_sig_stage2(fgs_ctx *ctx, const double *es, int response, int split_p, int calc_nes,
                   int abs_flag, const double *sums_parts, int n_parts,
                   const long long *counts_total, long long n_perm_total, double *p,
                   double *p_high, double *p_low, double *nes, double *fwer,
                   double *fdr_bh, long long *null_counts, long long *glob);
int fgs_sig_finish(fgs_ctx *ctx, const double *nes, const long long *null_counts,
                   const long long *glob, int n_sets, int abs_flag, double *fdr,
                   long long *fdr_counts, long long *glob_counts);

/* ---- permutations sharded over GPUs (R/flexgsea.R:422-472, :314-321) ------ */
/* flexgsea_perm_parallel hands blocks of permutations to foreach workers
 * (%dopar%, :432-470) and flexgsea_calc_sig then sees the abind()-ed null.  Here
 * GPU g of N takes the contiguous span [g*P/N, (g+1)*P/N) of the permutations
 * (x, the gene sets and the observed ES are replicated; Philox streams are keyed
 * by the global permutation id, so the null does not depend on N), and the one
 * exchange step happens inside the library over NCCL (NVLink / NVSwitch):
 *   - ONE ncclAllGather of every rank's per-set double-double null sums [S x 4]
 *     with its p-value counts [S x 6] behind them; the sums are combined in rank
 *     order on every rank (=> the NES means are bit-identical for any N), the
 *     counts summed;
 *   - ncclAllReduce (int64) of the pooled-FDR exceedance counts [S + 2] -- the
 *     exact equivalent of all-gathering the null NES (calc_fdr_nes :500-533
 *     only counts), S*8 bytes instead of S*P*8;
 *   - fgs_comm_allgather_null: the null itself, device to device, when the
 *     caller wants es_null back (return_values) or asks for the exchange
 *     north_star names (exchange = 1 below).
 * NCCL is loaded at run time (libnccl.so.2); single-GPU use does not need it.
 *
 * Two ways in.  (1) One process per GPU (torchrun, MPI, ...): rank 0 makes an id,
 * the host program carries its FGS_COMM_ID_BYTES bytes to the other ranks by any
 * means, every rank calls fgs_comm_init on its own context. */
#define FGS_COMM_ID_BYTES 128
int fgs_comm_unique_id(char *id /* [FGS_COMM_ID_BYTES] */);
int fgs_comm_init(fgs_ctx *ctx, const char *id, int n_ranks, int rank);
int fgs_comm_destroy(fgs_ctx *ctx);
/* n_ranks = 1, rank = 0 when no communicator is attached */
int fgs_comm_info(fgs_ctx *ctx, int *n_ranks, int *rank);
/* the span of n_perm permutations rank `rank` of `n_ranks` takes: [*lo, *hi) */
int fgs_shard_span(long long n_perm, int n_ranks, int rank, long long *lo, long long *hi);
/* flexgsea_calc_sig over the sharded null: same arguments and outputs as
 * fgs_calc_sig, every rank receives the full result.  n_perm_total = P over all
 * ranks.  exchange = 0: counts (default); 1: all-gather the null shards first
 * (fgs_comm_allgather_null) and let every rank compute the whole significance
 * locally.  Without a communicator it is fgs_calc_sig. */
int fgs_calc_sig_sharded(fgs_ctx *ctx, const double *es, int response, int split_p,
                         int calc_nes, int abs_flag, long long n_perm_total, int exchange,
                         double *p, double *p_high, double *p_low, double *nes, double *fdr,
                         double *fwer, long long *p_counts, long long *fdr_counts,
                         long long *glob_counts);
/* All-gather the resident null shards (rank r holds fgs_shard_span(n_perm_total,
 * n_ranks, r)) device to device; afterwards every rank's resident null is the
 * full [S, R, n_perm_total].  es_null_out (host, optional) receives a copy. */
int fgs_comm_allgather_null(fgs_ctx *ctx, long long n_perm_total, double *es_null_out);
/* device time (ms) of the most recent significance call on this context: [0] the
 * whole call, [1] the NCCL collectives inside it (includes waiting for the slowest
 * rank), [2] from the start of the most recent null loop to the end of this call
 * (one "step" on the device timeline, host gaps included), [3] collectives issued */
int fgs_last_sig_timings(fgs_ctx *ctx, float *ms /* [4] */);

/* (2) One process, several GPUs -- what flexgsea(parallel = N) binds from R: a
 * group owns one context per device, a communicator made with ncclCommInitAll
 * and fans every call out over one host thread per GPU.  Uploads are replicated;
 * the observed pass runs on every GPU (S*R units) so that each holds the
 * observed ES for the near-tie hand-over of the null loop; the null loop takes
 * the whole permutation range (perm: the full [n_sample x n_perm] matrix or
 * NULL for Philox; es_null_out: the full [S, R, n_perm] host array or NULL, each
 * GPU writes its span); significance is fgs_calc_sig_sharded on every rank,
 * outputs from rank 0.  Arguments are those of the single-context functions. */
typedef struct fgs_group fgs_group;
int fgs_group_create(int n_gpu, const int *devices /* NULL: 0 .. n_gpu-1 */, fgs_group **out);
int fgs_group_destroy(fgs_group *g);
int fgs_group_size(const fgs_group *g);
fgs_ctx *fgs_group_ctx(fgs_group *g, int rank);
int fgs_group_set_x(fgs_group *g, const double *x, int n_sample, int n_gene);
int fgs_group_set_gene_sets(fgs_group *g, const int *offsets, const int *index, int n_sets);
int fgs_group_set_response_s2n(fgs_group *g, const int *y_bin, int n_sample, int n_response);
int fgs_group_set_response_lm(fgs_group *g, const double *design, int n_sample, int n_col,
                              int has_intercept);
int fgs_group_set_option(fgs_group *g, const char *name, long long value);
int fgs_group_score_observed(fgs_group *g, int score_kind, int abs_flag, double *scores);
int fgs_group_observed_es(fgs_group *g, int es_kind, double p_exponent, const double *scores,
                          int n_gene, int n_response, double *es, double *max_es_at,
                          double *le_prop, int *running_es_at, double *running_es_pos,
                          double *running_es_neg, int *le_first, int *le_count);
int fgs_group_perm_null(fgs_group *g, int score_kind, int es_kind, double p_exponent,
                        int abs_flag, const int *perm, unsigned long long seed, int n_perm,
                        double *es_null_out);
int fgs_group_es_from_scores(fgs_group *g, int es_kind, double p_exponent, const double *scores,
                             int n_gene, int n_response, int n_perm, double *es_out);
/* tiles of a user gene.score.fn (fgs_es_from_scores_tile): each goes to the GPU(s)
 * whose span it falls into; fgs_group_null_finish waits for all of them */
int fgs_group_es_from_scores_tile(fgs_group *g, int es_kind, double p_exponent,
                                  const double *scores, int n_gene, int n_response,
                                  int n_perm_tile, long long perm_offset, long long n_perm_total);
int fgs_group_null_finish(fgs_group *g, double *es_null_out);
int fgs_group_calc_sig(fgs_group *g, const double *es, int response, int split_p, int calc_nes,
                       int abs_flag, int exchange, double *p, double *p_high, double *p_low,
                       double *nes, double *fdr, double *fwer, long long *p_counts,
                       long long *fdr_counts, long long *glob_counts);
/* max over the group's GPUs of fgs_last_timings()[4] (null loop) and of
 * fgs_last_sig_timings(): ms[0] null loop, [1] significance, [2] collectives, [3] step */
int fgs_group_last_timings(fgs_group *g, float *ms /* [4] */);

/* ---- gene-set ingest (host only, no device work) ------------------------- */
/* GMT text -> CSR of 1-based gene indices, done once and natively.  Replaces
 * read_gmt (R/flexgsea.R:966-973: name <tab> description <tab> gene ...),
 * filter_gene_sets (:939-958: genes absent from gene_names are dropped, then
 * sets whose remaining size is outside [gs_size_min, gs_size_max]) and the
 * per-set match(set, gene.names) of :279 / :412 (first occurrence wins for a
 * duplicated gene name; duplicates inside a set are kept).  The CSR is what
 * fgs_set_gene_sets takes; set names come back in file order. */
typedef struct fgs_gmt fgs_gmt;
int fgs_gmt_parse(const char *text, size_t len, const char *const *gene_names, int n_gene,
                  int gs_size_min, int gs_size_max, fgs_gmt **out);
int fgs_gmt_read(const char *path, const char *const *gene_names, int n_gene, int gs_size_min,
                 int gs_size_max, fgs_gmt **out);
int fgs_gmt_n_sets(const fgs_gmt *g);
long long fgs_gmt_nnz(const fgs_gmt *g);
const int *fgs_gmt_offsets(const fgs_gmt *g); /* [n_sets + 1] */
const int *fgs_gmt_index(const fgs_gmt *g);   /* [nnz], 1-based */
const char *fgs_gmt_set_name(const fgs_gmt *g, int k);
/* the counts filter_gene_sets reports: members before / after the gene filter, sets before the size filter */
int fgs_gmt_stats(const fgs_gmt *g, long long *genes_before, long long *genes_after, int *sets_before);
void fgs_gmt_free(fgs_gmt *g);

/* ---- instrumentation --------------------------------------------------- */
/* number of kernels this library launched on ctx since creation */
int fgs_launch_count(fgs_ctx *ctx, long long *count);
/* device time (ms, CUDA events on the context's stream) of the stages of the
 * most recent fgs_perm_null / fgs_es_from_scores: [0] labels+perm, [1] score,
 * [2] rank/sort, [3] ES, [4] total; kernel launches per stage in launches[5] */
int fgs_last_timings(fgs_ctx *ctx, float *ms, long long *launches);
/* enable per-stage event timing (adds stream synchronisation points) */
int fgs_set_profiling(fgs_ctx *ctx, int enabled);
/* Progress / interrupt hook for the block loops of fgs_perm_null and
 * fgs_es_from_scores: called on the calling thread once per block of score columns
 * with the number of columns whose work is finished and the total; a non-zero return
 * stops the call with FGS_ERR_CANCELLED (the null is then incomplete).  With a hook
 * set the loop keeps at most two blocks in flight, so the answer to an interrupt
 * comes within tens of milliseconds.  This is where an R binding polls
 * R_CheckUserInterrupt (through R_ToplevelExec, so nothing longjmps over CUDA
 * state); the reference is interruptible between its interpreted steps
 * (R/flexgsea.R:372-417).  cb == NULL removes the hook. */
typedef int (*fgs_progress_fn)(void *user, long long columns_done, long long columns_total);
int fgs_set_progress(fgs_ctx *ctx, fgs_progress_fn cb, void *user);
/* tuning / test hooks: "perm_block" (permutations per block), "force_generic_es"
 * (route every set through the duplicate-safe O(m^2) kernel), "force_sort64"
 * (rank with the 64-bit global-memory sort), "s2n_mode" (1, default: s2n scores
 * of the null loop by the FP64 tensor-core contraction with rank certification;
 * 0: the bit-faithful three-pass kernel), "cert_cap" (certification list size),
 * "h2d_threads" (8, default: host threads that stage large uploads of user score
 * columns and the block-wise return of es.null through pinned memory; 0: plain
 * copies),
 * "es_exact_small" (1, default: problems that cost well under a millisecond that
 * way compute every weighted-KS unit in the reference's operation order),
 * "nperm_total" (0, default: this call's n_perm; a caller that shards the
 * permutations over contexts sets the total, so that routing decisions such as
 * es_exact_small do not depend on the shard -- the groups do it themselves),
 * "tail_split" (1, default: the ES kernels split the columns of a partial last
 * round over the idle SMs),
 * "bulk_stage" (1, default: the ES kernels stage a column's weights / scores in
 * shared memory with one cp.async.bulk copy; 0: a load / store loop),
 * "arrange_members" (1, default: member blocks of the mean / maxmean kernels
 * dealt out by shared-memory bank),
 * "dedup_labels" (1, default: with few samples -- at most 32, and fewer distinct
 * class labellings than half the permutations -- fgs_perm_null computes every
 * distinct labelling once and copies its column to the permutations that share
 * it; such problems also take the all-exact path up to a far larger size) */
int fgs_set_option(fgs_ctx *ctx, const char *name, long long value);
/* read-only counters of the most recent call: "dedup_labellings" (distinct
 * labellings the last fgs_perm_null computed, 0 when it did not deduplicate),
 * "fallback_blocks" (blocks of the last fgs_perm_null whose certification lists
 * overflowed and were rescored by the bit-faithful kernel), "duplicated_rows"
 * (1 when the resident x holds genes with bit-identical rows) */
int fgs_get_stat(fgs_ctx *ctx, const char *name, long long *value);

#ifdef __cplusplus
}
#endif
#endif
