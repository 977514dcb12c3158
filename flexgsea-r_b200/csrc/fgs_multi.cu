// Permutations sharded over GPUs: the exchange step of the significance stage over NCCL, and the
// single-process group of contexts that flexgsea(parallel = N) binds from R.
//
// Replaces flexgsea_perm_parallel's foreach(%dopar%) over permutation blocks and the abind() of
// their results (R/flexgsea.R:432-470) together with the per-response sig.fun call on the
// combined null (:314-321).  See include/flexgsea_b200.h ("permutations sharded over GPUs").
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already holds a copy
// (torch's bundled one, say) gets that copy, and single-GPU use needs none.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>

#include "fgs_common.cuh"

namespace fgs {

int calc_sig_core(fgs_ctx *c, const double *es, int response, int split_p, int calc_nes, int abs_flag,
                  long long P_total, bool sharded, double *p, double *p_high, double *p_low, double *nes,
                  double *fdr, double *fwer, long long *p_counts, long long *fdr_counts,
                  long long *glob_counts);

namespace {

struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool tried = false;
};
NcclApi g_nccl;

const NcclApi *nccl() {
  static std::once_flag once;
  std::call_once(once, [] {
    g_nccl.tried = true;
    // FLEXGSEA_NCCL_LIB: an explicit copy (a process that will load a NEWER NCCL later -- torch's
    // bundled one, say -- must either load that first or name it here: both answer to the soname
    // libnccl.so.2 and the first one loaded serves everybody)
    // The exchanges are S-length arrays (hundreds of KB): NVLS multicast buys nothing for them, and with it
    // enabled an nvidia-smi query against a GPU of the communicator stalled that rank for milliseconds
    // (profiles/README.md).  Only a default: an NCCL_NVLS_ENABLE the user set wins, and a copy of NCCL that
    // was initialised earlier in the process keeps its own setting.
    setenv("NCCL_NVLS_ENABLE", "0", 0);
    const char *names[] = {getenv("FLEXGSEA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      if (!nm || !*nm) continue;
      g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return;
    bool ok = true;
    auto sym = [&](const char *nm) -> void * {
      void *p = dlsym(g_nccl.handle, nm);
      if (!p) ok = false;
      return p;
    };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast))sym("ncclBroadcast");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    if (!ok) { dlclose(g_nccl.handle); g_nccl.handle = nullptr; }
  });
  if (!g_nccl.handle) {
    set_error("NCCL is not available: libnccl.so.2 could not be loaded (multi-GPU runs need it on the library path)");
    return nullptr;
  }
  return &g_nccl;
}

#define FGS_NCCL(api, call)                                                                      \
  do {                                                                                           \
    ncclResult_t r__ = (call);                                                                   \
    if (r__ != ncclSuccess) {                                                                    \
      fgs::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, (api)->GetErrorString(r__)); \
      return FGS_ERR_COMM;                                                                       \
    }                                                                                            \
  } while (0)

static_assert(sizeof(ncclUniqueId) == FGS_COMM_ID_BYTES, "FGS_COMM_ID_BYTES must match ncclUniqueId");

void span_of(long long P, int n, int r, long long *lo, long long *hi) {
  *lo = (P * r) / n;
  *hi = (P * (r + 1)) / n;
}

}  // namespace

int comm_allgather_f64(fgs_ctx *c, const double *send, double *recv, size_t count) {
  if (!c->comm || c->comm_size <= 1) return FGS_OK;
  const NcclApi *a = nccl();
  if (!a) return FGS_ERR_COMM;
  FGS_NCCL(a, a->AllGather(send, recv, count, ncclDouble, (ncclComm_t)c->comm, c->stream));
  return FGS_OK;
}
int comm_allreduce_i64(fgs_ctx *c, long long *buf, size_t count) {
  if (!c->comm || c->comm_size <= 1) return FGS_OK;
  const NcclApi *a = nccl();
  if (!a) return FGS_ERR_COMM;
  FGS_NCCL(a, a->AllReduce(buf, buf, count, ncclInt64, ncclSum, (ncclComm_t)c->comm, c->stream));
  return FGS_OK;
}

}  // namespace fgs

using namespace fgs;

// a group call runs one host thread per GPU; the first failure's message is handed to the caller
struct fgs_group {
  std::vector<fgs_ctx *> ctx;
  std::vector<std::string> err;
  template <typename F>
  int each(F f) {
    const int n = (int)ctx.size();
    std::vector<int> rc(n, FGS_OK);
    auto run = [&](int r) {
      rc[r] = f(r, ctx[r]);
      if (rc[r] != FGS_OK) err[r] = fgs_last_error();
    };
    std::vector<std::thread> th;
    for (int r = 1; r < n; r++) th.emplace_back(run, r);
    run(0);
    for (auto &t : th) t.join();
    for (int r = 0; r < n; r++)
      if (rc[r] != FGS_OK) {
        set_error("GPU %d of %d: %s", r, n, err[r].c_str());
        return rc[r];
      }
    return FGS_OK;
  }
};

extern "C" {

int fgs_comm_unique_id(char *id) {
  if (!id) return FGS_ERR_ARG;
  const NcclApi *a = nccl();
  if (!a) return FGS_ERR_COMM;
  ncclUniqueId u;
  FGS_NCCL(a, a->GetUniqueId(&u));
  std::memcpy(id, &u, sizeof(u));
  return FGS_OK;
}

int fgs_comm_init(fgs_ctx *c, const char *id, int n_ranks, int rank) {
  if (!c || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_error("fgs_comm_init: bad argument"); return FGS_ERR_ARG; }
  FGS_TRY(fgs_comm_destroy(c));
  const NcclApi *a = nccl();
  if (!a) return FGS_ERR_COMM;
  FGS_CUDA(cudaSetDevice(c->device));
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  FGS_NCCL(a, a->CommInitRank(&comm, n_ranks, u, rank));
  c->comm = comm; c->comm_size = n_ranks; c->comm_rank = rank; c->comm_owned = true;
  return FGS_OK;
}

int fgs_comm_destroy(fgs_ctx *c) {
  if (!c) return FGS_ERR_ARG;
  if (c->comm && c->comm_owned) {
    const NcclApi *a = nccl();
    if (a) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); a->CommDestroy((ncclComm_t)c->comm); }
  }
  c->comm = nullptr; c->comm_size = 1; c->comm_rank = 0; c->comm_owned = true;
  return FGS_OK;
}

int fgs_comm_info(fgs_ctx *c, int *n_ranks, int *rank) {
  if (!c) return FGS_ERR_ARG;
  if (n_ranks) *n_ranks = c->comm ? c->comm_size : 1;
  if (rank) *rank = c->comm ? c->comm_rank : 0;
  return FGS_OK;
}

int fgs_shard_span(long long n_perm, int n_ranks, int rank, long long *lo, long long *hi) {
  if (n_perm < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks || !lo || !hi) return FGS_ERR_ARG;
  span_of(n_perm, n_ranks, rank, lo, hi);
  return FGS_OK;
}

int fgs_comm_allgather_null(fgs_ctx *c, long long n_perm_total, double *es_null_out) {
  if (!c || n_perm_total < 0) return FGS_ERR_ARG;
  FGS_CUDA(cudaSetDevice(c->device));
  if (c->null_S <= 0 || c->null_R <= 0) { set_error("no resident null to gather"); return FGS_ERR_STATE; }
  const int N = c->comm ? c->comm_size : 1, me = c->comm ? c->comm_rank : 0;
  const size_t unit = (size_t)c->null_S * c->null_R;
  long long lo, hi;
  span_of(n_perm_total, N, me, &lo, &hi);
  if (hi - lo != c->null_P) {
    set_error("resident null holds %d permutations, the span of rank %d of %d over %lld is %lld", c->null_P, me, N,
              n_perm_total, hi - lo);
    return FGS_ERR_STATE;
  }
  if (N > 1) {
    const NcclApi *a = nccl();
    if (!a) return FGS_ERR_COMM;
    FGS_TRY(c->null_full.ensure(std::max<size_t>(unit * (size_t)n_perm_total, 1)));
    // spans differ by at most one permutation: an all-gather "v" as one broadcast per rank, grouped
    FGS_NCCL(a, a->GroupStart());
    for (int r = 0; r < N; r++) {
      long long rlo, rhi;
      span_of(n_perm_total, N, r, &rlo, &rhi);
      if (rhi == rlo) continue;
      ncclResult_t rr = a->Broadcast(c->es_null.p, c->null_full.p + unit * (size_t)rlo, unit * (size_t)(rhi - rlo),
                                     ncclDouble, r, (ncclComm_t)c->comm, c->stream);
      if (rr != ncclSuccess) { a->GroupEnd(); set_error("ncclBroadcast failed: %s", a->GetErrorString(rr)); return FGS_ERR_COMM; }
    }
    FGS_NCCL(a, a->GroupEnd());
    std::swap(c->es_null, c->null_full);
    c->null_P = (int)n_perm_total;
  }
  if (es_null_out) {
    FGS_CUDA(cudaMemcpyAsync(es_null_out, c->es_null.p, unit * (size_t)n_perm_total * sizeof(double),
                             cudaMemcpyDeviceToHost, c->stream));
  }
  FGS_CUDA(cudaStreamSynchronize(c->stream));
  return FGS_OK;
}

int fgs_calc_sig_sharded(fgs_ctx *c, const double *es, int response, int split_p, int calc_nes,
                         int abs_flag, long long n_perm_total, int exchange, double *p, double *p_high,
                         double *p_low, double *nes, double *fdr, double *fwer, long long *p_counts,
                         long long *fdr_counts, long long *glob_counts) {
  if (!c || !es) return FGS_ERR_ARG;
  FGS_CUDA(cudaSetDevice(c->device));
  if (c->null_S <= 0 || c->null_R <= 0) { set_error("no resident null: run fgs_perm_null / fgs_es_from_scores / fgs_set_null first"); return FGS_ERR_STATE; }
  if (response < 0 || response >= c->null_R) { set_error("response %d out of range (%d)", response, c->null_R); return FGS_ERR_ARG; }
  const bool sharded = c->comm && c->comm_size > 1;
  if (!sharded) n_perm_total = c->null_P;
  if (n_perm_total == 0) return FGS_OK;  // R/flexgsea.R:575-577
  if (sharded && exchange == 1) {
    // the exchange north_star names: every rank receives the whole null and computes locally
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float t_gather = 0.f;
    if (c->null_P != n_perm_total) {
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0, c->stream);
      int rc = fgs_comm_allgather_null(c, n_perm_total, nullptr);
      cudaEventRecord(e1, c->stream);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&t_gather, e0, e1);
      cudaEventDestroy(e0); cudaEventDestroy(e1);
      if (rc != FGS_OK) return rc;
    }
    FGS_TRY(calc_sig_core(c, es, response, split_p, calc_nes, abs_flag, n_perm_total, false, p, p_high, p_low, nes,
                          fdr, fwer, p_counts, fdr_counts, glob_counts));
    c->last_sig_ms[0] += t_gather; c->last_sig_ms[1] = t_gather; c->last_sig_ms[3] = t_gather > 0.f ? 1.f : 0.f;
    return FGS_OK;
  }
  return calc_sig_core(c, es, response, split_p, calc_nes, abs_flag, n_perm_total, sharded, p, p_high, p_low, nes,
                       fdr, fwer, p_counts, fdr_counts, glob_counts);
}

// ---- single-process group -------------------------------------------------------------------
int fgs_group_create(int n_gpu, const int *devices, fgs_group **out) {
  if (!out || n_gpu < 1) { set_error("fgs_group_create: bad argument"); return FGS_ERR_ARG; }
  *out = nullptr;
  int avail = 0;
  FGS_TRY(fgs_device_count(&avail));
  std::vector<int> dev(n_gpu);
  for (int i = 0; i < n_gpu; i++) {
    dev[i] = devices ? devices[i] : i;
    if (dev[i] < 0 || dev[i] >= avail) { set_error("parallel = %d GPUs asked for, %d visible", n_gpu, avail); return FGS_ERR_ARG; }
  }
  fgs_group *g = new fgs_group();
  g->err.resize(n_gpu);
  for (int i = 0; i < n_gpu; i++) {
    fgs_ctx *c = nullptr;
    const int rc = fgs_create(dev[i], &c);
    if (rc != FGS_OK) { fgs_group_destroy(g); return rc; }
    g->ctx.push_back(c);
  }
  if (n_gpu > 1) {
    const NcclApi *a = nccl();
    if (!a) { fgs_group_destroy(g); return FGS_ERR_COMM; }
    std::vector<ncclComm_t> comms(n_gpu);
    ncclResult_t r = a->CommInitAll(comms.data(), n_gpu, dev.data());
    if (r != ncclSuccess) {
      set_error("ncclCommInitAll over %d GPUs failed: %s", n_gpu, a->GetErrorString(r));
      fgs_group_destroy(g);
      return FGS_ERR_COMM;
    }
    for (int i = 0; i < n_gpu; i++) {
      g->ctx[i]->comm = comms[i]; g->ctx[i]->comm_size = n_gpu; g->ctx[i]->comm_rank = i; g->ctx[i]->comm_owned = true;
    }
  }
  *out = g;
  return FGS_OK;
}

int fgs_group_destroy(fgs_group *g) {
  if (!g) return FGS_OK;
  for (fgs_ctx *c : g->ctx) fgs_destroy(c);  // destroys the rank's communicator too
  delete g;
  return FGS_OK;
}

int fgs_group_size(const fgs_group *g) { return g ? (int)g->ctx.size() : 0; }
fgs_ctx *fgs_group_ctx(fgs_group *g, int rank) {
  return (g && rank >= 0 && rank < (int)g->ctx.size()) ? g->ctx[rank] : nullptr;
}

int fgs_group_set_x(fgs_group *g, const double *x, int n_sample, int n_gene) {
  if (!g) return FGS_ERR_ARG;
  return g->each([&](int, fgs_ctx *c) { return fgs_set_x(c, x, n_sample, n_gene); });
}
int fgs_group_set_gene_sets(fgs_group *g, const int *offsets, const int *index, int n_sets) {
  if (!g) return FGS_ERR_ARG;
  return g->each([&](int, fgs_ctx *c) { return fgs_set_gene_sets(c, offsets, index, n_sets); });
}
int fgs_group_set_response_s2n(fgs_group *g, const int *y_bin, int n_sample, int n_response) {
  if (!g) return FGS_ERR_ARG;
  return g->each([&](int, fgs_ctx *c) { return fgs_set_response_s2n(c, y_bin, n_sample, n_response); });
}
int fgs_group_set_response_lm(fgs_group *g, const double *design, int n_sample, int n_col, int has_intercept) {
  if (!g) return FGS_ERR_ARG;
  return g->each([&](int, fgs_ctx *c) { return fgs_set_response_lm(c, design, n_sample, n_col, has_intercept); });
}
int fgs_group_set_option(fgs_group *g, const char *name, long long value) {
  if (!g) return FGS_ERR_ARG;
  for (fgs_ctx *c : g->ctx) FGS_TRY(fgs_set_option(c, name, value));
  return FGS_OK;
}

int fgs_group_score_observed(fgs_group *g, int score_kind, int abs_flag, double *scores) {
  if (!g) return FGS_ERR_ARG;
  return fgs_score_observed(g->ctx[0], score_kind, abs_flag, scores);
}

int fgs_group_observed_es(fgs_group *g, int es_kind, double p_exponent, const double *scores, int n_gene,
                          int n_response, double *es, double *max_es_at, double *le_prop, int *running_es_at,
                          double *running_es_pos, double *running_es_neg, int *le_first, int *le_count) {
  if (!g) return FGS_ERR_ARG;
  // every GPU runs the observed pass (S * R units) and keeps the observed ES for the near-tie hand-over
  // of its share of the null loop; the caller's arrays are filled by rank 0, the others only need `es`
  const size_t SR = (size_t)std::max(g->ctx[0]->S, 0) * (size_t)std::max(n_response, 0);
  std::vector<std::vector<double>> tmp(g->ctx.size());
  return g->each([&](int r, fgs_ctx *c) {
    if (r == 0)
      return fgs_observed_es(c, es_kind, p_exponent, scores, n_gene, n_response, es, max_es_at, le_prop,
                             running_es_at, running_es_pos, running_es_neg, le_first, le_count);
    tmp[r].resize(std::max<size_t>(SR, 1));
    return fgs_observed_es(c, es_kind, p_exponent, scores, n_gene, n_response, tmp[r].data(), nullptr, nullptr,
                           nullptr, nullptr, nullptr, nullptr, nullptr);
  });
}

int fgs_group_perm_null(fgs_group *g, int score_kind, int es_kind, double p_exponent, int abs_flag,
                        const int *perm, unsigned long long seed, int n_perm, double *es_null_out) {
  if (!g || n_perm < 0) return FGS_ERR_ARG;
  const int N = (int)g->ctx.size();
  return g->each([&](int r, fgs_ctx *c) {
    long long lo, hi;
    span_of(n_perm, N, r, &lo, &hi);
    int R = 0;
    FGS_TRY(fgs_n_response(c, score_kind, &R));
    c->opt_nperm_total = n_perm;
    const int *pm = perm ? perm + (size_t)lo * c->n : nullptr;
    double *out = es_null_out ? es_null_out + (size_t)lo * R * c->S : nullptr;
    return fgs_perm_null(c, score_kind, es_kind, p_exponent, abs_flag, pm, seed, lo, (int)(hi - lo), out);
  });
}

int fgs_group_es_from_scores(fgs_group *g, int es_kind, double p_exponent, const double *scores, int n_gene,
                             int n_response, int n_perm, double *es_out) {
  if (!g || n_perm < 0 || !scores) return FGS_ERR_ARG;
  const int N = (int)g->ctx.size();
  return g->each([&](int r, fgs_ctx *c) {
    long long lo, hi;
    span_of(n_perm, N, r, &lo, &hi);
    c->opt_nperm_total = n_perm;
    const double *sc = scores + (size_t)lo * n_response * n_gene;
    double *out = es_out ? es_out + (size_t)lo * n_response * c->S : nullptr;
    return fgs_es_from_scores(c, es_kind, p_exponent, sc, n_gene, n_response, (int)(hi - lo), out);
  });
}

// A tile of user scores goes to the GPU(s) whose span of the permutations it falls into; the
// calls only queue work (fgs_es_from_scores_tile), so the caller can go on scoring the next tile.
int fgs_group_es_from_scores_tile(fgs_group *g, int es_kind, double p_exponent, const double *scores, int n_gene,
                                  int n_response, int n_perm_tile, long long perm_offset, long long n_perm_total) {
  if (!g || !scores || n_perm_tile < 0 || perm_offset < 0 || perm_offset + n_perm_tile > n_perm_total) return FGS_ERR_ARG;
  const int N = (int)g->ctx.size();
  // the previous tile's buffer is the caller's again when this call returns: every GPU is through with it
  for (int r = 0; r < N; r++) FGS_TRY(settle_tile(g->ctx[r]));
  for (int r = 0; r < N; r++) {
    long long lo, hi;
    span_of(n_perm_total, N, r, &lo, &hi);
    const long long a = std::max(lo, perm_offset), b = std::min(hi, perm_offset + n_perm_tile);
    if (a >= b) continue;
    g->ctx[r]->opt_nperm_total = n_perm_total;
    const int rc = fgs_es_from_scores_tile(g->ctx[r], es_kind, p_exponent,
                                           scores + (size_t)(a - perm_offset) * n_response * n_gene, n_gene,
                                           n_response, (int)(b - a), a - lo, hi - lo);
    if (rc != FGS_OK) { g->err[r] = fgs_last_error(); set_error("GPU %d of %d: %s", r, N, g->err[r].c_str()); return rc; }
  }
  return FGS_OK;
}

int fgs_group_null_finish(fgs_group *g, double *es_null_out) {
  if (!g) return FGS_ERR_ARG;
  long long before = 0;
  std::vector<long long> off(g->ctx.size());
  for (size_t r = 0; r < g->ctx.size(); r++) { off[r] = before; before += g->ctx[r]->null_P; }
  return g->each([&](int r, fgs_ctx *c) {
    if (c->null_S <= 0) return (int)FGS_OK;   // a rank without permutations
    double *out = es_null_out ? es_null_out + (size_t)off[r] * c->null_R * c->null_S : nullptr;
    return fgs_null_finish(c, out);
  });
}

int fgs_group_calc_sig(fgs_group *g, const double *es, int response, int split_p, int calc_nes, int abs_flag,
                       int exchange, double *p, double *p_high, double *p_low, double *nes, double *fdr,
                       double *fwer, long long *p_counts, long long *fdr_counts, long long *glob_counts) {
  if (!g || !es) return FGS_ERR_ARG;
  long long P_total = 0;
  for (fgs_ctx *c : g->ctx) P_total += c->null_P;
  const int S = g->ctx[0]->null_S;
  if (S <= 0) { set_error("no resident null: run fgs_group_perm_null first"); return FGS_ERR_STATE; }
  std::vector<std::vector<double>> tmp(g->ctx.size());
  return g->each([&](int r, fgs_ctx *c) {
    if (r == 0)
      return fgs_calc_sig_sharded(c, es, response, split_p, calc_nes, abs_flag, P_total, exchange, p, p_high,
                                  p_low, nes, fdr, fwer, p_counts, fdr_counts, glob_counts);
    // the other ranks take part in the collectives; their copies of the result are dropped
    tmp[r].assign((size_t)S * 6, 0.0);
    double *t = tmp[r].data();
    std::vector<long long> fc((size_t)S * 2), gc(4);
    return fgs_calc_sig_sharded(c, es, response, split_p, calc_nes, abs_flag, P_total, exchange, t, t + S,
                                t + 2 * (size_t)S, t + 3 * (size_t)S, t + 4 * (size_t)S, t + 5 * (size_t)S, nullptr,
                                fc.data(), gc.data());
  });
}

int fgs_group_last_timings(fgs_group *g, float *ms) {
  if (!g || !ms) return FGS_ERR_ARG;
  for (int i = 0; i < 4; i++) ms[i] = 0.f;
  for (fgs_ctx *c : g->ctx) {
    ms[0] = std::max(ms[0], c->last_ms[4]);
    ms[1] = std::max(ms[1], c->last_sig_ms[0]);
    ms[2] = std::max(ms[2], c->last_sig_ms[1]);
    ms[3] = std::max(ms[3], c->last_sig_ms[2]);
  }
  return FGS_OK;
}

}  // extern "C"
