// Rcpp shim a maintainer of flexgsea adds next to src/ggsea.cpp.  Thin .Call
// wrappers over the C ABI of libflexgsea_b200.so (include/flexgsea_b200.h); the
// R objects stay R-owned and read-only, results are allocated on the R heap
// here, device state lives in an external pointer with a finalizer.
// NOT compiled in this repository's image (no R / Rcpp headers); see
// INTEGRATION.md.  Link with: PKG_LIBS = -lflexgsea_b200 (src/Makevars).
#include <Rcpp.h>
#include "flexgsea_b200.h"
using namespace Rcpp;

static void ctx_finalizer(SEXP p) {
  fgs_ctx *c = static_cast<fgs_ctx *>(R_ExternalPtrAddr(p));
  if (c) { fgs_destroy(c); R_ClearExternalPtr(p); }
}
static fgs_ctx *ctx_of(SEXP p) {
  fgs_ctx *c = static_cast<fgs_ctx *>(R_ExternalPtrAddr(p));
  if (!c) stop("flexgsea: GPU context was released");
  return c;
}
static void check(int rc) {
  if (rc == FGS_OK) return;
  // never Rf_error across frames that own CUDA resources: everything device side
  // is behind the C ABI and already unwound when we get the code back
  if (rc == FGS_ERR_NONFINITE) stop("Gene scoring function returned NA or Inf.");  // R/flexgsea.R:401
  if (rc == FGS_ERR_CANCELLED) { Rf_onintr(); return; }  // the user interrupt seen by progress_hook_, raised here
  stop("flexgsea_b200: %s", fgs_last_error());
}

// [[Rcpp::export]]
SEXP fgs_create_(int device) {
  fgs_ctx *c = nullptr;
  check(fgs_create(device, &c));
  SEXP p = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(p, ctx_finalizer, TRUE);
  UNPROTECT(1);
  return p;
}

// replaces s2n_C(x, y) (src/ggsea.cpp:6, R/RcppExports.R:4-6) bit for bit
// [[Rcpp::export]]
NumericMatrix s2n_gpu(SEXP ctx, NumericMatrix x, LogicalMatrix y) {
  NumericMatrix coef(x.ncol(), y.ncol());
  check(fgs_s2n(ctx_of(ctx), x.begin(), x.nrow(), x.ncol(), LOGICAL(y), y.nrow(), y.ncol(), coef.begin()));
  return coef;
}

// [[Rcpp::export]]
void fgs_upload_(SEXP ctx, NumericMatrix x, IntegerVector offsets, IntegerVector index) {
  check(fgs_set_x(ctx_of(ctx), x.begin(), x.nrow(), x.ncol()));
  check(fgs_set_gene_sets(ctx_of(ctx), offsets.begin(), index.begin(), offsets.size() - 1));
}
// gene.sets given as a GMT file: read_gmt + filter_gene_sets + match(set, gene.names)
// (R/flexgsea.R:966-973, :939-958, :279/:412) natively and once
// [[Rcpp::export]]
List fgs_gmt_read_(std::string path, CharacterVector gene_names, int gs_size_min, int gs_size_max) {
  std::vector<const char *> names(gene_names.size());
  for (R_xlen_t i = 0; i < gene_names.size(); i++) names[i] = CHAR(STRING_ELT(gene_names, i));
  fgs_gmt *g = nullptr;
  check(fgs_gmt_read(path.c_str(), names.data(), (int)names.size(), gs_size_min, gs_size_max, &g));
  const int S = fgs_gmt_n_sets(g);
  const long long nnz = fgs_gmt_nnz(g);
  IntegerVector offsets(fgs_gmt_offsets(g), fgs_gmt_offsets(g) + S + 1);
  IntegerVector index(fgs_gmt_index(g), fgs_gmt_index(g) + nnz);
  CharacterVector set_names(S);
  for (int k = 0; k < S; k++) set_names[k] = fgs_gmt_set_name(g, k);
  long long gb = 0, ga = 0; int sb = 0;
  fgs_gmt_stats(g, &gb, &ga, &sb);
  fgs_gmt_free(g);
  return List::create(_["names"] = set_names, _["offsets"] = offsets, _["index"] = index,
                      _["genes_before"] = (double)gb, _["genes_after"] = (double)ga, _["sets_before"] = sb);
}

// [[Rcpp::export]]
void fgs_response_s2n_(SEXP ctx, LogicalMatrix y_bin) {
  check(fgs_set_response_s2n(ctx_of(ctx), LOGICAL(y_bin), y_bin.nrow(), y_bin.ncol()));
}
// [[Rcpp::export]]
void fgs_response_lm_(SEXP ctx, NumericMatrix design, bool has_intercept) {
  check(fgs_set_response_lm(ctx_of(ctx), design.begin(), design.nrow(), design.ncol(), has_intercept));
}

// es.null <- perm.fun(...) (R/flexgsea.R:305): perm is matrix(sample.int(n) per column), made in R so the
// null is the one the reference's sequential loop draws under set.seed(); returns [S, R, P] like :363-368
// Interrupt polling between blocks of the null loop: R_CheckUserInterrupt longjmps, so it runs
// inside R_ToplevelExec and the answer goes back to the library as "stop" (FGS_ERR_CANCELLED),
// which unwinds normally and releases nothing it should keep.
static void poll_interrupt_(void *) { R_CheckUserInterrupt(); }
static int progress_hook_(void *, long long, long long) {
  return R_ToplevelExec(poll_interrupt_, nullptr) == FALSE;
}

// [[Rcpp::export]]
void fgs_interruptible_(SEXP ctx, bool on) {
  check(fgs_set_progress(ctx_of(ctx), on ? progress_hook_ : nullptr, nullptr));
}

// [[Rcpp::export]]
NumericVector fgs_perm_null_(SEXP ctx, int score_kind, int es_kind, bool abs, IntegerMatrix perm,
                             int n_sets, int n_response, bool want_null) {
  const int P = perm.ncol();
  NumericVector out(want_null ? (R_xlen_t)n_sets * n_response * P : 0);
  check(fgs_perm_null(ctx_of(ctx), score_kind, es_kind, 1.0, abs, perm.begin(), 0ull, 0, P,
                      want_null ? out.begin() : nullptr));
  if (want_null) out.attr("dim") = IntegerVector::create(n_sets, n_response, P);
  return out;
}

// user-defined gene.score.fn: R made scores[G, R, P] (:393-394); ranking onward on the device
// [[Rcpp::export]]
NumericVector fgs_es_from_scores_(SEXP ctx, int es_kind, NumericVector scores, int n_sets) {
  IntegerVector d = scores.attr("dim");
  NumericVector out((R_xlen_t)n_sets * d[1] * d[2]);
  check(fgs_es_from_scores(ctx_of(ctx), es_kind, 1.0, scores.begin(), d[0], d[1], d[2], out.begin()));
  out.attr("dim") = IntegerVector::create(n_sets, d[1], d[2]);
  return out;
}

// sig.fun(es[, r], es.null[, r, ]) on the resident null (:314-330); columns as flexgsea_calc_sig's tibble
// [[Rcpp::export]]
DataFrame fgs_calc_sig_(SEXP ctx, NumericVector es, int response, bool split_p, bool calc_nes, bool abs) {
  const int S = es.size();
  NumericVector p(S), ph(S), pl(S), nes(S), fdr(S), fwer(S);
  check(fgs_calc_sig(ctx_of(ctx), es.begin(), response, split_p, calc_nes, abs, p.begin(), ph.begin(),
                     pl.begin(), nes.begin(), fdr.begin(), fwer.begin(), nullptr, nullptr, nullptr));
  if (split_p && calc_nes)
    return DataFrame::create(_["es"] = es, _["p"] = p, _["nes"] = nes, _["fdr"] = fdr, _["fwer"] = fwer);
  if (abs) return DataFrame::create(_["es"] = es, _["p.high"] = ph, _["p"] = p, _["fdr"] = fdr, _["fwer"] = fwer);
  return DataFrame::create(_["es"] = es, _["p.high"] = ph, _["p.low"] = pl, _["p"] = p, _["fdr"] = fdr,
                           _["fwer"] = fwer);
}
