// .Call entry points a maintainer of flexgsea adds next to src/ggsea.cpp: thin wrappers over
// the C ABI of libflexgsea_b200.so (include/flexgsea_b200.h).  R objects stay R-owned and read
// only, results are allocated on the R heap here, device state lives in an external pointer
// with a finalizer.  The registration table at the end replaces the generated one of
// src/RcppExports.cpp:21-29 and KEEPS `_flexgsea_s2n_C` with arity 2 (R/RcppExports.R:4-6 and
// flexgsea_s2n, R/flexgsea.R:712, call it unchanged).
//
// Written against R's C API plus Rcpp's BEGIN_RCPP / END_RCPP / Rcpp::stop (C++ exceptions
// become R errors only after every C++ frame has unwound; nothing here calls Rf_error while a
// C++ object with a destructor is alive).  R is not installed in this repository's image: the
// file is compiled and driven by tests/test_r_shim.py against r-package/standin/Rcpp.h, a small
// functional stand-in of that API, so arity and type errors surface without R.
// Build inside the package: src/Makevars adds -I<repo>/include and links -lflexgsea_b200.
#include <Rcpp.h>

#include <cstring>
#include <string>
#include <vector>

#include "flexgsea_b200.h"

namespace {

// ---- errors ---------------------------------------------------------------------------------
void check(int rc) {
  if (rc == FGS_OK) return;
  if (rc == FGS_ERR_NONFINITE)  // the reference's own message, R/flexgsea.R:251-253, :400-402
    Rcpp::stop("Gene scoring function returned NA or Inf. Check that all genes in x have a positive variance.");
  if (rc == FGS_ERR_CANCELLED) Rcpp::stop("flexgsea: interrupted");
  Rcpp::stop(std::string("flexgsea_b200: ") + fgs_last_error());
}

// ---- argument helpers -------------------------------------------------------------------------
int dim_of(SEXP x, int k) {
  SEXP d = Rf_getAttrib(x, R_DimSymbol);
  if (Rf_isNull(d) || Rf_length(d) <= k) Rcpp::stop("flexgsea: expected an array with at least %d dimensions", k + 1);
  return INTEGER(d)[k];
}
const double *real_arg(SEXP x, const char *what) {
  if (TYPEOF(x) != REALSXP) Rcpp::stop("flexgsea: %s must be a double vector / matrix", what);
  return REAL(x);
}
const int *int_arg(SEXP x, const char *what) {
  if (TYPEOF(x) != INTSXP && TYPEOF(x) != LGLSXP) Rcpp::stop("flexgsea: %s must be an integer / logical vector", what);
  return TYPEOF(x) == LGLSXP ? LOGICAL(x) : INTEGER(x);
}
SEXP named_list(const std::vector<const char *> &names, const std::vector<SEXP> &vals) {
  SEXP out = PROTECT(Rf_allocVector(VECSXP, (R_xlen_t)vals.size()));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, (R_xlen_t)vals.size()));
  for (size_t i = 0; i < vals.size(); i++) {
    SET_VECTOR_ELT(out, (R_xlen_t)i, vals[i]);
    SET_STRING_ELT(nm, (R_xlen_t)i, Rf_mkChar(names[i]));
  }
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(2);
  return out;
}

// ---- the session: a group of one context per GPU (flexgsea(parallel = N): N GPUs) ------------------
void session_finalizer(SEXP p) {
  fgs_group *g = static_cast<fgs_group *>(R_ExternalPtrAddr(p));
  if (g) { fgs_group_destroy(g); R_ClearExternalPtr(p); }
}
fgs_group *session_of(SEXP p) {
  if (TYPEOF(p) != EXTPTRSXP) Rcpp::stop("flexgsea: not a GPU session");
  fgs_group *g = static_cast<fgs_group *>(R_ExternalPtrAddr(p));
  if (!g) Rcpp::stop("flexgsea: the GPU session was released");
  return g;
}

// the tile a GPU helper thread is still reading (fgs_es_from_scores_tile): kept from R's collector
// until the next tile or the finish has returned
SEXP g_tile_in_flight = nullptr;
void hold_tile(SEXP s) {
  if (g_tile_in_flight) R_ReleaseObject(g_tile_in_flight);
  g_tile_in_flight = s;
  if (s) R_PreserveObject(s);
}

// the package-level context s2n_C() runs on (created on first use, lives as long as the DLL)
fgs_ctx *g_s2n_ctx = nullptr;
fgs_ctx *s2n_ctx() {
  if (!g_s2n_ctx) check(fgs_create(0, &g_s2n_ctx));
  return g_s2n_ctx;
}

// Interrupt polling between blocks of the null loop: R_CheckUserInterrupt longjmps, so it runs
// inside R_ToplevelExec and the answer goes back to the library as "stop" (FGS_ERR_CANCELLED),
// which unwinds normally; the wrapper then raises the interrupt from R's side.
void poll_interrupt(void *) { R_CheckUserInterrupt(); }
int progress_hook(void *, long long, long long) { return R_ToplevelExec(poll_interrupt, nullptr) == FALSE; }

}  // namespace

extern "C" {

// s2n_C(x, y): src/ggsea.cpp:6-76 bit for bit; the .Call entry of src/RcppExports.cpp:10-19
SEXP _flexgsea_s2n_C(SEXP xSEXP, SEXP ySEXP) {
  BEGIN_RCPP
  SEXP x = xSEXP;
  int nprot = 0;
  if (TYPEOF(x) == INTSXP) { x = PROTECT(Rf_coerceVector(x, REALSXP)); nprot++; }  // Rcpp's input_parameter does the same (:14)
  const int n = dim_of(x, 0), G = dim_of(x, 1), ny = dim_of(ySEXP, 0), R = dim_of(ySEXP, 1);
  SEXP coef = PROTECT(Rf_allocMatrix(REALSXP, G, R)); nprot++;
  check(fgs_s2n(s2n_ctx(), real_arg(x, "x"), n, G, int_arg(ySEXP, "y"), ny, R, REAL(coef)));
  UNPROTECT(nprot);
  return coef;
  END_RCPP
}

SEXP _flexgsea_n_gpu() {
  BEGIN_RCPP
  int n = 0;
  if (fgs_device_count(&n) != FGS_OK) n = 0;  // no driver / no device: the R closures stay in charge
  return Rf_ScalarInteger(n);
  END_RCPP
}

SEXP _flexgsea_session_create(SEXP n_gpuSEXP) {
  BEGIN_RCPP
  fgs_group *g = nullptr;
  check(fgs_group_create(Rf_asInteger(n_gpuSEXP), nullptr, &g));
  SEXP p = PROTECT(R_MakeExternalPtr(g, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(p, session_finalizer, TRUE);
  UNPROTECT(1);
  return p;
  END_RCPP
}

SEXP _flexgsea_session_close(SEXP sess) {
  BEGIN_RCPP
  if (TYPEOF(sess) == EXTPTRSXP) session_finalizer(sess);
  return R_NilValue;
  END_RCPP
}

// x: [samples x genes] (t.x = FALSE) or an EList's $E [genes x samples] (t.x = TRUE,
// R/flexgsea.R:159-173); the library wants a gene's samples contiguous, so E is transposed here
SEXP _flexgsea_session_upload(SEXP sess, SEXP xSEXP, SEXP t_xSEXP, SEXP offsets, SEXP index) {
  BEGIN_RCPP
  fgs_group *g = session_of(sess);
  SEXP x = xSEXP;
  int nprot = 0;
  if (TYPEOF(x) == INTSXP) { x = PROTECT(Rf_coerceVector(x, REALSXP)); nprot++; }
  const double *xv = real_arg(x, "x");
  const int d0 = dim_of(x, 0), d1 = dim_of(x, 1);
  if (Rf_asLogical(t_xSEXP)) {
    const int G = d0, n = d1;
    std::vector<double> t((size_t)G * n);
    for (int s = 0; s < n; s++)
      for (int gi = 0; gi < G; gi++) t[(size_t)gi * n + s] = xv[(size_t)s * G + gi];
    check(fgs_group_set_x(g, t.data(), n, G));
  } else {
    check(fgs_group_set_x(g, xv, d0, d1));
  }
  check(fgs_group_set_gene_sets(g, int_arg(offsets, "offsets"), int_arg(index, "index"), Rf_length(offsets) - 1));
  UNPROTECT(nprot);
  return R_NilValue;
  END_RCPP
}

SEXP _flexgsea_session_response_s2n(SEXP sess, SEXP y_bin) {
  BEGIN_RCPP
  check(fgs_group_set_response_s2n(session_of(sess), int_arg(y_bin, "y_bin"), dim_of(y_bin, 0), dim_of(y_bin, 1)));
  return R_NilValue;
  END_RCPP
}

SEXP _flexgsea_session_response_lm(SEXP sess, SEXP design, SEXP has_intercept) {
  BEGIN_RCPP
  check(fgs_group_set_response_lm(session_of(sess), real_arg(design, "design"), dim_of(design, 0),
                                  dim_of(design, 1), Rf_asLogical(has_intercept)));
  return R_NilValue;
  END_RCPP
}

SEXP _flexgsea_session_option(SEXP sess, SEXP name, SEXP value) {
  BEGIN_RCPP
  check(fgs_group_set_option(session_of(sess), CHAR(STRING_ELT(name, 0)), (long long)Rf_asReal(value)));
  return R_NilValue;
  END_RCPP
}

SEXP _flexgsea_session_interruptible(SEXP sess, SEXP on) {
  BEGIN_RCPP
  fgs_group *g = session_of(sess);
  for (int r = 0; r < fgs_group_size(g); r++)  // rank 0 polls; the others stop with it at the exchange
    check(fgs_set_progress(fgs_group_ctx(g, r), (r == 0 && Rf_asLogical(on)) ? progress_hook : nullptr, nullptr));
  return R_NilValue;
  END_RCPP
}

// gene.score.fn(x, y, abs) for the built-ins, on the device: [n_gene x responses]
SEXP _flexgsea_session_score_observed_n(SEXP sess, SEXP score_kind, SEXP abs, SEXP n_gene) {
  BEGIN_RCPP
  fgs_group *g = session_of(sess);
  int R = 0;
  check(fgs_n_response(fgs_group_ctx(g, 0), Rf_asInteger(score_kind), &R));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, Rf_asInteger(n_gene), R));
  check(fgs_group_score_observed(g, Rf_asInteger(score_kind), Rf_asLogical(abs), REAL(out)));
  UNPROTECT(1);
  return out;
  END_RCPP
}

// es.fn$run on the observed scores with every extra (R/flexgsea.R:266-292, :841-908).
// scores [genes x responses]; nnz = length of the CSR index the session holds.  Returns
// list(es, max_es_at, le_prop [S x R]; running_es_at (int), running_es_pos, running_es_neg
// [nnz x R]; le_first, le_count [S x R]) -- the ragged ones share the CSR offsets and are NULL
// unless want_ragged.  The observed ES stays in the session for the null loop (near-tie hand-over).
SEXP _flexgsea_session_observed_es(SEXP sess, SEXP es_kind, SEXP p_exp, SEXP scores, SEXP n_sets, SEXP nnz,
                                   SEXP want_ragged) {
  BEGIN_RCPP
  fgs_group *g = session_of(sess);
  const int G = dim_of(scores, 0), R = dim_of(scores, 1), S = Rf_asInteger(n_sets), kind = Rf_asInteger(es_kind);
  const int NZ = Rf_asInteger(nnz);
  const bool wks = kind == FGS_ES_WEIGHTED_KS, ragged = wks && Rf_asLogical(want_ragged);
  SEXP es = PROTECT(Rf_allocMatrix(REALSXP, S, R));
  SEXP mea = PROTECT(wks ? Rf_allocMatrix(REALSXP, S, R) : R_NilValue);
  SEXP lep = PROTECT(wks ? Rf_allocMatrix(REALSXP, S, R) : R_NilValue);
  SEXP lf = PROTECT(wks ? Rf_allocMatrix(INTSXP, S, R) : R_NilValue);
  SEXP lc = PROTECT(wks ? Rf_allocMatrix(INTSXP, S, R) : R_NilValue);
  SEXP at = PROTECT(ragged ? Rf_allocMatrix(INTSXP, NZ, R) : R_NilValue);
  SEXP rp = PROTECT(ragged ? Rf_allocMatrix(REALSXP, NZ, R) : R_NilValue);
  SEXP rn = PROTECT(ragged ? Rf_allocMatrix(REALSXP, NZ, R) : R_NilValue);
  check(fgs_group_observed_es(g, kind, Rf_asReal(p_exp), real_arg(scores, "scores"), G, R, REAL(es),
                              wks ? REAL(mea) : nullptr, wks ? REAL(lep) : nullptr, ragged ? INTEGER(at) : nullptr,
                              ragged ? REAL(rp) : nullptr, ragged ? REAL(rn) : nullptr, wks ? INTEGER(lf) : nullptr,
                              wks ? INTEGER(lc) : nullptr));
  SEXP out = named_list({"es", "max_es_at", "le_prop", "running_es_at", "running_es_pos", "running_es_neg",
                         "le_first", "le_count"},
                        {es, mea, lep, at, rp, rn, lf, lc});
  UNPROTECT(8);
  return out;
  END_RCPP
}

// es.null <- perm.fun(...) (R/flexgsea.R:305).  perm: integer matrix [samples x nperm] whose
// columns are sample.int(n) draws made in R, in order (:384) -- the null is then the one the
// reference's sequential loop computes under the same set.seed() -- or NULL for permutations
// drawn on the device (Philox, `seed`).  Returns es.null [S, R, nperm] (:363-368) or, when
// want_null is FALSE (built-in sig.fun, es_null not in return_values), a zero-length vector:
// the null stays on the GPUs for _flexgsea_session_calc_sig.
SEXP _flexgsea_session_perm_null(SEXP sess, SEXP score_kind, SEXP es_kind, SEXP p_exp, SEXP abs, SEXP perm,
                                 SEXP seed, SEXP nperm, SEXP n_sets, SEXP want_null) {
  BEGIN_RCPP
  fgs_group *g = session_of(sess);
  const int P = Rf_asInteger(nperm), S = Rf_asInteger(n_sets), sk = Rf_asInteger(score_kind);
  int R = 0;
  check(fgs_n_response(fgs_group_ctx(g, 0), sk, &R));
  const int *pm = nullptr;
  if (!Rf_isNull(perm)) {
    if (dim_of(perm, 1) != P) Rcpp::stop("flexgsea: perm must have nperm columns");
    pm = int_arg(perm, "perm");
  }
  const bool want = Rf_asLogical(want_null);
  SEXP out;
  if (want) {
    out = PROTECT(Rf_alloc3DArray(REALSXP, S, R, P));
  } else {
    out = PROTECT(Rf_allocVector(REALSXP, 0));
  }
  const int rc = fgs_group_perm_null(g, sk, Rf_asInteger(es_kind), Rf_asReal(p_exp), Rf_asLogical(abs), pm,
                                     (unsigned long long)Rf_asReal(seed), P, want ? REAL(out) : nullptr);
  if (rc == FGS_ERR_CANCELLED) { UNPROTECT(1); Rf_onintr(); return R_NilValue; }
  check(rc);
  UNPROTECT(1);
  return out;
  END_RCPP
}

// the same loop with a user-defined gene.score.fn: the R closure made scores [G, R, P]
// (R/flexgsea.R:393-394); ranking onward on the GPUs, blocks streamed up while the previous is ranked
SEXP _flexgsea_session_es_from_scores(SEXP sess, SEXP es_kind, SEXP p_exp, SEXP scores, SEXP n_sets,
                                      SEXP want_null) {
  BEGIN_RCPP
  fgs_group *g = session_of(sess);
  const int G = dim_of(scores, 0), R = dim_of(scores, 1), P = dim_of(scores, 2), S = Rf_asInteger(n_sets);
  const bool want = Rf_asLogical(want_null);
  SEXP out = PROTECT(want ? Rf_alloc3DArray(REALSXP, S, R, P) : Rf_allocVector(REALSXP, 0));
  const int rc = fgs_group_es_from_scores(g, Rf_asInteger(es_kind), Rf_asReal(p_exp), real_arg(scores, "scores"), G,
                                          R, P, want ? REAL(out) : nullptr);
  if (rc == FGS_ERR_CANCELLED) { UNPROTECT(1); Rf_onintr(); return R_NilValue; }
  check(rc);
  UNPROTECT(1);
  return out;
  END_RCPP
}

// One tile of that loop: scores [G, R, B] of permutations perm_offset .. (0-based) out of nperm.
// Only queues the work, so R goes on scoring the next tile while this one is uploaded, ranked
// and scored; _flexgsea_session_null_finish waits and returns es.null (or a zero-length vector).
SEXP _flexgsea_session_es_from_scores_tile(SEXP sess, SEXP es_kind, SEXP p_exp, SEXP scores, SEXP perm_offset,
                                           SEXP nperm) {
  BEGIN_RCPP
  fgs_group *g = session_of(sess);
  const int G = dim_of(scores, 0), R = dim_of(scores, 1), B = dim_of(scores, 2);
  const int rc = fgs_group_es_from_scores_tile(g, Rf_asInteger(es_kind), Rf_asReal(p_exp), real_arg(scores, "scores"),
                                               G, R, B, (long long)Rf_asReal(perm_offset), (long long)Rf_asReal(nperm));
  hold_tile(rc == FGS_OK ? scores : nullptr);   // (the previous tile is through once the call has returned)
  if (rc == FGS_ERR_CANCELLED) { Rf_onintr(); return R_NilValue; }
  check(rc);
  return R_NilValue;
  END_RCPP
}

SEXP _flexgsea_session_null_finish(SEXP sess, SEXP n_sets, SEXP n_response, SEXP nperm, SEXP want_null) {
  BEGIN_RCPP
  fgs_group *g = session_of(sess);
  const bool want = Rf_asLogical(want_null);
  SEXP out = PROTECT(want ? Rf_alloc3DArray(REALSXP, Rf_asInteger(n_sets), Rf_asInteger(n_response), Rf_asInteger(nperm))
                          : Rf_allocVector(REALSXP, 0));
  const int rc = fgs_group_null_finish(g, want ? REAL(out) : nullptr);
  hold_tile(nullptr);
  if (rc == FGS_ERR_CANCELLED) { UNPROTECT(1); Rf_onintr(); return R_NilValue; }
  check(rc);
  UNPROTECT(1);
  return out;
  END_RCPP
}

// sig.fun(es[, r], es.null[, r, ]) on the null the GPUs hold (R/flexgsea.R:314-330): the columns
// of flexgsea_calc_sig's tibble (:564-629) or of flexgsea_calc_sig_simple's (:544-547), in order
SEXP _flexgsea_session_calc_sig(SEXP sess, SEXP esSEXP, SEXP response, SEXP split_p, SEXP calc_nes, SEXP abs) {
  BEGIN_RCPP
  fgs_group *g = session_of(sess);
  const int S = Rf_length(esSEXP);
  const bool sp = Rf_asLogical(split_p), cn = Rf_asLogical(calc_nes), ab = Rf_asLogical(abs);
  SEXP p = PROTECT(Rf_allocVector(REALSXP, S)), ph = PROTECT(Rf_allocVector(REALSXP, S));
  SEXP pl = PROTECT(Rf_allocVector(REALSXP, S)), nes = PROTECT(Rf_allocVector(REALSXP, S));
  SEXP fdr = PROTECT(Rf_allocVector(REALSXP, S)), fwer = PROTECT(Rf_allocVector(REALSXP, S));
  check(fgs_group_calc_sig(g, real_arg(esSEXP, "es"), Rf_asInteger(response), sp, cn, ab, 0, REAL(p), REAL(ph),
                           REAL(pl), REAL(nes), REAL(fdr), REAL(fwer), nullptr, nullptr, nullptr));
  SEXP out;
  if (sp && cn) out = named_list({"es", "p", "nes", "fdr", "fwer"}, {esSEXP, p, nes, fdr, fwer});
  else if (ab) out = named_list({"es", "p.high", "p", "fdr", "fwer"}, {esSEXP, ph, p, fdr, fwer});
  else out = named_list({"es", "p.high", "p.low", "p", "fdr", "fwer"}, {esSEXP, ph, pl, p, fdr, fwer});
  UNPROTECT(6);
  return out;
  END_RCPP
}

// flexgsea_calc_sig(es, es_null) called on its own (exported, R/flexgsea.R:564): the null comes from R
SEXP _flexgsea_session_set_null(SEXP sess, SEXP es_null) {
  BEGIN_RCPP
  fgs_group *g = session_of(sess);
  if (fgs_group_size(g) != 1) Rcpp::stop("flexgsea: an R-made null is loaded on a single-GPU session");
  check(fgs_set_null(fgs_group_ctx(g, 0), real_arg(es_null, "es_null"), dim_of(es_null, 0), 1, dim_of(es_null, 1)));
  return R_NilValue;
  END_RCPP
}

// read_gmt + filter_gene_sets + match(set, gene.names) (R/flexgsea.R:966-973, :939-958, :279 / :412)
// natively and once: list(names, offsets, index, genes_before, genes_after, sets_before)
SEXP _flexgsea_gmt_read(SEXP path, SEXP gene_names, SEXP gs_size_min, SEXP gs_size_max) {
  BEGIN_RCPP
  const R_xlen_t ng = Rf_length(gene_names);
  std::vector<const char *> names((size_t)ng);
  for (R_xlen_t i = 0; i < ng; i++) names[(size_t)i] = CHAR(STRING_ELT(gene_names, i));
  fgs_gmt *h = nullptr;
  check(fgs_gmt_read(CHAR(STRING_ELT(path, 0)), names.data(), (int)ng, Rf_asInteger(gs_size_min),
                     Rf_asInteger(gs_size_max), &h));
  const int S = fgs_gmt_n_sets(h);
  const long long nnz = fgs_gmt_nnz(h);
  SEXP off = PROTECT(Rf_allocVector(INTSXP, S + 1)), idx = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)nnz));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, S));
  std::memcpy(INTEGER(off), fgs_gmt_offsets(h), sizeof(int) * (size_t)(S + 1));
  if (nnz > 0) std::memcpy(INTEGER(idx), fgs_gmt_index(h), sizeof(int) * (size_t)nnz);
  for (int k = 0; k < S; k++) SET_STRING_ELT(nm, k, Rf_mkChar(fgs_gmt_set_name(h, k)));
  long long gb = 0, ga = 0;
  int sb = 0;
  fgs_gmt_stats(h, &gb, &ga, &sb);
  fgs_gmt_free(h);
  SEXP out = named_list({"names", "offsets", "index", "genes_before", "genes_after", "sets_before"},
                        {nm, off, idx, Rf_ScalarReal((double)gb), Rf_ScalarReal((double)ga), Rf_ScalarInteger(sb)});
  UNPROTECT(3);
  return out;
  END_RCPP
}

// ---- registration (replaces src/RcppExports.cpp:21-29) ------------------------------------------
static const R_CallMethodDef CallEntries[] = {
    {"_flexgsea_s2n_C", (DL_FUNC)&_flexgsea_s2n_C, 2},
    {"_flexgsea_n_gpu", (DL_FUNC)&_flexgsea_n_gpu, 0},
    {"_flexgsea_session_create", (DL_FUNC)&_flexgsea_session_create, 1},
    {"_flexgsea_session_close", (DL_FUNC)&_flexgsea_session_close, 1},
    {"_flexgsea_session_upload", (DL_FUNC)&_flexgsea_session_upload, 5},
    {"_flexgsea_session_response_s2n", (DL_FUNC)&_flexgsea_session_response_s2n, 2},
    {"_flexgsea_session_response_lm", (DL_FUNC)&_flexgsea_session_response_lm, 3},
    {"_flexgsea_session_option", (DL_FUNC)&_flexgsea_session_option, 3},
    {"_flexgsea_session_interruptible", (DL_FUNC)&_flexgsea_session_interruptible, 2},
    {"_flexgsea_session_score_observed_n", (DL_FUNC)&_flexgsea_session_score_observed_n, 4},
    {"_flexgsea_session_observed_es", (DL_FUNC)&_flexgsea_session_observed_es, 7},
    {"_flexgsea_session_perm_null", (DL_FUNC)&_flexgsea_session_perm_null, 10},
    {"_flexgsea_session_es_from_scores", (DL_FUNC)&_flexgsea_session_es_from_scores, 6},
    {"_flexgsea_session_es_from_scores_tile", (DL_FUNC)&_flexgsea_session_es_from_scores_tile, 6},
    {"_flexgsea_session_null_finish", (DL_FUNC)&_flexgsea_session_null_finish, 5},
    {"_flexgsea_session_calc_sig", (DL_FUNC)&_flexgsea_session_calc_sig, 6},
    {"_flexgsea_session_set_null", (DL_FUNC)&_flexgsea_session_set_null, 2},
    {"_flexgsea_gmt_read", (DL_FUNC)&_flexgsea_gmt_read, 4},
    {NULL, NULL, 0}};

void R_init_flexgsea(DllInfo *dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);  // only registered symbols resolve, as in the reference
}

}  // extern "C"
