// The stand-in's runtime (see Rcpp.h here) and the C entry points tests/test_r_shim.py uses to
// play R: build SEXPs, .Call a registered routine by name (arity checked like R does), read
// results, raise a pending interrupt, finalize external pointers.  Test infrastructure only.
#include <cstring>
#include <map>

#include "Rcpp.h"

static SEXPREC nil_obj, dim_sym, names_sym;
extern "C" {
SEXP R_NilValue = &nil_obj, R_DimSymbol = &dim_sym, R_NamesSymbol = &names_sym;
}
static std::string g_condition;
static bool g_have_condition = false, g_interrupt_pending = false, g_interrupt_raised = false;
static std::map<std::string, std::pair<DL_FUNC, int>> g_routines;
static bool g_dynamic_symbols = true;
struct Unwind {};  // what a longjmp out of R_CheckUserInterrupt becomes here

extern "C" {

SEXP Rf_allocVector(int type, R_xlen_t n) {
  SEXP s = new SEXPREC();
  s->type = type;
  if (type == REALSXP) s->real.assign((size_t)n, 0.0);
  else if (type == INTSXP || type == LGLSXP) s->ints.assign((size_t)n, 0);
  else if (type == VECSXP) s->elts.assign((size_t)n, R_NilValue);
  else if (type == STRSXP) s->elts.assign((size_t)n, Rf_mkChar(""));
  return s;
}
static SEXP dims(std::initializer_list<int> d) {
  SEXP s = Rf_allocVector(INTSXP, (R_xlen_t)d.size());
  int i = 0;
  for (int v : d) s->ints[i++] = v;
  return s;
}
SEXP Rf_allocMatrix(int type, int nr, int nc) {
  SEXP s = Rf_allocVector(type, (R_xlen_t)nr * nc);
  s->dim = dims({nr, nc});
  return s;
}
SEXP Rf_alloc3DArray(int type, int a, int b, int c) {
  SEXP s = Rf_allocVector(type, (R_xlen_t)a * b * c);
  s->dim = dims({a, b, c});
  return s;
}
SEXP Rf_coerceVector(SEXP x, int type) {
  if (x->type == type) return x;
  SEXP s = Rf_allocVector(type, Rf_length(x));
  s->dim = x->dim; s->names = x->names;
  if (type == REALSXP && (x->type == INTSXP || x->type == LGLSXP))
    for (size_t i = 0; i < x->ints.size(); i++) s->real[i] = x->ints[i];
  else if ((type == INTSXP || type == LGLSXP) && x->type == REALSXP)
    for (size_t i = 0; i < x->real.size(); i++) s->ints[i] = (int)x->real[i];
  else if ((type == INTSXP || type == LGLSXP) && (x->type == INTSXP || x->type == LGLSXP)) s->ints = x->ints;
  else throw Rcpp::exception("standin: unsupported coercion");
  return s;
}
double *REAL(SEXP x) { if (x->type != REALSXP) throw Rcpp::exception("REAL() can only be applied to a 'numeric'"); return x->real.data(); }
int *INTEGER(SEXP x) { if (x->type != INTSXP && x->type != LGLSXP) throw Rcpp::exception("INTEGER() can only be applied to a 'integer'"); return x->ints.data(); }
int *LOGICAL(SEXP x) { if (x->type != LGLSXP) throw Rcpp::exception("LOGICAL() can only be applied to a 'logical'"); return x->ints.data(); }
int TYPEOF(SEXP x) { return x->type; }
R_xlen_t Rf_length(SEXP x) {
  switch (x->type) {
    case REALSXP: return (R_xlen_t)x->real.size();
    case INTSXP: case LGLSXP: return (R_xlen_t)x->ints.size();
    case VECSXP: case STRSXP: return (R_xlen_t)x->elts.size();
    case NILSXP: return 0;
    default: return 1;
  }
}
int Rf_isNull(SEXP x) { return x == nullptr || x->type == NILSXP; }
SEXP Rf_getAttrib(SEXP x, SEXP what) {
  SEXP v = what == R_DimSymbol ? x->dim : (what == R_NamesSymbol ? x->names : nullptr);
  return v ? v : R_NilValue;
}
SEXP Rf_setAttrib(SEXP x, SEXP what, SEXP v) {
  if (what == R_DimSymbol) x->dim = v; else if (what == R_NamesSymbol) x->names = v;
  return v;
}
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return x->elts.at((size_t)i); }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { x->elts.at((size_t)i) = v; return v; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { if (x->type != STRSXP) throw Rcpp::exception("STRING_ELT() can only be applied to a 'character vector'"); return x->elts.at((size_t)i); }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { x->elts.at((size_t)i) = v; }
SEXP Rf_mkChar(const char *s) { SEXP c = new SEXPREC(); c->type = CHARSXP; c->chars = s; return c; }
const char *CHAR(SEXP x) { return x->chars.c_str(); }
int Rf_asInteger(SEXP x) {
  if (x->type == REALSXP && !x->real.empty()) return (int)x->real[0];
  if ((x->type == INTSXP || x->type == LGLSXP) && !x->ints.empty()) return x->ints[0];
  throw Rcpp::exception("standin: asInteger of a non-number");
}
double Rf_asReal(SEXP x) {
  if (x->type == REALSXP && !x->real.empty()) return x->real[0];
  if ((x->type == INTSXP || x->type == LGLSXP) && !x->ints.empty()) return x->ints[0];
  throw Rcpp::exception("standin: asReal of a non-number");
}
int Rf_asLogical(SEXP x) { return Rf_asInteger(x) != 0; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); s->ints[0] = v; return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); s->real[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); s->ints[0] = v; return s; }
SEXP R_MakeExternalPtr(void *p, SEXP, SEXP) { SEXP s = new SEXPREC(); s->type = EXTPTRSXP; s->ptr = p; return s; }
void *R_ExternalPtrAddr(SEXP s) { return s->ptr; }
void R_ClearExternalPtr(SEXP s) { s->ptr = nullptr; }
void R_RegisterCFinalizerEx(SEXP s, void (*fin)(SEXP), Rboolean) { s->fin = fin; }
static int g_preserved = 0;
void R_PreserveObject(SEXP) { g_preserved++; }
void R_ReleaseObject(SEXP) { g_preserved--; }
int standin_preserved(void) { return g_preserved; }
void R_CheckUserInterrupt(void) { if (g_interrupt_pending) throw Unwind(); }
Rboolean R_ToplevelExec(void (*fn)(void *), void *data) {
  try { fn(data); } catch (Unwind &) { return FALSE; }
  return TRUE;
}
void Rf_onintr(void) { g_interrupt_pending = false; g_interrupt_raised = true; }
int R_registerRoutines(DllInfo *, const void *, const R_CallMethodDef *call, const void *, const void *) {
  for (; call && call->name; call++) g_routines[call->name] = {call->fun, call->numArgs};
  return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *, Rboolean value) { g_dynamic_symbols = value; return TRUE; }
void standin_raise(const char *msg) { g_condition = msg; g_have_condition = true; }

// ---- what the Python harness calls ------------------------------------------------------------------
void R_init_flexgsea(DllInfo *);
void standin_load(void) { g_routines.clear(); R_init_flexgsea(nullptr); }
int standin_n_routines(void) { return (int)g_routines.size(); }
int standin_dynamic_symbols(void) { return g_dynamic_symbols; }
// arity of a registered routine, -1 when it is not registered
int standin_arity(const char *name) {
  auto it = g_routines.find(name);
  return it == g_routines.end() ? -1 : it->second.second;
}
const char *standin_routine_name(int k) {
  for (auto &kv : g_routines) if (k-- == 0) return kv.first.c_str();
  return nullptr;
}
// .Call(name, args...): NULL + a condition when the routine is unknown, the arity is wrong or it raised
SEXP standin_call(const char *name, int nargs, SEXP *a) {
  g_have_condition = false; g_interrupt_raised = false;
  auto it = g_routines.find(name);
  if (it == g_routines.end()) { standin_raise((std::string("\"") + name + "\" not available for .Call() for package \"flexgsea\"").c_str()); return nullptr; }
  if (it->second.second != nargs) {
    char b[256];
    snprintf(b, sizeof(b), "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, it->second.second, name);
    standin_raise(b);
    return nullptr;
  }
  typedef SEXP (*F0)(); typedef SEXP (*F1)(SEXP); typedef SEXP (*F2)(SEXP, SEXP); typedef SEXP (*F3)(SEXP, SEXP, SEXP);
  typedef SEXP (*F4)(SEXP, SEXP, SEXP, SEXP); typedef SEXP (*F5)(SEXP, SEXP, SEXP, SEXP, SEXP);
  typedef SEXP (*F6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP); typedef SEXP (*F7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
  typedef SEXP (*F10)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
  DL_FUNC f = it->second.first;
  SEXP r = nullptr;
  try {
    switch (nargs) {
      case 0: r = ((F0)f)(); break;
      case 1: r = ((F1)f)(a[0]); break;
      case 2: r = ((F2)f)(a[0], a[1]); break;
      case 3: r = ((F3)f)(a[0], a[1], a[2]); break;
      case 4: r = ((F4)f)(a[0], a[1], a[2], a[3]); break;
      case 5: r = ((F5)f)(a[0], a[1], a[2], a[3], a[4]); break;
      case 6: r = ((F6)f)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
      case 7: r = ((F7)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
      case 10: r = ((F10)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]); break;
      default: standin_raise("standin: arity not wired"); return nullptr;
    }
  } catch (std::exception &e) { standin_raise(e.what()); return nullptr; }
  return g_have_condition ? nullptr : r;
}
const char *standin_condition(void) { return g_have_condition ? g_condition.c_str() : nullptr; }
void standin_set_interrupt(int pending) { g_interrupt_pending = pending != 0; }
int standin_interrupt_raised(void) { return g_interrupt_raised; }
void standin_finalize(SEXP s) { if (s && s->type == EXTPTRSXP && s->fin) s->fin(s); }

SEXP standin_nil(void) { return R_NilValue; }
SEXP standin_real(const double *v, int nd, const int *d) {
  R_xlen_t n = 1;
  for (int i = 0; i < nd; i++) n *= d[i];
  SEXP s = Rf_allocVector(REALSXP, n);
  if (n) std::memcpy(s->real.data(), v, sizeof(double) * (size_t)n);
  if (nd > 1) { s->dim = Rf_allocVector(INTSXP, nd); for (int i = 0; i < nd; i++) s->dim->ints[i] = d[i]; }
  return s;
}
SEXP standin_int(const int *v, int nd, const int *d, int logical) {
  R_xlen_t n = 1;
  for (int i = 0; i < nd; i++) n *= d[i];
  SEXP s = Rf_allocVector(logical ? LGLSXP : INTSXP, n);
  if (n) std::memcpy(s->ints.data(), v, sizeof(int) * (size_t)n);
  if (nd > 1) { s->dim = Rf_allocVector(INTSXP, nd); for (int i = 0; i < nd; i++) s->dim->ints[i] = d[i]; }
  return s;
}
SEXP standin_strings(int n, const char *const *v) {
  SEXP s = Rf_allocVector(STRSXP, n);
  for (int i = 0; i < n; i++) s->elts[i] = Rf_mkChar(v[i]);
  return s;
}
int standin_type(SEXP s) { return s ? s->type : -1; }
long standin_length(SEXP s) { return Rf_length(s); }
int standin_ndim(SEXP s) { return s->dim ? (int)s->dim->ints.size() : 0; }
int standin_dim(SEXP s, int k) { return s->dim->ints.at(k); }
const double *standin_real_ptr(SEXP s) { return s->real.data(); }
const int *standin_int_ptr(SEXP s) { return s->ints.data(); }
const char *standin_string(SEXP s, int i) { return s->elts.at(i)->chars.c_str(); }
SEXP standin_elt(SEXP s, int i) { return s->elts.at(i); }
const char *standin_name(SEXP s, int i) { return s->names ? s->names->elts.at(i)->chars.c_str() : ""; }

}  // extern "C"
