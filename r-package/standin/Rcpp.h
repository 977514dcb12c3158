// Functional stand-in for the slice of <Rcpp.h> / R's C API that r-package/src/flexgsea_native.cpp
// uses.  TEST INFRASTRUCTURE: R is not installed in this repository's image, so the .Call entry
// points are compiled against this header and driven by tests/test_r_shim.py the way R would
// drive them (lookup in the registered table, arity check, SEXP arguments, errors as conditions).
// Values are heap objects that are never collected (PROTECT is a no-op); external pointers run
// their finalizer when the harness asks.  Not R, not Rcpp: just enough of their surface.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

typedef long R_xlen_t;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

enum { NILSXP = 0, CHARSXP = 9, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22 };

struct SEXPREC;
typedef SEXPREC *SEXP;
struct SEXPREC {
  int type = NILSXP;
  std::vector<double> real;
  std::vector<int> ints;        // INTSXP / LGLSXP
  std::vector<SEXP> elts;       // VECSXP / STRSXP (CHARSXP elements)
  std::string chars;            // CHARSXP
  SEXP dim = nullptr, names = nullptr;
  void *ptr = nullptr;          // EXTPTRSXP
  void (*fin)(SEXP) = nullptr;
};

extern "C" {
extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;
SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nr, int nc);
SEXP Rf_alloc3DArray(int type, int a, int b, int c);
SEXP Rf_coerceVector(SEXP x, int type);
double *REAL(SEXP x);
int *INTEGER(SEXP x);
int *LOGICAL(SEXP x);
int TYPEOF(SEXP x);
R_xlen_t Rf_length(SEXP x);
int Rf_isNull(SEXP x);
SEXP Rf_getAttrib(SEXP x, SEXP what);
SEXP Rf_setAttrib(SEXP x, SEXP what, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP Rf_mkChar(const char *s);
const char *CHAR(SEXP x);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
int Rf_asLogical(SEXP x);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_ScalarReal(double v);
SEXP Rf_ScalarLogical(int v);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, void (*fin)(SEXP), Rboolean onexit);
void R_PreserveObject(SEXP s);
void R_ReleaseObject(SEXP s);
Rboolean R_ToplevelExec(void (*fn)(void *), void *data);
void R_CheckUserInterrupt(void);
void Rf_onintr(void);

typedef void *(*DL_FUNC)();
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call, const void *f, const void *e);
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value);
// the harness's side of an R error: the message of the condition the entry point raised
void standin_raise(const char *msg);
}

#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)(n))

namespace Rcpp {
struct exception : std::runtime_error { using std::runtime_error::runtime_error; };
inline void stop(const std::string &msg) { throw exception(msg); }
inline void stop(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw exception(buf);
}
}  // namespace Rcpp

// Rcpp turns C++ exceptions into R conditions after the C++ frames have unwound
#define BEGIN_RCPP try {
#define END_RCPP                                                   \
  }                                                                \
  catch (std::exception & e__) { standin_raise(e__.what()); }      \
  catch (...) { standin_raise("c++ exception (unknown reason)"); } \
  return R_NilValue;
