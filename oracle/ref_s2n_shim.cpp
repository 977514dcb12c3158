// Compiles the reference's src/ggsea.cpp VERBATIM (included from
// /root/reference at build time, never copied into this repo) against the
// stand-in Rcpp.h and exposes it with a C ABI for the tests.
// Output goes to oracle/_ref/ only.  Test infrastructure.
#include <Rcpp.h>
#include <cstring>
#include REF_GGSEA_CPP
extern "C" void ref_s2n_C(const double *x, int n_x, int G, const int *y, int n,
                          int R, double *coef) {
  Rcpp::NumericMatrix xm(const_cast<double *>(x), n_x, G);
  Rcpp::LogicalMatrix ym(const_cast<int *>(y), n, R);
  Rcpp::NumericMatrix out = s2n_C(xm, ym);
  std::memcpy(coef, out.data(), sizeof(double) * (size_t)G * R);
}
