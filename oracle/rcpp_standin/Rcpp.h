// Stand-in for <Rcpp.h>: just enough surface for the reference's
// src/ggsea.cpp to compile verbatim outside R (test infrastructure only).
// Column-major matrices with nrow()/ncol()/operator(), NA_LOGICAL = INT_MIN.
#pragma once
#include <climits>
#include <vector>
#ifndef NA_LOGICAL
#define NA_LOGICAL INT_MIN
#endif
namespace Rcpp {
template <typename T> class Mat {
  int nr_, nc_; std::vector<T> own_; T *p_;
 public:
  Mat(int nr, int nc) : nr_(nr), nc_(nc), own_((size_t)nr * nc, T(0)), p_(own_.data()) {}
  Mat(T *p, int nr, int nc) : nr_(nr), nc_(nc), p_(p) {}
  Mat(const Mat &o) : nr_(o.nr_), nc_(o.nc_), own_(o.own_), p_(o.own_.empty() ? o.p_ : own_.data()) {}
  int nrow() const { return nr_; }
  int ncol() const { return nc_; }
  T &operator()(int i, int j) { return p_[(size_t)j * nr_ + i]; }
  const T &operator()(int i, int j) const { return p_[(size_t)j * nr_ + i]; }
  const T *data() const { return p_; }
};
typedef Mat<double> NumericMatrix;
typedef Mat<int> LogicalMatrix;
}  // namespace Rcpp
