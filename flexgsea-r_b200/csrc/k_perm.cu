// K1: permutation draw and permuted class labels.
//
// Replaces `y.perm <- y[sample.int(nrow(y)), , drop=F]` of the reference
// (R/flexgsea.R:384, :443).  Either an R-made permutation matrix is used as is
// (same RNG stream as the reference), or each permutation is drawn on the
// device: Fisher-Yates driven by Philox4x32-10 keyed by (seed, global
// permutation id), so results do not depend on how permutations are sharded.
#include <algorithm>

#include "fgs_common.cuh"

namespace fgs {

__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// one thread per permutation; the permutation lives in its column of d_perm
__global__ void philox_perm_kernel(unsigned long long seed, long long perm_id0, int n, int P,
                                   int *__restrict__ perm) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int *a = perm + (size_t)p * n;
  for (int i = 0; i < n; i++) a[i] = i + 1;
  unsigned long long id = (unsigned long long)(perm_id0 + p);
  uint32_t buf[4];
  int have = 0;
  uint32_t chunk = 0;
  for (int i = n - 1; i >= 1; i--) {
    if (!have) {
      buf[0] = chunk++; buf[1] = 0; buf[2] = (uint32_t)id; buf[3] = (uint32_t)(id >> 32);
      philox4x32_10(buf, (uint32_t)seed, (uint32_t)(seed >> 32));
      have = 4;
    }
    uint32_t u = buf[4 - have];
    have--;
    uint32_t j = __umulhi(u, (uint32_t)(i + 1));
    int t = a[i]; a[i] = a[j]; a[j] = t;
  }
}

int launch_philox_perm(fgs_ctx *c, unsigned long long seed, long long perm_id0, int n, int P,
                       int *d_perm) {
  if (P <= 0) return FGS_OK;
  philox_perm_kernel<<<ceil_div(P, 128), 128, 0, c->stream>>>(seed, perm_id0, n, P, d_perm);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// labels[col][0][w] = class-1 mask, labels[col][1][w] = class-2 mask of the
// permuted y_bin column; col = p * R + r.  NA_LOGICAL samples are in neither
// mask (src/ggsea.cpp:29).  One warp per column.  A permutation entry outside
// [1, n] raises flag bit 8 in *bad.
__global__ void labels_kernel(const int *__restrict__ perm, const int *__restrict__ ybin, int n,
                              int R, long long C, int W, uint32_t *__restrict__ labels,
                              int *__restrict__ bad) {
  long long col = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (col >= C) return;
  long long p = col / R;
  int r = (int)(col % R);
  const int *pi = perm ? perm + (size_t)p * n : nullptr;
  const int *yc = ybin + (size_t)r * n;
  for (int wb = 0; wb < W; wb++) {
    int i = wb * 32 + lane;
    int v = INT_MIN;
    if (i < n) {
      int src = pi ? pi[i] - 1 : i;
      if (src < 0 || src >= n) { atomicOr(bad, 8); src = 0; }
      v = yc[src];
    }
    uint32_t m1 = __ballot_sync(0xffffffffu, v != INT_MIN && v != 0);
    uint32_t m2 = __ballot_sync(0xffffffffu, v != INT_MIN && v == 0);
    if (lane == 0) {
      labels[(col * 2 + 0) * W + wb] = m1;
      labels[(col * 2 + 1) * W + wb] = m2;
    }
  }
}

// ident[p] = 1 when permutation p leaves the class labels of EVERY response as they are
// (label-preserving permutation: one in twenty at 3 + 3 samples).  Its scores, ranks and
// ES are the observed ones, which fgs_perm_null then copies instead of recomputing them
// in the reference's arithmetic unit by unit.  obs = labels of the identity, [R][2][W].
__global__ void label_identity_kernel(const uint32_t *__restrict__ labels, const uint32_t *__restrict__ obs,
                                      int R, int W, int P, int *__restrict__ ident) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const uint32_t *lp = labels + (size_t)p * R * 2 * W;
  int same = 1;
  for (int i = 0; i < R * 2 * W; i++) same &= lp[i] == obs[i];
  ident[p] = same;
}
int launch_label_identity(fgs_ctx *c, const uint32_t *d_labels, const uint32_t *d_obs, int R, int W, int P,
                          int *d_ident) {
  if (P <= 0) return FGS_OK;
  label_identity_kernel<<<ceil_div(P, 128), 128, 0, c->stream>>>(d_labels, d_obs, R, W, P, d_ident);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// es_null[, , p] = es_uniq[, , map[p]]: every permutation takes the column of its labelling
__global__ void scatter_null_kernel(const double *__restrict__ uniq, const int *__restrict__ map, long long unit,
                                    double *__restrict__ null) {
  const int p = blockIdx.y;
  const double *src = uniq + (size_t)map[p] * unit;
  double *dst = null + (size_t)p * unit;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < unit; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
int launch_scatter_null(fgs_ctx *c, const double *d_uniq, const int *d_map, long long unit, int P, double *d_null) {
  if (P <= 0 || unit <= 0) return FGS_OK;
  const int gx = (int)std::min<long long>(ceil_div(unit, 256), 64);
  for (int p0 = 0; p0 < P; p0 += 65535) {  // grid.y limit
    const int np = std::min(65535, P - p0);
    scatter_null_kernel<<<dim3(gx, np), 256, 0, c->stream>>>(d_uniq, d_map + p0, unit, d_null + (size_t)p0 * unit);
  }
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

int launch_labels(fgs_ctx *c, const int *d_perm, const int *d_ybin, int n, int R, int P,
                  uint32_t *d_labels) {
  long long C = (long long)P * R;
  if (C <= 0) return FGS_OK;
  int W = (n + 31) / 32;
  labels_kernel<<<ceil_div(C, 8), 256, 0, c->stream>>>(d_perm, d_ybin, n, R, C, W, d_labels,
                                                       c->flags.p + (c->flags.n - 1));
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

}  // namespace fgs
