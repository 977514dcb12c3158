// K2': bit-faithful signal-to-noise gene score for every (gene, response,
// permutation).
//
// Reproduces the reference's s2n_C (src/ggsea.cpp:6-76) operation by
// operation: per (gene, column) the samples are visited in index order, three
// passes (class sums -> corrected means -> squared deviations), the
// `mean2 = sum1 / n2` first-pass quirk (:40, corrected by pass 2), population
// variance (:69-70), IEEE division and square root, and explicitly
// non-contracted adds/multiplies (__dadd_rn/__dmul_rn; the reference is built
// for plain x86-64, no FMA).  Hence scores are bit-identical to the reference
// and so are the ranks derived from them.
//
// Mapping: lane <-> gene (32 genes per CTA tile, transposed into shared memory
// so a warp reads 32 consecutive doubles per sample), warp <-> PPT permutation
// columns.  The class label of a sample is then warp-uniform, so the
// class-conditional accumulation is a uniform predicate, and every x value
// loaded from shared memory is reused by PPT columns.  FP64-pipe bound:
// ~(n1 + 5n) double ops per (gene, column).
#include "fgs_common.cuh"

namespace fgs {

constexpr int S2N_THREADS = 256;
constexpr int S2N_PPT = 4;

template <int PPT, bool SMEM_TILE>
__global__ void __launch_bounds__(S2N_THREADS)
s2n_kernel(const double *__restrict__ x, int n, int G, const uint32_t *__restrict__ labels,
           const int *__restrict__ n1n2, int R, long long C, int W, long long cols_per_cta,
           double *__restrict__ scores, int *__restrict__ flags) {
  extern __shared__ double xs[];  // [n][32] when SMEM_TILE
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = blockIdx.x * 32 + lane;
  const bool gvalid = g < G;
  const double *xg = x + (size_t)(gvalid ? g : 0) * n;
  if (SMEM_TILE) {
    // lanes over genes: conflict-free shared stores; the strided global reads
    // hit L2 (x is resident there) and happen once per CTA
    for (int i = warp; i < n; i += nwarps) xs[i * 32 + lane] = gvalid ? xg[i] : 0.0;
    __syncthreads();
  }
  const long long cbeg = (long long)blockIdx.y * cols_per_cta;
  const long long cend = (cbeg + cols_per_cta < C) ? cbeg + cols_per_cta : C;

  for (long long c0 = cbeg + (long long)warp * PPT; c0 < cend; c0 += (long long)nwarps * PPT) {
    double s1[PPT], s2[PPT], mean1[PPT], mean2[PPT];
    int n1[PPT], n2[PPT];
    const uint32_t *lab[PPT];
#pragma unroll
    for (int q = 0; q < PPT; q++) {
      long long col = (c0 + q < cend) ? c0 + q : c0;  // tail columns recompute c0, not stored
      lab[q] = labels + (size_t)col * 2 * W;
      int r = (int)(col % R);
      n1[q] = n1n2[2 * r]; n2[q] = n1n2[2 * r + 1];
      s1[q] = 0.0;
    }
    // pass 1 (src/ggsea.cpp:28-38): only sum1 is consumed (mean2 = sum1 / n2, :40)
    for (int wb = 0; wb < W; wb++) {
      uint32_t m1[PPT];
#pragma unroll
      for (int q = 0; q < PPT; q++) m1[q] = __ldg(lab[q] + wb);
      const int i0 = wb * 32, jn = (n - i0 < 32) ? n - i0 : 32;
      for (int j = 0; j < jn; j++) {
        double xv = SMEM_TILE ? xs[(i0 + j) * 32 + lane] : xg[i0 + j];
#pragma unroll
        for (int q = 0; q < PPT; q++) {
          if (m1[q] & 1u) s1[q] = __dadd_rn(s1[q], xv);
          m1[q] >>= 1;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < PPT; q++) {
      mean1[q] = __ddiv_rn(s1[q], (double)n1[q]);
      mean2[q] = __ddiv_rn(s1[q], (double)n2[q]);  // sic
      s1[q] = 0.0; s2[q] = 0.0;
    }
    // pass 2 (:44-54)
    for (int wb = 0; wb < W; wb++) {
      uint32_t m1[PPT], m2[PPT];
#pragma unroll
      for (int q = 0; q < PPT; q++) { m1[q] = __ldg(lab[q] + wb); m2[q] = __ldg(lab[q] + W + wb); }
      const int i0 = wb * 32, jn = (n - i0 < 32) ? n - i0 : 32;
      for (int j = 0; j < jn; j++) {
        double xv = SMEM_TILE ? xs[(i0 + j) * 32 + lane] : xg[i0 + j];
#pragma unroll
        for (int q = 0; q < PPT; q++) {
          if (m1[q] & 1u) s1[q] = __dadd_rn(s1[q], __dsub_rn(xv, mean1[q]));
          else if (m2[q] & 1u) s2[q] = __dadd_rn(s2[q], __dsub_rn(xv, mean2[q]));
          m1[q] >>= 1; m2[q] >>= 1;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < PPT; q++) {
      mean1[q] = __dadd_rn(mean1[q], __ddiv_rn(s1[q], (double)n1[q]));
      mean2[q] = __dadd_rn(mean2[q], __ddiv_rn(s2[q], (double)n2[q]));
      s1[q] = 0.0; s2[q] = 0.0;
    }
    // pass 3 (:58-68)
    for (int wb = 0; wb < W; wb++) {
      uint32_t m1[PPT], m2[PPT];
#pragma unroll
      for (int q = 0; q < PPT; q++) { m1[q] = __ldg(lab[q] + wb); m2[q] = __ldg(lab[q] + W + wb); }
      const int i0 = wb * 32, jn = (n - i0 < 32) ? n - i0 : 32;
      for (int j = 0; j < jn; j++) {
        double xv = SMEM_TILE ? xs[(i0 + j) * 32 + lane] : xg[i0 + j];
#pragma unroll
        for (int q = 0; q < PPT; q++) {
          if (m1[q] & 1u) { double m = __dsub_rn(xv, mean1[q]); s1[q] = __dadd_rn(s1[q], __dmul_rn(m, m)); }
          else if (m2[q] & 1u) { double m = __dsub_rn(xv, mean2[q]); s2[q] = __dadd_rn(s2[q], __dmul_rn(m, m)); }
          m1[q] >>= 1; m2[q] >>= 1;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < PPT; q++) {
      if (c0 + q < cend && gvalid) {
        double var1 = __ddiv_rn(s1[q], (double)n1[q]);  // :69-70
        double var2 = __ddiv_rn(s2[q], (double)n2[q]);
        double coef = __ddiv_rn(__dsub_rn(mean1[q], mean2[q]),
                                __dadd_rn(__dsqrt_rn(var1), __dsqrt_rn(var2)));  // :71
        scores[(size_t)(c0 + q) * G + g] = coef;
        if (!isfinite(coef)) {
          int bit = isnan(coef) ? FLAG_NAN : (coef > 0 ? FLAG_POS_INF : FLAG_NEG_INF);
          atomicOr(flags + (c0 + q) / R, bit);
        }
      }
    }
  }
}

int launch_s2n(fgs_ctx *c, const double *d_x, int n, int G, const uint32_t *d_labels,
               const int *d_n1n2, int R, long long C, double *d_scores, int *d_flags) {
  if (C <= 0 || G <= 0) return FGS_OK;
  const int W = (n + 31) / 32;
  const int tiles = ceil_div(G, 32);
  const size_t smem = (size_t)n * 32 * sizeof(double);
  const bool tile_fits = smem <= (size_t)c->smem_optin;
  const int nwarps = S2N_THREADS / 32;
  // enough CTAs for ~2 waves at 4 CTAs/SM, but every CTA keeps >= 1 full sweep
  long long want_y = ((long long)c->num_sms * 8 + tiles - 1) / tiles;
  long long max_y = (C + (long long)nwarps * S2N_PPT - 1) / ((long long)nwarps * S2N_PPT);
  long long gy = want_y < 1 ? 1 : (want_y > max_y ? max_y : want_y);
  if (gy > 65535) gy = 65535;
  long long cols_per_cta = (C + gy - 1) / gy;
  cols_per_cta = ((cols_per_cta + S2N_PPT - 1) / S2N_PPT) * S2N_PPT;
  gy = (C + cols_per_cta - 1) / cols_per_cta;
  dim3 grid(tiles, (unsigned)gy);
  if (tile_fits) {
    auto k = s2n_kernel<S2N_PPT, true>;
    FGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, S2N_THREADS, smem, c->stream>>>(d_x, n, G, d_labels, d_n1n2, R, C, W, cols_per_cta,
                                              d_scores, d_flags);
  } else {
    s2n_kernel<S2N_PPT, false><<<grid, S2N_THREADS, 0, c->stream>>>(
        d_x, n, G, d_labels, d_n1n2, R, C, W, cols_per_cta, d_scores, d_flags);
  }
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// +-Inf patch of flexgsea_s2n (R/flexgsea.R:715-716) for the (rare)
// permutations whose flag word has an Inf bit: +Inf -> 1.1 * max(finite), then
// -Inf -> 1.1 * min(finite), both over the whole [G x R] score matrix of that
// permutation.  One CTA per permutation; unflagged permutations exit at once.
__global__ void s2n_patch_kernel(double *__restrict__ scores, int G, int R, int *__restrict__ flags) {
  const int p = blockIdx.x;
  const int f = flags[p];
  if (!(f & (FLAG_POS_INF | FLAG_NEG_INF))) return;
  __shared__ double red[32];
  __shared__ double bcast;
  double *m = scores + (size_t)p * R * G;
  const long long len = (long long)G * R;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int phase = 0; phase < 2; phase++) {
    const bool want = phase == 0 ? (f & FLAG_POS_INF) : (f & FLAG_NEG_INF);
    if (!want) continue;  // uniform
    double v = phase == 0 ? -INFINITY : INFINITY;
    for (long long i = threadIdx.x; i < len; i += blockDim.x) {
      double s = m[i];
      if (isfinite(s)) v = phase == 0 ? fmax(v, s) : fmin(v, s);
    }
    for (int o = 16; o; o >>= 1) {
      double t = __shfl_xor_sync(0xffffffffu, v, o);
      v = phase == 0 ? fmax(v, t) : fmin(v, t);
    }
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = red[0];
      for (int w = 1; w < nw; w++) a = phase == 0 ? fmax(a, red[w]) : fmin(a, red[w]);
      bcast = a * 1.1;
    }
    __syncthreads();
    const double rep = bcast;
    for (long long i = threadIdx.x; i < len; i += blockDim.x) {
      double s = m[i];
      if (isinf(s) && (phase == 0 ? s > 0 : s < 0)) m[i] = rep;
    }
    __syncthreads();
    // a patch value that is itself infinite (no finite score at all) stays an error
    if (threadIdx.x == 0 && !isfinite(rep)) atomicOr(flags + p, FLAG_NAN);
  }
  if (threadIdx.x == 0) atomicAnd(flags + p, ~(FLAG_POS_INF | FLAG_NEG_INF));
}

int launch_s2n_patch(fgs_ctx *c, double *d_scores, int G, int R, int P, int *d_flags) {
  if (P <= 0) return FGS_OK;
  s2n_patch_kernel<<<P, 256, 0, c->stream>>>(d_scores, G, R, d_flags);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// any non-finite value in v raises FLAG_NAN in flags[i / per_flag]
__global__ void check_finite_kernel(const double *__restrict__ v, long long len, long long per_flag,
                                    int *__restrict__ flags) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < len; i += stride)
    if (!isfinite(v[i])) atomicOr(flags + i / per_flag, FLAG_NAN);
}
int launch_check_finite(fgs_ctx *c, const double *d_v, long long len, long long per_flag,
                        int *d_flags) {
  if (len <= 0) return FGS_OK;
  int blocks = (int)((len + 255) / 256);
  if (blocks > c->num_sms * 16) blocks = c->num_sms * 16;
  check_finite_kernel<<<blocks, 256, 0, c->stream>>>(d_v, len, per_flag, d_flags);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

__global__ void abs_kernel(double *__restrict__ v, long long len) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < len; i += stride) v[i] = fabs(v[i]);
}
int launch_abs_inplace(fgs_ctx *c, double *d_v, long long len) {
  if (len <= 0) return FGS_OK;
  int blocks = (int)((len + 255) / 256);
  if (blocks > c->num_sms * 16) blocks = c->num_sms * 16;
  abs_kernel<<<blocks, 256, 0, c->stream>>>(d_v, len);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

}  // namespace fgs
