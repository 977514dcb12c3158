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
// columns.  The reference walks the samples once per pass and picks the class
// accumulator by a branch (:30-37, :46-53, :60-67); an accumulator therefore only
// ever sees the samples of ITS class, in index order.  Here each column's samples
// are first split into the two index lists (class 1 in index order, then class 2:
// class_lists_kernel) and every pass runs one loop per class over its list -- the
// same operation sequence per accumulator, hence the same bits, but no
// class-conditional arithmetic at all.  (Written as `if (class1) s1 += d; else s2
// += d` the compiler if-converts to "compute both, select": twice the FP64 work.
// Round 1 measured that form at 47 ms per 10 000 columns at C2.)
// FP64-pipe bound: n1 + 5n double operations per (gene, column); a visit costs one add for the
// address (the lists hold ready-made offsets), one shared-memory load and its FP64 work.
#include "fgs_common.cuh"

namespace fgs {

constexpr int S2N_THREADS = 256;
constexpr int S2N_PPT = 2;

// lists[col][0 .. n1) = samples of class 1 in index order, lists[col][n1 .. n1 + n2) = class 2
// (NA samples are in neither mask and in neither list).  A warp per column.
__global__ void __launch_bounds__(256)
class_lists_kernel(const uint32_t *__restrict__ labels, int n, int W, long long C,
                   uint16_t *__restrict__ lists, const int *__restrict__ only_if,
                   const int *__restrict__ unless) {
  if (only_if && *only_if == 0) return;
  if (unless && *unless != 0) return;
  const long long col = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (col >= C) return;
  uint16_t *out = lists + (size_t)col * n;
  int at = 0;
  for (int cls = 0; cls < 2; cls++) {
    const uint32_t *m = labels + ((size_t)col * 2 + cls) * W;
    for (int wb = 0; wb < W; wb++) {
      const uint32_t w = __ldg(m + wb);
      if ((w >> lane) & 1u) out[at + __popc(w & ((1u << lane) - 1u))] = (uint16_t)(wb * 32 + lane);
      at += __popc(w);
    }
  }
}

template <int PPT, bool SMEM_TILE>
__global__ void __launch_bounds__(S2N_THREADS)
s2n_kernel(const double *__restrict__ x, int n, int G, const uint16_t *__restrict__ lists,
           const int *__restrict__ n1n2, int R, long long C, long long cols_per_cta,
           double *__restrict__ scores, int *__restrict__ flags, const int *__restrict__ only_if,
           const int *__restrict__ unless) {
  // a conditional launch: the block-level fallback of the tensor-core path (fgs_api.cu) queues
  // this kernel behind a device flag and it leaves at once when the flag says "not needed"
  if (only_if && *only_if == 0) return;
  if (unless && *unless != 0) return;
  extern __shared__ double xs[];  // [n][32] when SMEM_TILE, then the warps' sample lists
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = blockIdx.x * 32 + lane;
  const bool gvalid = g < G;
  const double *xg = x + (size_t)(gvalid ? g : 0) * n;
  // per warp and column: the samples as ready-made offsets (bytes into the tile row of this
  // lane's gene, or elements of xg), so that a visit is one add, one load and its FP64 work
  // (each class's list starts on a 16-byte boundary: four offsets per shared-memory load)
  const int npad = ((n + 3) & ~3) + 4;
  uint32_t *sl = (uint32_t *)(xs + (SMEM_TILE ? (size_t)n * 32 : 0)) + (size_t)warp * PPT * npad;
  if (SMEM_TILE) {
    // lanes over genes: conflict-free shared stores; the strided global reads
    // hit L2 (x is resident there) and happen once per CTA
    for (int i = warp; i < n; i += nwarps) xs[i * 32 + lane] = gvalid ? xg[i] : 0.0;
    __syncthreads();
  }
  const long long cbeg = (long long)blockIdx.y * cols_per_cta;
  const long long cend = (cbeg + cols_per_cta < C) ? cbeg + cols_per_cta : C;
  const char *xrow = (const char *)(xs + lane);
  auto xat = [&](uint32_t o) -> double {
    return SMEM_TILE ? *(const double *)(xrow + o) : __ldg(xg + o);
  };

  for (long long c0 = cbeg + (long long)warp * PPT; c0 < cend; c0 += (long long)nwarps * PPT) {
    double s1[PPT], s2[PPT], mean1[PPT], mean2[PPT];
    int n1[PPT], n2[PPT], o2[PPT];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < PPT; q++) {
      const long long col = (c0 + q < cend) ? c0 + q : c0;  // tail columns recompute c0, not stored
      const int r = (int)(col % R);
      n1[q] = n1n2[2 * r]; n2[q] = n1n2[2 * r + 1];
      o2[q] = (n1[q] + 3) & ~3;
      const uint16_t *src = lists + (size_t)col * n;
      for (int i = lane; i < n1[q] + n2[q]; i += 32)
        sl[q * npad + (i < n1[q] ? i : o2[q] + (i - n1[q]))] = SMEM_TILE ? (uint32_t)src[i] * 256u : (uint32_t)src[i];
      s1[q] = 0.0;
    }
    __syncwarp();
    // The columns of a warp advance together while both have samples left (two independent
    // FP64 chains in flight), then each finishes on its own.  `OP` is the pass's update.
#define S2N_SWEEP(ACC, MEAN, OFF, CNT, OP)                                                         \
    {                                                                                              \
      int kc = CNT[0];                                                                             \
      _Pragma("unroll") for (int q = 1; q < PPT; q++) kc = min(kc, CNT[q]);                        \
      kc &= ~3;                                                                                    \
      for (int k = 0; k < kc; k += 4) {                                                            \
        uint4 o4[PPT];                                                                             \
        _Pragma("unroll") for (int q = 0; q < PPT; q++) o4[q] = *(const uint4 *)(sl + q * npad + OFF[q] + k); \
        _Pragma("unroll") for (int q = 0; q < PPT; q++) { const double xv = xat(o4[q].x); OP(ACC[q], xv, MEAN[q]); } \
        _Pragma("unroll") for (int q = 0; q < PPT; q++) { const double xv = xat(o4[q].y); OP(ACC[q], xv, MEAN[q]); } \
        _Pragma("unroll") for (int q = 0; q < PPT; q++) { const double xv = xat(o4[q].z); OP(ACC[q], xv, MEAN[q]); } \
        _Pragma("unroll") for (int q = 0; q < PPT; q++) { const double xv = xat(o4[q].w); OP(ACC[q], xv, MEAN[q]); } \
      }                                                                                            \
      _Pragma("unroll") for (int q = 0; q < PPT; q++) {                                            \
        for (int k = kc; k < CNT[q]; k++) {                                                        \
          const double xv = xat(sl[q * npad + OFF[q] + k]);                                        \
          OP(ACC[q], xv, MEAN[q]);                                                                 \
        }                                                                                          \
      }                                                                                            \
    }
#define S2N_OP_SUM(A, X, M) A = __dadd_rn(A, X)
#define S2N_OP_DEV(A, X, M) A = __dadd_rn(A, __dsub_rn(X, M))
#define S2N_OP_SQ(A, X, M) { const double d__ = __dsub_rn(X, M); A = __dadd_rn(A, __dmul_rn(d__, d__)); }
    int zero[PPT];
#pragma unroll
    for (int q = 0; q < PPT; q++) zero[q] = 0;
    // pass 1 (src/ggsea.cpp:28-38): only sum1 is consumed (mean2 = sum1 / n2, :40)
    S2N_SWEEP(s1, mean1, zero, n1, S2N_OP_SUM)
#pragma unroll
    for (int q = 0; q < PPT; q++) {
      mean1[q] = __ddiv_rn(s1[q], (double)n1[q]);
      mean2[q] = __ddiv_rn(s1[q], (double)n2[q]);  // sic
      s1[q] = 0.0; s2[q] = 0.0;
    }
    // pass 2 (:44-54)
    S2N_SWEEP(s1, mean1, zero, n1, S2N_OP_DEV)
    S2N_SWEEP(s2, mean2, o2, n2, S2N_OP_DEV)
#pragma unroll
    for (int q = 0; q < PPT; q++) {
      mean1[q] = __dadd_rn(mean1[q], __ddiv_rn(s1[q], (double)n1[q]));
      mean2[q] = __dadd_rn(mean2[q], __ddiv_rn(s2[q], (double)n2[q]));
      s1[q] = 0.0; s2[q] = 0.0;
    }
    // pass 3 (:58-68)
    S2N_SWEEP(s1, mean1, zero, n1, S2N_OP_SQ)
    S2N_SWEEP(s2, mean2, o2, n2, S2N_OP_SQ)
#undef S2N_SWEEP
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
               const int *d_n1n2, int R, long long C, double *d_scores, int *d_flags,
               const int *only_if, const int *unless) {
  if (C <= 0 || G <= 0) return FGS_OK;
  if (n > 65535) { set_error("more than 65535 samples"); return FGS_ERR_UNSUPPORTED; }
  const int W = (n + 31) / 32;
  FGS_TRY(c->class_lists.ensure((size_t)C * n));
  class_lists_kernel<<<ceil_div(C, 8), 256, 0, c->stream>>>(d_labels, n, W, C, c->class_lists.p, only_if, unless);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  const int tiles = ceil_div(G, 32);
  const int nwarps = S2N_THREADS / 32;
  const size_t smem_lists = (size_t)nwarps * S2N_PPT * (((n + 3) & ~3) + 4) * sizeof(uint32_t);
  const size_t smem = (size_t)n * 32 * sizeof(double) + smem_lists;
  const bool tile_fits = smem <= (size_t)c->smem_optin;
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
    k<<<grid, S2N_THREADS, smem, c->stream>>>(d_x, n, G, c->class_lists.p, d_n1n2, R, C, cols_per_cta,
                                              d_scores, d_flags, only_if, unless);
  } else {
    if (smem_lists > (size_t)c->smem_optin) { set_error("%d samples: the class lists do not fit on chip", n); return FGS_ERR_UNSUPPORTED; }
    FGS_CUDA(cudaFuncSetAttribute(s2n_kernel<S2N_PPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lists));
    s2n_kernel<S2N_PPT, false><<<grid, S2N_THREADS, smem_lists, c->stream>>>(
        d_x, n, G, c->class_lists.p, d_n1n2, R, C, cols_per_cta, d_scores, d_flags, only_if, unless);
  }
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// redo[0 .. C) = 1 when the device flag says so (block-level fallback: every column is ranked again)
__global__ void mark_all_if_kernel(int *__restrict__ redo, long long C, const int *__restrict__ only_if,
                                   const int *__restrict__ unless) {
  if (*only_if == 0 || (unless && *unless != 0)) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) redo[i] = 1;
}
int launch_mark_all_if(fgs_ctx *c, int *d_redo, long long C, const int *only_if, const int *unless) {
  if (C <= 0) return FGS_OK;
  mark_all_if_kernel<<<ceil_div(C, 256), 256, 0, c->stream>>>(d_redo, C, only_if, unless);
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
