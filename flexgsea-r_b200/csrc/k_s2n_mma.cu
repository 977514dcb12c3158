// K2: signal-to-noise gene scores as an FP64 tensor-core contraction, with the
// bookkeeping that keeps the RANKS derived from them bit-exact.
//
// north_star (2): "class sums and sums of squares give s2n".  For every gene g
// and score column c (a permuted two-class labelling without NA):
//     S1 = sum_i xc[g,i] * z[i,c]      Q1 = sum_i xc[g,i]^2 * z[i,c]
// with xc the expression centred per gene (s2n is shift invariant; centring
// removes the cancellation in E[x^2] - mean^2) and z the 0/1 class-1 indicator.
// That is one GEMM [2G x n] . [n x C] on the FP64 tensor pipe
// (mma.sync.m8n8k4.f64 -> SASS DMMA); class 2 follows from the per-gene totals.
// Epilogue: mean_c, population variances, (mean1 - mean2) / (sd1 + sd2)
// (src/ggsea.cpp:69-71).
//
// These scores differ from the reference's three-pass arithmetic by ~1e-14
// relative -- harmless for ES (tolerance 1e-10) but not for ranks when two
// scores nearly tie.  Certification ("two modes"):
//   * the epilogue lists every (column, gene) whose variance suffered
//     cancellation (E[x^2] / var > S2N_CANCEL_LIMIT) or is not finite,
//   * the ranking kernel lists every adjacent pair of the sorted column closer
//     than the tie tolerance (TieCert, set in fgs_api.cu),
// and s2n_exact_entries_kernel recomputes exactly those entries with the
// bit-faithful arithmetic of k_s2n.cu; the affected columns are ranked again.
// Error bound of the contraction after the cancellation guard: n * u * limit
// ~ 4e-13 relative for n = 200 (plus ~1e-12 worst case from the class sums), two
// orders below the tie threshold (1e-10 relative, 1e-12 absolute), so every pair whose order could differ from the reference's is recomputed exactly.
#include "fgs_common.cuh"

namespace fgs {

// centred expression xc[g][i] = x[i + n g] - mean_g, and the per-gene totals
// tot[g] = {sum_i xc, sum_i xc^2} (double-double accumulation: exact to fp64)
__global__ void center_kernel(const double *__restrict__ x, int n, int G, double *__restrict__ xc,
                              double *__restrict__ tot) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const double *xg = x + (size_t)g * n;
  double *cg = xc + (size_t)g * n;
  double s = 0.0, e = 0.0;
  for (int i = 0; i < n; i++) {  // Kahan: the mean only has to be a good centre
    const double y = xg[i] - e, t = s + y;
    e = (t - s) - y; s = t;
  }
  const double mean = s / n;
  double t1 = 0.0, t2 = 0.0;
  for (int i = 0; i < n; i++) {
    const double d = xg[i] - mean;
    cg[i] = d;
    t1 += d;
    t2 = fma(d, d, t2);
  }
  tot[2 * g] = t1;
  tot[2 * g + 1] = t2;
}

int launch_center(fgs_ctx *c, const double *d_x, int n, int G, double *d_xc, double *d_tot) {
  center_kernel<<<ceil_div(G, 128), 128, 0, c->stream>>>(d_x, n, G, d_xc, d_tot);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

constexpr int SM_BG = 32, SM_BN = 128, SM_BK = 16, SM_PAD = 4;  // genes x columns x samples per tile
constexpr double S2N_CANCEL_LIMIT = 16.0;

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void push_entry(int2 *__restrict__ list, int *__restrict__ count, int cap,
                                           int *__restrict__ overflow, int col, int gene) {
  const int at = atomicAdd(count, 1);
  if (at < cap) list[at] = make_int2(col, gene);
  else atomicOr(overflow, 16);  // the caller reruns the block bit-faithfully
}

// 8-byte asynchronous copy global -> shared, zero-filled when `ok` is false
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int nbytes = ok ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(d), "l"(src), "r"(nbytes));
}

// CTA tile: 32 genes x 128 columns; 8 warps = 2 (gene halves of 16) x 4 (column
// groups of 32).  A warp holds, for its 16 genes and 32 columns, the S1 and the
// Q1 accumulators: m-tiles {0,1} multiply xc, m-tiles {2,3} multiply xc^2 (the
// square is taken on the A fragment, so xc is staged once).
//   A (centred expression): 32 x 16 tiles, double-buffered in shared memory by
//     cp.async, so the next tile lands while the tensor pipe works on this one.
//   B (0/1 class-1 indicator): never materialised.  A lane needs it for four
//     fixed columns only; it keeps those columns' label words in registers and
//     builds the 1.0 / 0.0 operand from a bit (two integer instructions).
__global__ void __launch_bounds__(256)
s2n_mma_kernel(const double *__restrict__ xc, const double *__restrict__ tot, int n, int G,
               const uint32_t *__restrict__ labels, const double *__restrict__ inv_n, int R, long long C,
               int W, double *__restrict__ scores, int2 *__restrict__ list, int *__restrict__ count,
               int cap, int *__restrict__ overflow) {
  __shared__ __align__(16) double As[2][SM_BG][SM_BK + SM_PAD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g0 = blockIdx.x * SM_BG;
  const long long c0 = (long long)blockIdx.y * SM_BN;
  const int wg = (warp & 1) * 16, wn = (warp >> 1) * 32;
  const int grp = lane >> 2, tig = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // this lane's B columns (operand fragment: column = grp of each 8-column tile)
  const uint32_t *lab[4];
  bool colok[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const long long col = c0 + wn + j * 8 + grp;
    colok[j] = col < C;
    lab[j] = labels + (size_t)(colok[j] ? col : 0) * 2 * W;
  }
  // staging assignment: thread -> (gene row, two samples) of the 32 x 16 tile
  const int sr = tid >> 3, sk = (tid & 7) * 2;
  const bool rowok = g0 + sr < G;
  const double *srow = xc + (size_t)(rowok ? g0 + sr : 0) * n;
  auto stage = [&](int buf, int k0) {
    cp_async8(&As[buf][sr][sk], srow + k0 + sk, rowok && k0 + sk < n);
    cp_async8(&As[buf][sr][sk + 1], srow + k0 + sk + 1, rowok && k0 + sk + 1 < n);
    asm volatile("cp.async.commit_group;\n" ::);
  };

  const int ntile = (n + SM_BK - 1) / SM_BK;
  stage(0, 0);
  uint32_t wcur[4] = {0u, 0u, 0u, 0u};
  for (int t = 0; t < ntile; t++) {
    const int k0 = t * SM_BK;
    if ((t & 1) == 0) {  // a 32-sample label word covers two 16-sample tiles
#pragma unroll
      for (int j = 0; j < 4; j++) wcur[j] = colok[j] ? __ldg(lab[j] + (k0 >> 5)) : 0u;
    }
    if (t + 1 < ntile) {
      stage((t + 1) & 1, k0 + SM_BK);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    const int buf = t & 1;
    const int bit0 = (k0 & 31) + tig;
#pragma unroll
    for (int kk = 0; kk < SM_BK; kk += 4) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; i++) a[i] = As[buf][wg + i * 8 + grp][kk + tig];
#pragma unroll
      for (int j = 0; j < 4; j++)  // 1.0 or 0.0 straight from the label bit (samples >= n have bit 0)
        b[j] = __hiloint2double((int)(((wcur[j] >> (bit0 + kk)) & 1u) * 0x3ff00000u), 0);
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const double a2 = a[i] * a[i];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
          dmma884(acc[i + 2][j][0], acc[i + 2][j][1], a2, b[j]);
        }
      }
    }
    __syncthreads();  // the buffer is refilled two iterations later
  }
  // accumulator fragment: row = grp (gene), columns 2 * tig + {0, 1}
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int g = g0 + wg + i * 8 + grp;
    if (g >= G) continue;
    const double t1 = tot[2 * g], t2 = tot[2 * g + 1];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const long long col = c0 + wn + j * 8 + 2 * tig + h;
        if (col >= C) continue;
        const int r = (int)(col % R);
        // (reciprocal class sizes, made once per response: the tensor path is certified to 1e-10, not bit-faithful)
        const double in1 = __ldg(inv_n + 2 * r), in2 = __ldg(inv_n + 2 * r + 1);
        const double s1 = acc[i][j][h], q1 = acc[i + 2][j][h];
        const double s2 = t1 - s1, q2 = t2 - q1;
        const double m1 = s1 * in1, m2 = s2 * in2;
        const double e1 = q1 * in1, e2 = q2 * in2;          // E[xc^2] per class
        const double v1 = e1 - m1 * m1, v2 = e2 - m2 * m2;  // population variances
        const double coef = (m1 - m2) / (sqrt(fmax(v1, 0.0)) + sqrt(fmax(v2, 0.0)));
        scores[(size_t)col * G + g] = coef;
        // (a score this close to zero is either an exact zero in the reference -- equal class
        // means, e.g. tied data -- or lost to cancellation in m1 - m2: an exact zero must stay
        // one, it is the weight 0 of the weighted KS)
        const bool shaky = !(v1 * S2N_CANCEL_LIMIT > e1) || !(v2 * S2N_CANCEL_LIMIT > e2) || !isfinite(coef) ||
                           fabs(coef) <= 1e-9;
        if (shaky) push_entry(list, count, cap, overflow, (int)col, g);
      }
  }
}

int launch_s2n_mma(fgs_ctx *c, const double *d_xc, const double *d_tot, int n, int G,
                   const uint32_t *d_labels, const double *d_inv_n, int R, long long C, double *d_scores,
                   int2 *d_list, int *d_count, int cap, int *d_overflow) {
  if (C <= 0 || G <= 0) return FGS_OK;
  dim3 grid(ceil_div(G, SM_BG), ceil_div(C, SM_BN));
  s2n_mma_kernel<<<grid, 256, 0, c->stream>>>(d_xc, d_tot, n, G, d_labels, d_inv_n, R, C, (n + 31) / 32,
                                              d_scores, d_list, d_count, cap, d_overflow);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// Bit-faithful s2n (the arithmetic of s2n_kernel / src/ggsea.cpp:19-73) for a
// list of (column, gene) entries; a thread per entry, grid-stride over the
// device-side count.  Writes the score and raises the same non-finite flags.
__global__ void s2n_exact_entries_kernel(const double *__restrict__ x, int n, int G,
                                         const uint32_t *__restrict__ labels,
                                         const int *__restrict__ n1n2, int R, int W,
                                         const int2 *__restrict__ list, const int *__restrict__ count,
                                         int cap, double *__restrict__ scores, int *__restrict__ flags,
                                         int after_patch, const int *__restrict__ unless) {
  if (unless && *unless != 0) return;  // the list overflowed: the whole block is rescored instead
  int total = *count;
  if (total > cap) total = cap;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int col = list[e].x, g = list[e].y;
    const double *xg = x + (size_t)g * n;
    const uint32_t *m1p = labels + (size_t)col * 2 * W, *m2p = m1p + W;
    const int r = col % R;
    const double dn1 = (double)n1n2[2 * r], dn2 = (double)n1n2[2 * r + 1];
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < n; i++)
      if ((m1p[i >> 5] >> (i & 31)) & 1u) s1 = __dadd_rn(s1, xg[i]);
    double mean1 = __ddiv_rn(s1, dn1), mean2 = __ddiv_rn(s1, dn2);  // sic, :40
    s1 = 0.0;
    for (int i = 0; i < n; i++) {
      if ((m1p[i >> 5] >> (i & 31)) & 1u) s1 = __dadd_rn(s1, __dsub_rn(xg[i], mean1));
      else if ((m2p[i >> 5] >> (i & 31)) & 1u) s2 = __dadd_rn(s2, __dsub_rn(xg[i], mean2));
    }
    mean1 = __dadd_rn(mean1, __ddiv_rn(s1, dn1));
    mean2 = __dadd_rn(mean2, __ddiv_rn(s2, dn2));
    s1 = 0.0; s2 = 0.0;
    for (int i = 0; i < n; i++) {
      if ((m1p[i >> 5] >> (i & 31)) & 1u) { const double d = __dsub_rn(xg[i], mean1); s1 = __dadd_rn(s1, __dmul_rn(d, d)); }
      else if ((m2p[i >> 5] >> (i & 31)) & 1u) { const double d = __dsub_rn(xg[i], mean2); s2 = __dadd_rn(s2, __dmul_rn(d, d)); }
    }
    const double coef = __ddiv_rn(__dsub_rn(mean1, mean2),
                                  __dadd_rn(__dsqrt_rn(__ddiv_rn(s1, dn1)), __dsqrt_rn(__ddiv_rn(s2, dn2))));
    // +-Inf is replaced by 1.1 * the permutation's largest / smallest finite score
    // (s2n_patch_kernel, R/flexgsea.R:715-716) right after the first scoring pass.  An
    // entry recomputed later (rank certification, tied ES units) that is infinite has
    // been patched then: its patched value stays
    if (after_patch && isinf(coef)) continue;
    scores[(size_t)col * G + g] = coef;
    if (!isfinite(coef)) {
      const int bit = isnan(coef) ? FLAG_NAN : (coef > 0 ? FLAG_POS_INF : FLAG_NEG_INF);
      atomicOr(flags + col / R, bit);
    }
  }
}

int launch_s2n_exact_entries(fgs_ctx *c, const double *d_x, int n, int G, const uint32_t *d_labels,
                             const int *d_n1n2, int R, const int2 *d_list, const int *d_count, int cap,
                             double *d_scores, int *d_flags, bool after_patch, const int *unless) {
  s2n_exact_entries_kernel<<<c->num_sms * 4, 128, 0, c->stream>>>(d_x, n, G, d_labels, d_n1n2, R,
                                                                  (n + 31) / 32, d_list, d_count, cap,
                                                                  d_scores, d_flags, after_patch ? 1 : 0, unless);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

}  // namespace fgs
