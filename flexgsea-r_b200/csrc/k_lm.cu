// K2 (lm): regression gene score as a dense FP64 contraction.
//
// Replaces flexgsea_lm (R/flexgsea.R:647-676): lm.fit(design_perm, x)
// coefficients (intercept dropped, :667-669) divided by sd(x[,g]) (:671).
// Permuting the rows of the design leaves its QR factor R unchanged and
// permutes the rows of Q, so with H = R^-1 Q^T (computed once on the host in
// fp64 Householder form) the coefficients of permutation p are
//     coef_p[c, g] = sum_i H[c, perm_p(i)] * x[i, g]
// i.e. one GEMM  X^T [G x n] . B [n x (R P)]  with B[i, (p, c)] = H[c, perm_p(i)],
// followed by the 1/sd epilogue.  This is the tensor-core stage of the path:
// the inner product runs on the FP64 tensor pipe (mma.sync m8n8k4 f64 -> SASS
// DMMA; tcgen05 has no f64 kind), tiles staged in shared memory.
#include "fgs_common.cuh"

namespace fgs {

// B[i + n * (p * R + c)] = H[c * n + perm[i + n * p] - 1]
__global__ void lm_build_B_kernel(const double *__restrict__ H, const int *__restrict__ perm, int n,
                                  int R, int P, double *__restrict__ B, int *__restrict__ bad) {
  const long long total = (long long)n * R * P;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % n);
    const long long col = t / n;
    const int c = (int)(col % R);
    const long long p = col / R;
    int src = perm ? perm[(size_t)p * n + i] - 1 : i;
    if (src < 0 || src >= n) { atomicOr(bad, 8); src = 0; }
    B[t] = H[(size_t)c * n + src];
  }
}

int launch_lm_build_B(fgs_ctx *c, const double *d_H, const int *d_perm, int n, int R, int P,
                      double *d_B) {
  long long total = (long long)n * R * P;
  if (total <= 0) return FGS_OK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > c->num_sms * 16) blocks = c->num_sms * 16;
  lm_build_B_kernel<<<blocks, 256, 0, c->stream>>>(d_H, d_perm, n, R, P, d_B,
                                                   c->flags.p + (c->flags.n - 1));
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// sd(x[, g]) with R's two-pass mean (cov.c); sums in double-double so the
// result rounds like R's long-double accumulation.
__device__ __forceinline__ void dd_acc(double &hi, double &lo, double b) {
  double s = __dadd_rn(hi, b);
  double bb = __dsub_rn(s, hi);
  double e = __dadd_rn(__dsub_rn(hi, __dsub_rn(s, bb)), __dsub_rn(b, bb));
  e = __dadd_rn(e, lo);
  hi = __dadd_rn(s, e);
  lo = __dsub_rn(e, __dsub_rn(hi, s));
}
__global__ void gene_sd_kernel(const double *__restrict__ x, int n, int G, double *__restrict__ sd) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const double *xg = x + (size_t)g * n;
  double hi = 0, lo = 0;
  for (int i = 0; i < n; i++) dd_acc(hi, lo, xg[i]);
  double mean = (hi + lo) / n;
  double h2 = 0, l2 = 0;
  for (int i = 0; i < n; i++) dd_acc(h2, l2, xg[i] - mean);
  mean += (h2 + l2) / n;
  double h3 = 0, l3 = 0;
  for (int i = 0; i < n; i++) { double d = xg[i] - mean; dd_acc(h3, l3, d * d); }
  sd[g] = sqrt((h3 + l3) / (n - 1));
}
// 64-bit hash of every gene's expression row (bit pattern of its n doubles): groups candidate
// duplicated rows for fgs_set_x, which confirms them on the host by comparing the rows
__global__ void row_hash_kernel(const double *__restrict__ x, int n, int G, unsigned long long *__restrict__ hash) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const unsigned long long *row = (const unsigned long long *)(x + (size_t)g * n);
  unsigned long long h = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < n; i++) {
    h ^= row[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
  }
  hash[g] = h;
}
int launch_row_hash(fgs_ctx *c, const double *d_x, int n, int G, unsigned long long *d_hash) {
  row_hash_kernel<<<ceil_div(G, 128), 128, 0, c->stream>>>(d_x, n, G, d_hash);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

int launch_gene_sd(fgs_ctx *c, const double *d_x, int n, int G, double *d_sd) {
  gene_sd_kernel<<<ceil_div(G, 128), 128, 0, c->stream>>>(d_x, n, G, d_sd);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// ---------------------------------------------------------------------------
// DMMA GEMM: scores[col][g] = (sum_i x[i, g] * B[i, col]) / sd[g]
//   A = x viewed as [G x n] row-major (x is [n x G] column-major: a gene's
//       samples contiguous)            -> mma "row" operand
//   B = [n x C] column-major (a column's n values contiguous) -> mma "col"
// CTA tile 64 genes x 64 columns, K step 16; 8 warps, each warp 32 x 16 of the
// tile = 4 x 2 m8n8k4 accumulators.  Shared tiles padded to kill bank
// conflicts on the fragment loads.
// ---------------------------------------------------------------------------
constexpr int LM_BM = 64, LM_BN = 64, LM_BK = 16, LM_PAD = 4;  // row stride 20: conflict-free f64 fragment loads

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// 8-byte asynchronous copy global -> shared, zero-filled when `ok` is false
__device__ __forceinline__ void lm_cp_async8(void *dst_smem, const void *src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int nbytes = ok ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(d), "l"(src), "r"(nbytes));
}

__global__ void __launch_bounds__(256)
lm_gemm_kernel(const double *__restrict__ x, int n, int G, const double *__restrict__ B, long long C,
               const double *__restrict__ sd, int R, double *__restrict__ scores,
               int *__restrict__ flags) {
  // both operand tiles are double-buffered by cp.async: the next 16-sample slab lands while
  // the tensor pipe works on this one
  __shared__ __align__(16) double As[2][LM_BM][LM_BK + LM_PAD];  // [gene][k]
  __shared__ __align__(16) double Bs[2][LM_BN][LM_BK + LM_PAD];  // [col][k]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g0 = blockIdx.x * LM_BM;
  const long long c0 = (long long)blockIdx.y * LM_BN;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16;  // warp tile origin
  const int grp = lane >> 2, tig = lane & 3;              // mma thread coords
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // 64 x 16 doubles per tile = 1024 elements, 4 per thread and matrix; k is contiguous in
  // both sources
  auto stage = [&](int buf, int k0) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = tid + u * 256;
      const int r = e / LM_BK, k = e % LM_BK;
      const int g = g0 + r;
      const long long col = c0 + r;
      const bool ka = k0 + k < n;
      lm_cp_async8(&As[buf][r][k], x + (size_t)(g < G ? g : 0) * n + (ka ? k0 + k : 0), g < G && ka);
      lm_cp_async8(&Bs[buf][r][k], B + (size_t)(col < C ? col : 0) * n + (ka ? k0 + k : 0), col < C && ka);
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  const int ntile = (n + LM_BK - 1) / LM_BK;
  stage(0, 0);
  for (int t = 0; t < ntile; t++) {
    if (t + 1 < ntile) {
      stage((t + 1) & 1, (t + 1) * LM_BK);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    const int buf = t & 1;
#pragma unroll
    for (int kk = 0; kk < LM_BK; kk += 4) {
      double a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[buf][wm + i * 8 + grp][kk + tig];
#pragma unroll
      for (int j = 0; j < 2; j++) b[j] = Bs[buf][wn + j * 8 + grp][kk + tig];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();  // the buffer is refilled two iterations later
  }
  // accumulator fragment: row = grp, cols = 2 * tig + {0, 1}
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int g = g0 + wm + i * 8 + grp;
    if (g >= G) continue;
    const double s = sd[g];
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const long long col = c0 + wn + j * 8 + 2 * tig + h;
        if (col < C) {
          const double v = acc[i][j][h] / s;  // :671
          scores[(size_t)col * G + g] = v;
          if (!isfinite(v)) atomicOr(flags + col / R, FLAG_NAN);
        }
      }
  }
}

int launch_lm_gemm(fgs_ctx *c, const double *d_x, int n, int G, const double *d_B, long long C,
                   const double *d_gene_sd, double *d_scores, int *d_flags, int R) {
  if (C <= 0 || G <= 0) return FGS_OK;
  dim3 grid(ceil_div(G, LM_BM), ceil_div(C, LM_BN));
  lm_gemm_kernel<<<grid, 256, 0, c->stream>>>(d_x, n, G, d_B, C, d_gene_sd, R, d_scores, d_flags);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

}  // namespace fgs
