// K3: segmented stable radix sort -> descending gene rank with R's tie rule.
//
// Replaces flexgsea_weighted_ks$prepare (R/flexgsea.R:927-933):
//   gene.rank = G - rank(s, 'keep', 'first') + 1
// i.e. rank 1 = largest score; among equal scores (and -0.0 == 0.0) the gene
// with the LARGER index gets the smaller descending rank.  We get exactly that
// order from a stable LSD radix sort, ascending on key = ~orderable(score),
// started from the genes in reverse index order.
//
// One CTA per score column (segment), persistent over columns.  8 passes of
// 8 bits; a pass in which every key has the same digit is skipped.  Ranking
// inside a pass is the warp-match scheme: each warp owns a contiguous slice of
// the current order, per round `__match_any_sync` groups equal digits, and a
// [digit][warp] counter table (scanned digit-major) gives stable destinations.
//
// Outputs per column: rank16[g] (1-based descending rank, uint16) and
// sabs[r] = |score|^p of the gene at rank r+1 (the weighted-KS hit weights,
// R/flexgsea.R:819, already in rank order so the ES kernel needs no second
// indirection).
#include "fgs_common.cuh"

namespace fgs {

constexpr int SORT_THREADS = 512;
constexpr int SORT_WARPS = SORT_THREADS / 32;

__device__ __forceinline__ unsigned long long score_key_desc(double s) {
  if (s == 0.0) s = 0.0;  // -0.0 ties with 0.0 in R's comparator
  unsigned long long u = (unsigned long long)__double_as_longlong(s);
  u ^= (u >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull;  // ascending-orderable
  return ~u;                                                        // descending
}
__device__ __forceinline__ double key_desc_to_score(unsigned long long k) {
  unsigned long long a = ~k;
  a ^= (a >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull;
  return __longlong_as_double((long long)a);
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_rank_kernel(const double *__restrict__ scores, int G, long long C, int abs_flag, double p_exp,
                 unsigned long long *__restrict__ key_scratch, uint32_t *__restrict__ idx_scratch,
                 uint16_t *__restrict__ rank16, double *__restrict__ sabs) {
  __shared__ uint32_t cnt[256 * SORT_WARPS];
  __shared__ uint32_t wsum[SORT_WARPS];
  __shared__ int s_skip;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const size_t Gp = pitch_of(G);
  int Lw = (G + SORT_WARPS - 1) / SORT_WARPS;
  Lw = (Lw + 31) & ~31;
  const int wbeg = warp * Lw < G ? warp * Lw : G;
  const int wend = wbeg + Lw < G ? wbeg + Lw : G;

  unsigned long long *kA = key_scratch + (size_t)blockIdx.x * 2 * G, *kB = kA + G;
  uint32_t *iA = idx_scratch + (size_t)blockIdx.x * 2 * G, *iB = iA + G;

  for (long long col = blockIdx.x; col < C; col += gridDim.x) {
    const double *sc = scores + (size_t)col * G;
    // materialise keys in reverse gene order (position q holds gene G-1-q)
    for (int q = tid; q < G; q += SORT_THREADS) {
      int g = G - 1 - q;
      double s = sc[g];
      if (abs_flag) s = fabs(s);
      kA[q] = score_key_desc(s);
      iA[q] = (uint32_t)g;
    }
    unsigned long long *kin = kA, *kout = kB;
    uint32_t *iin = iA, *iout = iB;
    __syncthreads();

    for (int pass = 0; pass < 8; pass++) {
      const int shift = pass * 8;
      for (int i = tid; i < 256 * SORT_WARPS; i += SORT_THREADS) cnt[i] = 0;
      if (tid == 0) s_skip = 0;
      __syncthreads();
      // sweep 1: per-warp digit counts
      for (int base = wbeg; base < wend; base += 32) {
        int q = base + lane;
        bool valid = q < wend;
        uint32_t d = valid ? (uint32_t)((kin[q] >> shift) & 255ull) : 256u + lane;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (valid && (peers & lt_mask) == 0) cnt[d * SORT_WARPS + warp] += __popc(peers);
        __syncwarp();
      }
      __syncthreads();
      // exclusive scan over (digit-major, warp-minor); 8 entries per thread
      {
        uint32_t v[8], tot = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { v[k] = cnt[tid * 8 + k]; tot += v[k]; }
        // a digit owns entries [16 d, 16 d + 16) = threads 2d, 2d+1
        uint32_t pair = tot + __shfl_xor_sync(0xffffffffu, tot, 1);
        if (pair == (uint32_t)G) s_skip = 1;
        uint32_t inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < warp; w++) woff += wsum[w];
        uint32_t run = woff + inc - tot;
#pragma unroll
        for (int k = 0; k < 8; k++) { cnt[tid * 8 + k] = run; run += v[k]; }
      }
      __syncthreads();
      if (s_skip) { __syncthreads(); continue; }  // all keys share this digit
      // sweep 2: stable scatter
      for (int base = wbeg; base < wend; base += 32) {
        int q = base + lane;
        bool valid = q < wend;
        unsigned long long key = valid ? kin[q] : 0ull;
        uint32_t id = valid ? iin[q] : 0u;
        uint32_t d = valid ? (uint32_t)((key >> shift) & 255ull) : 256u + lane;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t dst = 0;
        if (valid) dst = cnt[d * SORT_WARPS + warp] + __popc(peers & lt_mask);
        __syncwarp();
        if (valid && (peers & lt_mask) == 0) cnt[d * SORT_WARPS + warp] += __popc(peers);
        __syncwarp();
        if (valid) { kout[dst] = key; iout[dst] = id; }
      }
      __syncthreads();
      { unsigned long long *t = kin; kin = kout; kout = t; }
      { uint32_t *t = iin; iin = iout; iout = t; }
    }
    // emit rank (scatter by gene) and rank-ordered hit weights (coalesced)
    uint16_t *rk = rank16 + (size_t)col * Gp;
    double *sa = sabs + (size_t)col * Gp;
    for (int q = tid; q < G; q += SORT_THREADS) {
      double w = fabs(key_desc_to_score(kin[q]));
      if (p_exp != 1.0) w = pow(w, p_exp);
      sa[q] = w;
      rk[iin[q]] = (uint16_t)(q + 1);
    }
    __syncthreads();
  }
}

int launch_sort(fgs_ctx *c, const double *d_scores, int G, long long C, int abs_flag, double p_exp,
                uint16_t *d_rank16, double *d_sabs) {
  if (C <= 0 || G <= 0) return FGS_OK;
  if (G > 65535) {
    set_error("weighted KS ranking supports up to 65535 genes (got %d)", G);
    return FGS_ERR_UNSUPPORTED;
  }
  int grid = c->num_sms * 3;
  if (grid > C) grid = (int)C;
  FGS_TRY(c->sort_keys.ensure((size_t)grid * 2 * G));
  FGS_TRY(c->sort_idx.ensure((size_t)grid * 2 * G));
  sort_rank_kernel<<<grid, SORT_THREADS, 0, c->stream>>>(d_scores, G, C, abs_flag, p_exp,
                                                         c->sort_keys.p, c->sort_idx.p, d_rank16,
                                                         d_sabs);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

}  // namespace fgs
