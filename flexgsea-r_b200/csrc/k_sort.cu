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

// adjacent scores of the sorted column closer than this are listed for an exact
// recomputation (k_s2n_mma.cu); a column with such a pair is ranked again
__device__ __forceinline__ void certify_pair(const TieCert &tc, long long col, double a, double b,
                                             uint32_t ga, uint32_t gb) {
  const double gap = fabs(a - b), scale = fmax(fabs(a), fabs(b));
  if (gap <= tc.abs_tol || gap <= tc.rel_tol * scale) {
    const int at = atomicAdd(tc.count, 2);
    if (at + 1 < tc.cap) {
      tc.list[at] = make_int2((int)col, (int)ga);
      tc.list[at + 1] = make_int2((int)col, (int)gb);
    } else {
      atomicOr(tc.overflow, 16);
    }
    tc.redo[col] = 1;
  }
}

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
                 uint16_t *__restrict__ rank16, double *__restrict__ sabs,
                 const int *__restrict__ only_flagged, TieCert tc) {
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
    if (only_flagged && !only_flagged[col]) continue;  // uniform per CTA
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
      const double sq = key_desc_to_score(kin[q]);
      double w = fabs(sq);
      if (p_exp != 1.0) w = pow(w, p_exp);
      sa[q] = w;
      rk[iin[q]] = (uint16_t)(q + 1);
      if (tc.list && q + 1 < G) certify_pair(tc, col, sq, key_desc_to_score(kin[q + 1]), iin[q], iin[q + 1]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Shared-memory variant (G <= 32768): the column is sorted on chip by the TOP
// 32 bits of the key -- 4 radix passes instead of 8, 6 bytes per gene in shared
// memory (uint32 key + uint16 gene), items staged in registers between the
// read and the scatter of a pass -- and the (few, short) runs of equal 32-bit
// keys are then put in exact 64-bit order by an insertion sort on the low key
// halves.  Exact ties need no work: the stable sort from reverse gene order
// already left them in R's order.  A column with a run longer than
// SORT_MAX_RUN is flagged and redone by the 64-bit kernel above.
// ---------------------------------------------------------------------------
constexpr int SORT32_THREADS = 1024;
constexpr int SORT32_WARPS = 32;
constexpr int SORT32_CSTRIDE = 33;  // counter row stride: (digit + warp) mod 32 banks
constexpr int SORT_MAX_RUN = 48;

template <int IPT>
__global__ void __launch_bounds__(SORT32_THREADS, 1)
sort_rank32_kernel(const double *__restrict__ scores, int G, long long C, int abs_flag, double p_exp,
                   uint16_t *__restrict__ rank16, double *__restrict__ sabs, int *__restrict__ redo,
                   const int *__restrict__ only_cols, TieCert tc) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int GS = SORT32_THREADS * IPT;  // padded capacity
  uint32_t *skey = (uint32_t *)smem;
  uint16_t *sidx = (uint16_t *)(skey + GS);
  uint32_t *cnt = (uint32_t *)(sidx + GS);
  __shared__ uint32_t wsum[SORT32_WARPS];
  __shared__ int s_redo;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const size_t Gp = pitch_of(G);
  const int wbase = warp * (32 * IPT);  // this warp's slice of the current order

  for (long long col = blockIdx.x; col < C; col += gridDim.x) {
    if (only_cols && !only_cols[col]) continue;  // uniform per CTA
    const double *sc = scores + (size_t)col * G;
    if (tid == 0) s_redo = 0;
    // position q holds gene G-1-q (reverse order: the stable sort then breaks ties
    // by descending gene index, R's rule); pads sort last
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      const int q = tid + i * SORT32_THREADS;
      uint32_t k = 0xffffffffu, g = 0xffffu;
      if (q < G) {
        g = (uint32_t)(G - 1 - q);
        double v = __ldg(sc + g);
        if (abs_flag) v = fabs(v);
        k = (uint32_t)(score_key_desc(v) >> 32);
      }
      skey[q] = k;
      sidx[q] = (uint16_t)g;
    }
    __syncthreads();

    for (int pass = 0; pass < 4; pass++) {
      const int shift = pass * 8;
      for (int i = tid; i < 256 * SORT32_CSTRIDE; i += SORT32_THREADS) cnt[i] = 0;
      uint32_t k[IPT];
#pragma unroll
      for (int r = 0; r < IPT; r++) k[r] = skey[wbase + r * 32 + lane];
      __syncthreads();
      // sweep 1: per-warp digit counts
#pragma unroll
      for (int r = 0; r < IPT; r++) {
        const uint32_t d = (k[r] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if ((peers & lt_mask) == 0) cnt[d * SORT32_CSTRIDE + warp] += __popc(peers);
        __syncwarp();
      }
      __syncthreads();
      {  // exclusive scan over (digit-major, warp-minor): 8 entries per thread
        const int d = tid >> 2, w0 = (tid & 3) * 8;
        uint32_t v[8], tot = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { v[j] = cnt[d * SORT32_CSTRIDE + w0 + j]; tot += v[j]; }
        uint32_t inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < warp; w++) woff += wsum[w];
        uint32_t run = woff + inc - tot;
#pragma unroll
        for (int j = 0; j < 8; j++) { cnt[d * SORT32_CSTRIDE + w0 + j] = run; run += v[j]; }
      }
      // the gene ids travel with the keys: read them all before anyone scatters
      uint32_t id2[(IPT + 1) / 2];
#pragma unroll
      for (int r = 0; r < IPT; r += 2) {
        const uint32_t a = sidx[wbase + r * 32 + lane];
        const uint32_t b = (r + 1 < IPT) ? sidx[wbase + (r + 1) * 32 + lane] : 0u;
        id2[r >> 1] = a | (b << 16);
      }
      __syncthreads();
      // sweep 2: stable scatter
#pragma unroll
      for (int r = 0; r < IPT; r++) {
        const uint32_t d = (k[r] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t dst = cnt[d * SORT32_CSTRIDE + warp] + __popc(peers & lt_mask);
        __syncwarp();
        if ((peers & lt_mask) == 0) cnt[d * SORT32_CSTRIDE + warp] += __popc(peers);
        __syncwarp();
        skey[dst] = k[r];
        sidx[dst] = (uint16_t)(id2[r >> 1] >> (16 * (r & 1)));
      }
      __syncthreads();
    }

    // exact order inside runs of equal 32-bit keys (by the low key halves, stable)
#pragma unroll 1
    for (int i = 0; i < IPT; i++) {
      const int q = tid + i * SORT32_THREADS;
      if (q + 1 < G) {
        const uint32_t kq = skey[q];
        if (skey[q + 1] == kq && (q == 0 || skey[q - 1] != kq)) {  // q starts a run
          int e = q + 2;
          while (e < G && skey[e] == kq) e++;
          if (e - q > SORT_MAX_RUN) {
            s_redo = 1;
          } else {
            for (int a = q + 1; a < e; a++) {  // insertion sort, strict compare keeps ties stable
              const uint16_t ga = sidx[a];
              double va = __ldg(sc + ga);
              if (abs_flag) va = fabs(va);
              const uint32_t la = (uint32_t)score_key_desc(va);
              int b = a - 1;
              while (b >= q) {
                double vb = __ldg(sc + sidx[b]);
                if (abs_flag) vb = fabs(vb);
                if ((uint32_t)score_key_desc(vb) <= la) break;
                sidx[b + 1] = sidx[b];
                b--;
              }
              sidx[b + 1] = ga;
            }
          }
        }
      }
    }
    __syncthreads();
    if (s_redo) {
      if (tid == 0) redo[col] = 1;
    } else {
      uint16_t *rk = rank16 + (size_t)col * Gp;
      double *sa = sabs + (size_t)col * Gp;
#pragma unroll 1
      for (int i = 0; i < IPT; i++) {
        const int q = tid + i * SORT32_THREADS;
        if (q < G) {
          const uint32_t g = sidx[q];
          double sq = sc[g];
          if (abs_flag) sq = fabs(sq);
          double w = fabs(sq);
          if (p_exp != 1.0) w = pow(w, p_exp);
          sa[q] = w;
          rk[g] = (uint16_t)(q + 1);
          // a neighbour whose 32-bit key differs by >= 2 is > 1e-6 away in relative
          // terms; only closer keys (or scores so small that the absolute tolerance
          // could bite) need the exact comparison
          if (tc.list && q + 1 < G && (skey[q + 1] - skey[q] <= 1u || fabs(sq) <= 1e-5)) {
            const uint32_t g2 = sidx[q + 1];
            double s2 = sc[g2];
            if (abs_flag) s2 = fabs(s2);
            certify_pair(tc, col, sq, s2, g, g2);
          }
        }
      }
    }
    __syncthreads();
  }
}

template <int IPT>
static int launch_sort32_t(fgs_ctx *c, const double *d_scores, int G, long long C, int abs_flag,
                           double p_exp, uint16_t *d_rank16, double *d_sabs, int *d_redo, int grid,
                           const int *only_cols, const TieCert &tc) {
  const size_t smem = (size_t)SORT32_THREADS * IPT * 6 + (size_t)256 * SORT32_CSTRIDE * 4;
  auto k = sort_rank32_kernel<IPT>;
  FGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, SORT32_THREADS, smem, c->stream>>>(d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, d_redo,
                                               only_cols, tc);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

int launch_sort(fgs_ctx *c, const double *d_scores, int G, long long C, int abs_flag, double p_exp,
                uint16_t *d_rank16, double *d_sabs, const int *only_cols, const TieCert *cert) {
  TieCert tc{};
  if (cert) tc = *cert;
  if (C <= 0 || G <= 0) return FGS_OK;
  if (G > 65535) {
    set_error("weighted KS ranking supports up to 65535 genes (got %d)", G);
    return FGS_ERR_UNSUPPORTED;
  }
  const int *only = nullptr;
  if (G <= 32768 && !c->opt_sort64) {
    int grid = c->num_sms;
    if (grid > C) grid = (int)C;
    FGS_TRY(c->sort_redo.ensure((size_t)C + 8));
    int *redo = c->sort_redo.p;
    FGS_CUDA(cudaMemsetAsync(redo, 0, (size_t)C * sizeof(int), c->stream));
    const int ipt = (G + SORT32_THREADS - 1) / SORT32_THREADS;
    if (ipt <= 4) FGS_TRY(launch_sort32_t<4>(c, d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, redo, grid, only_cols, tc));
    else if (ipt <= 8) FGS_TRY(launch_sort32_t<8>(c, d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, redo, grid, only_cols, tc));
    else if (ipt <= 12) FGS_TRY(launch_sort32_t<12>(c, d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, redo, grid, only_cols, tc));
    else if (ipt <= 16) FGS_TRY(launch_sort32_t<16>(c, d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, redo, grid, only_cols, tc));
    else if (ipt <= 20) FGS_TRY(launch_sort32_t<20>(c, d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, redo, grid, only_cols, tc));
    else if (ipt <= 24) FGS_TRY(launch_sort32_t<24>(c, d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, redo, grid, only_cols, tc));
    else if (ipt <= 28) FGS_TRY(launch_sort32_t<28>(c, d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, redo, grid, only_cols, tc));
    else FGS_TRY(launch_sort32_t<32>(c, d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, redo, grid, only_cols, tc));
    only = redo;  // the 64-bit kernel below only redoes flagged columns
  } else {
    only = only_cols;
  }
  int grid = c->num_sms * 3;
  if (grid > C) grid = (int)C;
  FGS_TRY(c->sort_keys.ensure((size_t)grid * 2 * G));
  FGS_TRY(c->sort_idx.ensure((size_t)grid * 2 * G));
  sort_rank_kernel<<<grid, SORT_THREADS, 0, c->stream>>>(d_scores, G, C, abs_flag, p_exp,
                                                         c->sort_keys.p, c->sort_idx.p, d_rank16,
                                                         d_sabs, only, tc);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

}  // namespace fgs
