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
// inside a pass is the warp-peer scheme: each warp owns a contiguous slice of
// the current order, per round eight ballots group the lanes with equal digits, and a
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
    if (gap == 0.0 && tc.dup && tc.dup[ga] == tc.dup[gb]) return;  // duplicated rows: tied in the reference too
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

// lanes holding the same 8-bit digit, by eight ballots (one per digit bit).
// __match_any_sync gives the same mask but serialises over the distinct values
// in the warp (~30 of them for random digits), which made it the bottleneck.
__device__ __forceinline__ uint32_t digit_peers(uint32_t d) {
  uint32_t peers = 0xffffffffu;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t bit = (d >> i) & 1u;
    const uint32_t b = __ballot_sync(0xffffffffu, bit);
    peers &= b ^ (bit - 1u);  // bit set: lanes with the bit; clear: lanes without
  }
  return peers;
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
        uint32_t d = valid ? (uint32_t)((kin[q] >> shift) & 255ull) : 0u;
        uint32_t peers = digit_peers(d) & __ballot_sync(0xffffffffu, valid);
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
        uint32_t d = valid ? (uint32_t)((key >> shift) & 255ull) : 0u;
        uint32_t peers = digit_peers(d) & __ballot_sync(0xffffffffu, valid);
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
      w = hit_weight(w, p_exp);
      sa[q] = w;
      rk[iin[q]] = (uint16_t)(q + 1);
      if (tc.list && q + 1 < G) certify_pair(tc, col, sq, key_desc_to_score(kin[q + 1]), iin[q], iin[q + 1]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Shared-memory variant (G <= 29696): the column is sorted on chip by the TOP
// 32 bits of the key -- 6 bytes per gene in shared memory (uint32 key + uint16
// gene), 8 stable passes of 4 bits with per-thread byte counters (no warp votes),
// items staged in registers between the read and the scatter of a pass -- and
// the (few, short) runs of equal 32-bit
// keys are then put in exact 64-bit order by an insertion sort on the low key
// halves.  Exact ties need no work: the stable sort from reverse gene order
// already left them in R's order.  A column with a run longer than
// SORT_MAX_RUN is flagged and redone by the 64-bit kernel above.
// ---------------------------------------------------------------------------
constexpr int SORT32_THREADS = 1024;
constexpr int SORT32_WARPS = 32;
constexpr int SORT32_MAX_G = 29 * 1024;  // 6 bytes per gene + 48 KB of counters in 227 KB
constexpr int SORT_MAX_RUN = 48;

// The on-chip sort proper, shared by sort_rank32_kernel (a whole column) and
// sort_rank_big_kernel (a score-range bucket of a long column): n (key, gene) pairs
// in skey / sidx, pads behind them, are put in ascending key order.
template <int IPT>
__device__ __forceinline__ void sort32_core(uint32_t *__restrict__ skey, uint16_t *__restrict__ sidx,
                                            unsigned char *__restrict__ cnt8, uint16_t *__restrict__ pre16,
                                            uint32_t *__restrict__ wsum, int *__restrict__ s_redo,
                                            const double *__restrict__ sc, int n, int abs_flag) {
  constexpr int CROW = SORT32_THREADS + 16;  // bytes per cnt8 row
  constexpr int PROW = SORT32_THREADS + 8;   // uint16 per pre16 row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 8 stable passes of 4 bits.  Thread t owns the IPT consecutive positions
  // [t*IPT, (t+1)*IPT) of the current order (IPT is odd: conflict-free) and a
  // private byte counter per digit, so ranking needs no warp votes at all:
  // destination = prefix over (digit, thread) + arrival order inside the thread.
  for (int pass = 0; pass < 8; pass++) {
    const int shift = pass * 4;
    for (int i = tid; i < 16 * CROW / 4; i += SORT32_THREADS) ((uint32_t *)cnt8)[i] = 0;
    uint32_t k[IPT], id2[(IPT + 1) / 2];
#pragma unroll
    for (int i = 0; i < IPT; i++) k[i] = skey[tid * IPT + i];
#pragma unroll
    for (int i = 0; i < IPT; i += 2) {
      const uint32_t a0 = sidx[tid * IPT + i];
      const uint32_t b0 = (i + 1 < IPT) ? sidx[tid * IPT + i + 1] : 0u;
      id2[i >> 1] = a0 | (b0 << 16);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IPT; i++) cnt8[((k[i] >> shift) & 15u) * CROW + tid]++;
    __syncthreads();
    {  // exclusive scan over (digit-major, thread-minor): 16 byte counters per thread
      // thread t scans the 16 entries [16t, 16t + 16) of the digit-major order: row t / 64
      const uint4 w = ((const uint4 *)cnt8)[(tid >> 6) * (CROW / 16) + (tid & 63)];
      const uint32_t tot = __dp4a(w.x, 0x01010101u, __dp4a(w.y, 0x01010101u,
                            __dp4a(w.z, 0x01010101u, __dp4a(w.w, 0x01010101u, 0u))));
      uint32_t inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      uint32_t run = inc - tot;
      for (int w2 = 0; w2 < warp; w2++) run += wsum[w2];
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
      uint32_t outw[8];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const uint32_t c0 = (ws[j >> 2] >> (8 * (j & 3))) & 255u, c1 = (ws[j >> 2] >> (8 * ((j + 1) & 3))) & 255u;
        outw[j >> 1] = run | ((run + c0) << 16);
        run += c0 + c1;
      }
      uint4 *dst = (uint4 *)(pre16 + (tid >> 6) * PROW + 16 * (tid & 63));
      dst[0] = make_uint4(outw[0], outw[1], outw[2], outw[3]);
      dst[1] = make_uint4(outw[4], outw[5], outw[6], outw[7]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      uint16_t *slot = pre16 + ((k[i] >> shift) & 15u) * PROW + tid;
      const uint32_t dst = *slot;
      *slot = (uint16_t)(dst + 1);
      skey[dst] = k[i];
      sidx[dst] = (uint16_t)(id2[i >> 1] >> (16 * (i & 1)));
    }
    __syncthreads();
  }

  // exact order inside runs of equal 32-bit keys (by the low key halves, stable)
#pragma unroll 1
  for (int i = 0; i < IPT; i++) {
    const int q = tid + i * SORT32_THREADS;
    if (q + 1 < n) {
      const uint32_t kq = skey[q];
      if (skey[q + 1] == kq && (q == 0 || skey[q - 1] != kq)) {  // q starts a run
        int e = q + 2;
        while (e < n && skey[e] == kq) e++;
        if (e - q > SORT_MAX_RUN) {
          *s_redo = 1;
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
}

template <int IPT>
__global__ void __launch_bounds__(SORT32_THREADS, 1)
sort_rank32_kernel(const double *__restrict__ scores, int G, long long C, int abs_flag, double p_exp,
                   uint16_t *__restrict__ rank16, double *__restrict__ sabs, int *__restrict__ redo,
                   const int *__restrict__ only_cols, TieCert tc) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int GS = SORT32_THREADS * IPT;  // padded capacity
  uint32_t *skey = (uint32_t *)smem;
  uint16_t *sidx = (uint16_t *)(skey + GS);
  // per pass: cnt8[digit][thread] (16 rows of 1024 bytes) -> pre16[digit][thread] (exclusive
  // prefixes).  Rows are 16 bytes longer than their 1024 entries: without the skew the four
  // lanes that share a counter word's bank collide whenever their digits differ.
  constexpr int CROW = SORT32_THREADS + 16;  // bytes per cnt8 row
  constexpr int PROW = SORT32_THREADS + 8;   // uint16 per pre16 row
  unsigned char *cnt8 = (unsigned char *)(sidx + GS + (GS & 1));
  uint16_t *pre16 = (uint16_t *)(cnt8 + 16 * CROW);
  __shared__ uint32_t wsum[SORT32_WARPS];
  __shared__ int s_redo;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t Gp = pitch_of(G);

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

    sort32_core<IPT>(skey, sidx, cnt8, pre16, wsum, &s_redo, sc, G, abs_flag);
    __syncthreads();
    if (s_redo) {
      if (tid == 0) redo[col] = 1;
    } else {
      uint16_t *rk = rank16 + (size_t)col * Gp;
      double *sa = sabs + (size_t)col * Gp;
      // the scores come back from L2 by gene (a gather): several independent loads in
      // flight per thread.  (Streaming the column through shared memory in coalesced
      // slices and gathering on chip was measured: no faster.)
      constexpr int OB = 4;
#pragma unroll 1
      for (int i0 = 0; i0 < IPT; i0 += OB) {
        uint32_t gq[OB];
        double sv[OB];
#pragma unroll
        for (int u = 0; u < OB; u++) {
          const int q = tid + (i0 + u) * SORT32_THREADS;
          gq[u] = (i0 + u < IPT && q < G) ? (uint32_t)sidx[q] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < OB; u++) sv[u] = (gq[u] != 0xffffffffu) ? __ldg(sc + gq[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < OB; u++) {
          if (gq[u] == 0xffffffffu) continue;
          const int q = tid + (i0 + u) * SORT32_THREADS;
          const uint32_t g = gq[u];
          double sq = sv[u];
          if (abs_flag) sq = fabs(sq);
          double w = fabs(sq);
          w = hit_weight(w, p_exp);
          sa[q] = w;
          // a neighbour whose 32-bit key differs by >= 2 is > 1e-6 away in relative
          // terms; only closer keys (or scores so small that the absolute tolerance
          // could bite) need the exact comparison
          if (tc.list && q + 1 < G && (skey[q + 1] - skey[q] <= 1u || fabs(sq) <= 1e-5)) {
            const uint32_t g2 = sidx[q + 1];
            double s2 = __ldg(sc + g2);
            if (abs_flag) s2 = fabs(s2);
            certify_pair(tc, col, sq, s2, g, g2);
          }
        }
      }
      // ranks: scattered by gene into the (now free) key array, then written out as
      // one coalesced row -- a 2-byte store per gene straight to L2 is a transaction each
      __syncthreads();
      uint16_t *rks = (uint16_t *)skey;
#pragma unroll 1
      for (int i = 0; i < IPT; i++) {
        const int q = tid + i * SORT32_THREADS;
        if (q < G) rks[sidx[q]] = (uint16_t)(q + 1);
      }
      __syncthreads();
      for (int i = tid; i < (int)(Gp / 8); i += SORT32_THREADS) ((uint4 *)rk)[i] = ((const uint4 *)rks)[i];
    }
    __syncthreads();
  }
}

template <int IPT>
static int launch_sort32_t(fgs_ctx *c, const double *d_scores, int G, long long C, int abs_flag,
                           double p_exp, uint16_t *d_rank16, double *d_sabs, int *d_redo, int grid,
                           const int *only_cols, const TieCert &tc) {
  const size_t smem = (size_t)SORT32_THREADS * IPT * 6 + 2 + (size_t)(SORT32_THREADS + 16) * 16 * 3;
  auto k = sort_rank32_kernel<IPT>;
  FGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, SORT32_THREADS, smem, c->stream>>>(d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, d_redo,
                                               only_cols, tc);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// ---------------------------------------------------------------------------
// Long columns (29 696 < G <= 65 534) do not fit the on-chip sort in one piece.
// They are first split by SCORE RANGE into K buckets of about G / K genes, each
// of which does: (a) a histogram of the top 16 key bits (65 536 bins of packed
// uint16 counts in shared memory) places K - 1 cuts at bin boundaries, (b) a stable
// partition (reverse gene order kept, so the tie rule survives) writes the gene
// lists of the buckets, (c) every bucket is sorted on chip by sort32_core and its
// ranks / hit weights are written behind those of the buckets before it.
// A bucket larger than the kernel's capacity (one bin holding a quarter of all
// genes) or a run of > SORT_MAX_RUN equal 32-bit keys sends the column to the
// 64-bit kernel.  Replaces 8 passes through L2 (3 ms per 60 000-gene column) by
// two coalesced reads of the column and K on-chip sorts.
constexpr int BIG_MAXK = 4;
template <int IPT>
__global__ void __launch_bounds__(SORT32_THREADS, 1)
sort_rank_big_kernel(const double *__restrict__ scores, int G, long long C, int abs_flag, double p_exp,
                     uint16_t *__restrict__ rank16, double *__restrict__ sabs, int *__restrict__ redo,
                     const int *__restrict__ only_cols, TieCert tc, uint16_t *__restrict__ glist_all,
                     int K) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int GS = SORT32_THREADS * IPT;
  constexpr int CROW = SORT32_THREADS + 16;
  // sort phase
  uint32_t *skey = (uint32_t *)smem;
  uint16_t *sidx = (uint16_t *)(skey + GS);
  unsigned char *cnt8 = (unsigned char *)(sidx + GS + (GS & 1));
  uint16_t *pre16 = (uint16_t *)(cnt8 + 16 * CROW);
  // histogram phase: 32 768 words of two uint16 counts, then 1 024 per-thread prefixes
  uint32_t *hist32 = (uint32_t *)smem;
  uint32_t *tpre = hist32 + 32768;
  // partition phase: bucket of every position, then the gene lists
  unsigned char *bid = smem;
  uint16_t *lst = (uint16_t *)(smem + 65536);
  __shared__ uint32_t wsum[SORT32_WARPS];
  __shared__ int s_redo;
  __shared__ int s_cut[BIG_MAXK + 1], s_start[BIG_MAXK + 1];
  __shared__ uint32_t s_bsum[BIG_MAXK][SORT32_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t Gp = pitch_of(G);
  uint16_t *glist = glist_all + (size_t)blockIdx.x * Gp;

  for (long long col = blockIdx.x; col < C; col += gridDim.x) {
    if (only_cols && !only_cols[col]) continue;  // uniform per CTA
    const double *sc = scores + (size_t)col * G;
    uint16_t *rk = rank16 + (size_t)col * Gp;
    double *sa = sabs + (size_t)col * Gp;
    __syncthreads();
    // (a) histogram of the top 16 key bits
    for (int i = tid; i < 32768; i += SORT32_THREADS) hist32[i] = 0;
    if (tid == 0) s_redo = 0;
    __syncthreads();
    for (int g = tid; g < G; g += SORT32_THREADS) {
      double v = __ldg(sc + g);
      if (abs_flag) v = fabs(v);
      const uint32_t k16 = (uint32_t)(score_key_desc(v) >> 48);
      atomicAdd(&hist32[k16 >> 1], 1u << (16 * (k16 & 1u)));  // G <= 65 534: a half never carries
    }
    __syncthreads();
    {  // exclusive prefix of the 64 bins of every thread
      uint32_t tot = 0;
      for (int i = 0; i < 32; i++) { const uint32_t w = hist32[tid * 32 + i]; tot += (w & 0xffffu) + (w >> 16); }
      uint32_t inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      uint32_t run = inc - tot;
      for (int w2 = 0; w2 < warp; w2++) run += wsum[w2];
      tpre[tid] = run;
    }
    __syncthreads();
    if (tid == 0) {  // cut i = the first bin boundary with at least i * G / K keys below it
      s_cut[0] = 0; s_start[0] = 0;
      for (int i = 1; i < K; i++) {
        const uint32_t target = (uint32_t)(((long long)i * G) / K);
        int lo = 0, hi = SORT32_THREADS - 1;  // last thread whose prefix is < target
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tpre[mid] < target) lo = mid; else hi = mid - 1; }
        uint32_t cum = tpre[lo];
        int b = lo * 64;
        while (cum < target && b < (lo + 1) * 64) {
          const uint32_t w = hist32[b >> 1];
          cum += (b & 1) ? (w >> 16) : (w & 0xffffu);
          b++;
        }
        s_cut[i] = b; s_start[i] = (int)cum;
      }
      s_cut[K] = 65536; s_start[K] = G;
      for (int i = 0; i < K; i++) if (s_start[i + 1] - s_start[i] > GS) s_redo = 1;
    }
    __syncthreads();
    if (s_redo) {  // a bucket would not fit: the 64-bit kernel takes the column
      if (tid == 0) redo[col] = 1;
      continue;
    }
    int cut[BIG_MAXK + 1], start[BIG_MAXK + 1];
#pragma unroll
    for (int i = 0; i <= BIG_MAXK; i++) { cut[i] = i <= K ? s_cut[i] : 65536; start[i] = i <= K ? s_start[i] : G; }
    __syncthreads();
    // (b) stable partition of the positions (position q holds gene G-1-q: reverse order)
    for (int q = tid; q < G; q += SORT32_THREADS) {
      double v = __ldg(sc + (G - 1 - q));
      if (abs_flag) v = fabs(v);
      const int k16 = (int)(score_key_desc(v) >> 48);
      int b = 0;
#pragma unroll
      for (int i = 1; i < BIG_MAXK; i++) b += (i < K && k16 >= cut[i]);
      bid[q] = (unsigned char)b;
    }
    __syncthreads();
    const int L = (G + SORT32_THREADS - 1) / SORT32_THREADS;
    const int q0 = min(G, tid * L), q1 = min(G, q0 + L);
    uint32_t cnt[BIG_MAXK] = {0u, 0u, 0u, 0u};
    for (int q = q0; q < q1; q++) {
      const int b = bid[q];
#pragma unroll
      for (int i = 0; i < BIG_MAXK; i++) cnt[i] += (b == i);
    }
    uint32_t offs[BIG_MAXK];
#pragma unroll
    for (int i = 0; i < BIG_MAXK; i++) {  // exclusive prefix over the threads, per bucket
      uint32_t inc = cnt[i];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) s_bsum[i][warp] = inc;
      offs[i] = inc - cnt[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < BIG_MAXK; i++) {
      uint32_t run = (uint32_t)start[i] + offs[i];
      for (int w2 = 0; w2 < warp; w2++) run += s_bsum[i][w2];
      offs[i] = run;
    }
    for (int q = q0; q < q1; q++) {
      const int b = bid[q];
      uint32_t dst = 0;
#pragma unroll
      for (int i = 0; i < BIG_MAXK; i++) if (b == i) dst = offs[i]++;
      lst[dst] = (uint16_t)(G - 1 - q);
    }
    __syncthreads();
    for (int i = tid; i < (int)(Gp / 8); i += SORT32_THREADS) ((uint4 *)glist)[i] = ((const uint4 *)lst)[i];
    __syncthreads();
    // (c) the buckets, in score order
    bool failed = false;
    double edge_s = 0.0;   // last score of the bucket before (thread 0)
    uint32_t edge_g = 0u;
    for (int b = 0; b < K && !failed; b++) {
      const int base = start[b], nb = start[b + 1] - base;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < IPT; i++) {
        const int q = tid + i * SORT32_THREADS;
        uint32_t key = 0xffffffffu, g = 0xffffu;
        if (q < nb) {
          g = glist[base + q];
          double v = __ldg(sc + g);
          if (abs_flag) v = fabs(v);
          key = (uint32_t)(score_key_desc(v) >> 32);
        }
        skey[q] = key;
        sidx[q] = (uint16_t)g;
      }
      __syncthreads();
      sort32_core<IPT>(skey, sidx, cnt8, pre16, wsum, &s_redo, sc, nb, abs_flag);
      __syncthreads();
      if (s_redo) { failed = true; break; }
      constexpr int OB = 4;
#pragma unroll 1
      for (int i0 = 0; i0 < IPT; i0 += OB) {
        uint32_t gq[OB];
        double sv[OB];
#pragma unroll
        for (int u = 0; u < OB; u++) {
          const int q = tid + (i0 + u) * SORT32_THREADS;
          gq[u] = (i0 + u < IPT && q < nb) ? (uint32_t)sidx[q] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < OB; u++) sv[u] = (gq[u] != 0xffffffffu) ? __ldg(sc + gq[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < OB; u++) {
          if (gq[u] == 0xffffffffu) continue;
          const int q = tid + (i0 + u) * SORT32_THREADS;
          const uint32_t g = gq[u];
          double sq = sv[u];
          if (abs_flag) sq = fabs(sq);
          double w = fabs(sq);
          w = hit_weight(w, p_exp);
          sa[base + q] = w;
          rk[g] = (uint16_t)(base + q + 1);
          if (tc.list && q + 1 < nb && (skey[q + 1] - skey[q] <= 1u || fabs(sq) <= 1e-5)) {
            const uint32_t g2 = sidx[q + 1];
            double s2 = __ldg(sc + g2);
            if (abs_flag) s2 = fabs(s2);
            certify_pair(tc, col, sq, s2, g, g2);
          }
        }
      }
      if (tid == 0 && nb > 0) {  // the pair that straddles the cut is adjacent in the final order too
        double f = __ldg(sc + sidx[0]), l = __ldg(sc + sidx[nb - 1]);
        if (abs_flag) { f = fabs(f); l = fabs(l); }
        if (tc.list && base > 0) certify_pair(tc, col, edge_s, f, edge_g, sidx[0]);
        edge_s = l; edge_g = sidx[nb - 1];
      }
    }
    __syncthreads();
    if (failed && tid == 0) redo[col] = 1;
  }
}
template <int IPT>
static int launch_sort_big_t(fgs_ctx *c, const double *d_scores, int G, long long C, int abs_flag,
                             double p_exp, uint16_t *d_rank16, double *d_sabs, int *d_redo, int grid,
                             const int *only_cols, const TieCert &tc, int K) {
  size_t smem = (size_t)SORT32_THREADS * IPT * 6 + 2 + (size_t)(SORT32_THREADS + 16) * 16 * 3;
  const size_t need = 65536 + (size_t)pitch_of(G) * 2 + 64;  // partition phase; histogram phase: 132 KB
  if (smem < need) smem = need;
  if (smem < 32768 * 4 + 4096 + 64) smem = 32768 * 4 + 4096 + 64;
  auto k = sort_rank_big_kernel<IPT>;
  FGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, SORT32_THREADS, smem, c->stream>>>(d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, d_redo,
                                               only_cols, tc, c->sort_glist.p, K);
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
  if (G <= SORT32_MAX_G && !c->opt_sort64) {
    int grid = c->num_sms;
    if (grid > C) grid = (int)C;
    FGS_TRY(c->sort_redo.ensure((size_t)C + 8));
    int *redo = c->sort_redo.p;
    FGS_CUDA(cudaMemsetAsync(redo, 0, (size_t)C * sizeof(int), c->stream));
    const int ipt = (G + SORT32_THREADS - 1) / SORT32_THREADS;
#define FGS_SORT32(N) FGS_TRY(launch_sort32_t<N>(c, d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, redo, grid, only_cols, tc))
    if (ipt <= 5) FGS_SORT32(5);
    else if (ipt <= 9) FGS_SORT32(9);
    else if (ipt <= 13) FGS_SORT32(13);
    else if (ipt <= 17) FGS_SORT32(17);
    else if (ipt <= 21) FGS_SORT32(21);
    else if (ipt <= 25) FGS_SORT32(25);
    else FGS_SORT32(29);
#undef FGS_SORT32
    only = redo;  // the 64-bit kernel below only redoes flagged columns
  } else if (G <= 65534 && !c->opt_sort64) {
    // long columns: K score-range buckets of about G / K genes, each sorted on chip
    int grid = c->num_sms;
    if (grid > C) grid = (int)C;
    FGS_TRY(c->sort_redo.ensure((size_t)C + 8));
    int *redo = c->sort_redo.p;
    FGS_CUDA(cudaMemsetAsync(redo, 0, (size_t)C * sizeof(int), c->stream));
    FGS_TRY(c->sort_glist.ensure((size_t)grid * pitch_of(G)));
    const int K = (G + 19999) / 20000;                     // 2 .. 4
    const int want = (int)(1.2 * (double)((G + K - 1) / K));  // room for the bin that straddles a cut
    const int ipt = (want + SORT32_THREADS - 1) / SORT32_THREADS;
#define FGS_SORTBIG(N) FGS_TRY(launch_sort_big_t<N>(c, d_scores, G, C, abs_flag, p_exp, d_rank16, d_sabs, redo, grid, only_cols, tc, K))
    if (ipt <= 21) FGS_SORTBIG(21);
    else if (ipt <= 25) FGS_SORTBIG(25);
    else FGS_SORTBIG(29);
#undef FGS_SORTBIG
    only = redo;
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
