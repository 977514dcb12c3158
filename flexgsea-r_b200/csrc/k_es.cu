// K4 / K5: enrichment scores per (gene set, score column).
//
// K4 replaces the fast branch of flexgsea_weighted_ks_ (R/flexgsea.R:816-840):
//   w = |s[set]|^p / sum(.);  o = order(rank[set]);  d_hit = cumsum(w[o]);
//   d_miss = (rank[set][o] - 1..m) / (G - m);  pos = d_hit - d_miss;
//   neg = pos - w[o];  es = max(pos) > -min(neg) ? max(pos) : min(neg)
// K5 replaces flexgsea_maxmean_ (:725-741) and flexgsea_mean_ (:755-767).
//
// Weighted KS runs as two persistent kernels per block of columns (one CTA per
// SM, a column at a time, a warp per gene set), split so that each phase keeps
// only what it needs in shared memory and 32 warps fit beside it:
//  A  es_order_kernel: the column's rank array (uint16) is staged on chip.  A
//     set's members are ordered by rank WITHOUT a comparison sort: every member
//     sets one bit of a per-warp bitmap over rank space; lane l owns a
//     contiguous (odd-strided, bank-conflict-free) range of bitmap words, so a
//     popcount prefix over lanes gives each hit its ordinal and a walk over the
//     set bits emits the ranks in order (uint16 lists, 2 nnz bytes per column).
//     Bits are set with plain byte read-modify-writes; the rare same-byte
//     collisions inside one warp instruction are found by a read-back and
//     repaired.
//  B  es_scan_kernel: the column's rank-ordered hit weights (fp64) are staged on
//     chip; the running sum over a set's ordered hits is a blocked warp scan.
//
// Sets with duplicated members (same rank twice cannot live in a bitmap) or
// more than ES_LISTCAP members go through es_wks_generic_kernel, an
// O(m^2 / 32) warp kernel that needs no ordering at all: max/min over the
// members of pos/neg is order independent once each member knows its ordinal
// and prefix weight, which it gets by comparing against every other member.
#include "fgs_common.cuh"

namespace fgs {

constexpr int ES_LISTCAP = 512;  // max members per set in the bitmap kernel (16 per lane)

__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
  for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- phase A: order every set's members by rank (no weights needed) ----------
// Persistent CTA per column, up to 32 warps, a warp per set.  Shared memory:
// the column's rank array (uint16), and per warp a rank-space bitmap plus a
// per-word exclusive popcount prefix.  Output: the set's 0-based ranks in
// increasing order, uint16, at lists[col][loff ..).
//
// Per set, with NB = batches of 32 members (compile-time class 1/2/4/8/16):
//  1. every member sets its bit (byte read-modify-write; bits lost to same-byte
//     collisions inside one warp instruction are found by a read-back and
//     repaired),
//  2. lane l popcounts its own contiguous word range (uint4 loads; the range
//     stride is 4*odd words, so those are bank-conflict free), a warp scan turns
//     the lane totals into bases, and the exclusive prefix of every word is
//     stored,
//  3. every member reads its word's prefix, adds the popcount of the lower bits
//     of its word -- that is its ordinal -- and stores its rank at that slot.
// No comparison sort, no divergent bit walk.
template <int NB>
__device__ __forceinline__ void order_one_set(const uint16_t *__restrict__ s_rank, uint32_t *bm,
                                              uint16_t *wpre, int lane, int wpl,
                                              const int *__restrict__ set_idx, int off, int m,
                                              uint16_t *__restrict__ out) {
  unsigned char *bm8 = (unsigned char *)bm;
  uint32_t rv[NB];
#pragma unroll
  for (int b = 0; b < NB; b++) {  // member ids first: independent loads in flight
    const int j = b * 32 + lane;
    rv[b] = (j < m) ? (uint32_t)__ldg(set_idx + off + j) : 0xffffffffu;
  }
#pragma unroll
  for (int b = 0; b < NB; b++) {
    if (rv[b] != 0xffffffffu) {
      const uint32_t r0 = (uint32_t)s_rank[rv[b]] - 1u;
      const unsigned char v = bm8[r0 >> 3];
      bm8[r0 >> 3] = v | (unsigned char)(1u << (r0 & 7));
      rv[b] = r0;
    }
  }
  __syncwarp();
  for (;;) {
    bool miss = false;
#pragma unroll
    for (int b = 0; b < NB; b++) {
      const uint32_t r0 = rv[b];
      if (r0 != 0xffffffffu) {
        const unsigned char v = bm8[r0 >> 3], bit = (unsigned char)(1u << (r0 & 7));
        if (!(v & bit)) { miss = true; bm8[r0 >> 3] = v | bit; }
      }
    }
    __syncwarp();
    if (!__any_sync(0xffffffffu, miss)) break;
  }
  // per-word exclusive prefix over the lane's range, then the lane bases
  const int wpl4 = wpl >> 2;
  uint4 *my4 = (uint4 *)(bm + lane * wpl);
  uint2 *pre2 = (uint2 *)(wpre + lane * wpl);
  int cnt = 0;
  for (int i = 0; i < wpl4; i++) {
    const uint4 v = my4[i];
    const int c0 = cnt, c1 = c0 + __popc(v.x), c2 = c1 + __popc(v.y), c3 = c2 + __popc(v.z);
    cnt = c3 + __popc(v.w);
    pre2[i] = make_uint2((uint32_t)c0 | ((uint32_t)c1 << 16), (uint32_t)c2 | ((uint32_t)c3 << 16));
  }
  int inc = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  const int base = inc - cnt;  // hits owned by lower lanes
  // every lane needs every owner's base: broadcast through the prefix table
  const uint32_t base2 = (uint32_t)base * 0x00010001u;
  for (int i = 0; i < wpl4; i++) {
    uint2 v = pre2[i];
    v.x += base2; v.y += base2;  // no carry between halves: totals < 65536
    pre2[i] = v;
  }
  __syncwarp();
#pragma unroll
  for (int b = 0; b < NB; b++) {
    const uint32_t r0 = rv[b];
    if (r0 != 0xffffffffu) {
      const uint32_t wd = r0 >> 5;
      const int ord = (int)wpre[wd] + __popc(bm[wd] & ((1u << (r0 & 31)) - 1u));
      out[ord] = (uint16_t)r0;
    }
  }
  __syncwarp();
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int i = 0; i < wpl4; i++) my4[i] = z;
}

__global__ void __launch_bounds__(1024, 1)
es_order_kernel(const uint16_t *__restrict__ rank16, int G, long long C,
                const int4 *__restrict__ set_info, const int *__restrict__ set_idx, int S, int wpl,
                uint16_t *__restrict__ lists, long long Lp) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const size_t Gp = pitch_of(G);
  uint16_t *s_rank = (uint16_t *)smem;
  uint32_t *bm_all = (uint32_t *)(smem + Gp * 2);
  uint16_t *pre_all = (uint16_t *)(bm_all + (size_t)nwarps * 32 * wpl);
  uint32_t *bm = bm_all + (size_t)warp * 32 * wpl;
  uint16_t *wpre = pre_all + (size_t)warp * 32 * wpl;

  for (int i = tid; i < nwarps * 32 * wpl; i += blockDim.x) bm_all[i] = 0;

  for (long long col = blockIdx.x; col < C; col += gridDim.x) {
    __syncthreads();
    {
      const uint4 *src = (const uint4 *)(rank16 + (size_t)col * Gp);
      uint4 *dst = (uint4 *)s_rank;
      for (int i = tid; i < (int)(Gp / 8); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    uint16_t *lcol = lists + (size_t)col * Lp;
    for (int k = warp; k < S; k += nwarps) {
      const int4 info = __ldg(set_info + k);  // {member offset, m | slow << 31, list offset, set id}
      if (info.y < 0) continue;               // es_wks_generic_kernel does these
      const int off = info.x, m = info.y;
      uint16_t *out = lcol + info.z;
      if (m <= 32) order_one_set<1>(s_rank, bm, wpre, lane, wpl, set_idx, off, m, out);
      else if (m <= 64) order_one_set<2>(s_rank, bm, wpre, lane, wpl, set_idx, off, m, out);
      else if (m <= 128) order_one_set<4>(s_rank, bm, wpre, lane, wpl, set_idx, off, m, out);
      else if (m <= 256) order_one_set<8>(s_rank, bm, wpre, lane, wpl, set_idx, off, m, out);
      else order_one_set<16>(s_rank, bm, wpre, lane, wpl, set_idx, off, m, out);
    }
  }
}

// ---- phase B: running sum over the ordered hits ------------------------------
// Persistent CTA per column, a warp per set.  Shared memory: the column's
// rank-ordered hit weights (fp64).  Lane l takes hits [l*NB, l*NB + NB) (one
// vector load of its uint16 ranks): local sums, one warp scan, then the running
// maxima / minima.
template <int NB>
__device__ __forceinline__ void load_ranks(const uint16_t *__restrict__ p, uint32_t (&rr)[NB]) {
  if constexpr (NB == 1) {
    rr[0] = *p;
  } else if constexpr (NB == 2) {
    const uint32_t v = *(const uint32_t *)p;
    rr[0] = v & 0xffffu; rr[1] = v >> 16;
  } else if constexpr (NB == 4) {
    const uint2 v = *(const uint2 *)p;
    rr[0] = v.x & 0xffffu; rr[1] = v.x >> 16; rr[2] = v.y & 0xffffu; rr[3] = v.y >> 16;
  } else {
#pragma unroll
    for (int q = 0; q < NB / 8; q++) {
      const uint4 v = *(const uint4 *)(p + 8 * q);
      rr[8 * q + 0] = v.x & 0xffffu; rr[8 * q + 1] = v.x >> 16;
      rr[8 * q + 2] = v.y & 0xffffu; rr[8 * q + 3] = v.y >> 16;
      rr[8 * q + 4] = v.z & 0xffffu; rr[8 * q + 5] = v.z >> 16;
      rr[8 * q + 6] = v.w & 0xffffu; rr[8 * q + 7] = v.w >> 16;
    }
  }
}

template <int NB>
__device__ __forceinline__ double scan_one_set(const double *__restrict__ wsrc,
                                               const uint16_t *__restrict__ list, int lane, int m,
                                               int G) {
  const int k0 = lane * NB;
  uint32_t rr[NB];
  load_ranks<NB>(list + k0, rr);
  double wv[NB];
  double lsum = 0.0;
#pragma unroll
  for (int t = 0; t < NB; t++) {
    const double w = (k0 + t < m) ? wsrc[rr[t]] : 0.0;
    wv[t] = w;
    lsum += w;
  }
  double incw = lsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incw, d);
    if (lane >= d) incw += t;
  }
  const double W = __shfl_sync(0xffffffffu, incw, 31);
  double cum = incw - lsum;  // exclusive prefix of the lane (exact: incw = prefix + lsum)
  const double invW = 1.0 / W, invGm = 1.0 / (double)(G - m);
  double mx = -INFINITY, mn = INFINITY;
#pragma unroll
  for (int t = 0; t < NB; t++) {
    const int kk = k0 + t;
    if (kk < m) {
      cum += wv[t];
      const double ps = cum * invW - (double)((int)rr[t] - kk) * invGm;  // r - k with r = rr+1, k = kk+1
      const double ng = ps - wv[t] * invW;
      mx = fmax(mx, ps);
      mn = fmin(mn, ng);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  double es = (mx > -mn) ? mx : mn;                      // strict, :833
  if (W == 0.0 || m == G || m == 0) es = nan("");        // 0/0 in the reference
  return es;
}

constexpr int ES_SCAN_THREADS = 640;  // 20 warps: up to 102 registers each, no spills
template <bool W_SMEM>
__global__ void __launch_bounds__(ES_SCAN_THREADS, 1)
es_scan_kernel(const double *__restrict__ sabs, int G, long long C,
               const int4 *__restrict__ set_info, int S, const uint16_t *__restrict__ lists,
               long long Lp, double *__restrict__ es_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const size_t Gp = pitch_of(G);
  double *s_abs = (double *)smem;

  for (long long col = blockIdx.x; col < C; col += gridDim.x) {
    __syncthreads();
    if (W_SMEM) {
      const double2 *ws = (const double2 *)(sabs + (size_t)col * Gp);
      double2 *wd = (double2 *)s_abs;
      for (int i = tid; i < (int)(Gp / 2); i += blockDim.x) wd[i] = __ldg(ws + i);
    }
    __syncthreads();
    const double *wsrc = W_SMEM ? s_abs : sabs + (size_t)col * Gp;
    const uint16_t *lcol = lists + (size_t)col * Lp;
    double *es_col = es_out + (size_t)col * S;

    for (int k = warp; k < S; k += nwarps) {
      const int4 info = __ldg(set_info + k);
      if (info.y < 0) continue;
      const int m = info.y;
      const uint16_t *list = lcol + info.z;
      double es;
      if (m <= 32) es = scan_one_set<1>(wsrc, list, lane, m, G);
      else if (m <= 64) es = scan_one_set<2>(wsrc, list, lane, m, G);
      else if (m <= 128) es = scan_one_set<4>(wsrc, list, lane, m, G);
      else if (m <= 256) es = scan_one_set<8>(wsrc, list, lane, m, G);
      else es = scan_one_set<16>(wsrc, list, lane, m, G);
      if (lane == 0) es_col[info.w] = es;
    }
  }
}

// Order-free weighted KS for sets the bitmap kernel cannot take (duplicates,
// > ES_LISTCAP members).  One warp per (set, column).  Ties in rank (duplicate
// members) are ordered by member position, like R's stable order() (:822).
__global__ void __launch_bounds__(256)
es_wks_generic_kernel(const uint16_t *__restrict__ rank16, const double *__restrict__ sabs, int G,
                      long long C, const int *__restrict__ set_off, const int *__restrict__ set_idx,
                      const int *__restrict__ set_ids, int n_ids, int S, double *__restrict__ es_out) {
  const int lane = threadIdx.x & 31;
  const long long unit = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (unit >= (long long)n_ids * C) return;
  const long long col = unit / n_ids;
  const int s = set_ids ? set_ids[unit % n_ids] : (int)(unit % n_ids);
  const size_t Gp = pitch_of(G);
  const uint16_t *rk = rank16 + (size_t)col * Gp;
  const double *sa = sabs + (size_t)col * Gp;
  const int off = set_off[s], m = set_off[s + 1] - off;
  double lsum = 0.0;
  for (int j = lane; j < m; j += 32) lsum += sa[rk[set_idx[off + j]] - 1];
  const double W = warp_sum(lsum);
  const double invW = 1.0 / W, invGm = 1.0 / (double)(G - m);
  double mx = -INFINITY, mn = INFINITY;
  for (int j = lane; j < m; j += 32) {
    const int rj = rk[set_idx[off + j]];
    const double wj = sa[rj - 1];
    int kord = 0;
    double cum = 0.0;
    for (int t = 0; t < m; t++) {
      const int rt = rk[set_idx[off + t]];
      if (rt < rj || (rt == rj && t <= j)) { kord++; cum += sa[rt - 1]; }
    }
    const double ps = cum * invW - (double)(rj - kord) * invGm;
    const double ng = ps - wj * invW;
    mx = fmax(mx, ps);
    mn = fmin(mn, ng);
  }
  mx = warp_max(mx);
  mn = warp_min(mn);
  if (lane == 0) {
    double es = (mx > -mn) ? mx : mn;
    if (W == 0.0 || m == G || m == 0) es = nan("");
    es_out[(size_t)col * S + s] = es;
  }
}

// K5: mean / maxmean.  A warp per (set, column); scores read through L1/L2
// (a column is 8 G bytes and is re-read by every set of that column).
__global__ void __launch_bounds__(256)
es_mean_kernel(const double *__restrict__ scores, int G, long long C,
               const int *__restrict__ set_off, const int *__restrict__ set_idx, int S, int es_kind,
               int abs_flag, double *__restrict__ es_out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const long long nunits = (long long)S * C;
  for (long long unit = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); unit < nunits;
       unit += (long long)gridDim.x * wpb) {
    const long long col = unit / S;
    const int s = (int)(unit % S);
    const double *sc = scores + (size_t)col * G;
    const int off = set_off[s], m = set_off[s + 1] - off;
    double sp = 0.0, sn = 0.0;
    for (int j = lane; j < m; j += 32) {
      double v = sc[set_idx[off + j]];
      if (abs_flag) v = fabs(v);
      if (es_kind == FGS_ES_MEAN) sp += v;
      else { double a = fabs(v); sp += (v + a) / 2; sn += (-v + a) / 2; }  // :734-735
    }
    sp = warp_sum(sp);
    sn = warp_sum(sn);
    if (lane == 0) {
      double es;
      if (es_kind == FGS_ES_MEAN) es = sp / m;                 // :763
      else {
        double ps = sp / m, ng = sn / m;
        es = fmax(ps, ng);                                     // :736
        if (isnan(ps) || isnan(ng)) es = nan("");
        else if (ng > ps) es = -es;                            // :737, strict
      }
      es_out[(size_t)col * S + s] = es;
    }
  }
}

int launch_es_wks_generic(fgs_ctx *c, const uint16_t *d_rank16, const double *d_sabs, int G,
                          long long C, const int *d_set_ids, int n_ids, double *d_es) {
  if (C <= 0 || n_ids <= 0) return FGS_OK;
  long long units = (long long)n_ids * C;
  long long blocks = (units + 7) / 8;
  if (blocks > 0x7fffffffLL) { set_error("too many (set, column) units"); return FGS_ERR_ARG; }
  es_wks_generic_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(
      d_rank16, d_sabs, G, C, c->set_off.p, c->set_idx.p, d_set_ids, n_ids, c->S, d_es);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

int launch_es_wks(fgs_ctx *c, const uint16_t *d_rank16, const double *d_sabs, int G, long long C,
                  double *d_es) {
  if (C <= 0 || c->S <= 0) return FGS_OK;
  if (c->opt_force_generic)
    return launch_es_wks_generic(c, d_rank16, d_sabs, G, C, nullptr, c->S, d_es);
  const size_t Gp = pitch_of(G);
  // words per lane: 4 * odd, so the lanes' uint4 loads of their own range are conflict free
  int wpl = (int)((G + 1023) / 1024);
  wpl = (wpl + 3) & ~3;
  if (((wpl >> 2) & 1) == 0) wpl += 4;
  const size_t per_warp = (size_t)32 * wpl * 6;  // bitmap (4 B/word) + per-word prefix (2 B/word)
  const size_t budget = (size_t)c->smem_optin;
  if (Gp * 2 + per_warp > budget)  // not even one warp fits: the order-free kernel does everything
    return launch_es_wks_generic(c, d_rank16, d_sabs, G, C, nullptr, c->S, d_es);
  int grid = c->num_sms;
  if (grid > C) grid = (int)C;
  {  // phase A
    int nwarps = (int)((budget - Gp * 2) / per_warp);
    if (nwarps > 32) nwarps = 32;
    const size_t smem = Gp * 2 + (size_t)nwarps * per_warp;
    FGS_TRY(c->es_lists.ensure((size_t)c->list_pitch * C + 1024));  // + slack: vector loads past the last list
    FGS_CUDA(cudaFuncSetAttribute(es_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    es_order_kernel<<<grid, nwarps * 32, smem, c->stream>>>(d_rank16, G, C, c->set_info.p, c->set_idx.p,
                                                            c->S, wpl, c->es_lists.p, c->list_pitch);
    c->launches++;
    FGS_CUDA(cudaGetLastError());
  }
  {  // phase B
    const bool w_smem = Gp * 8 <= budget;
    const size_t smem = w_smem ? Gp * 8 : 0;
    if (w_smem) {
      auto k = es_scan_kernel<true>;
      FGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, ES_SCAN_THREADS, smem, c->stream>>>(d_sabs, G, C, c->set_info.p, c->S, c->es_lists.p, c->list_pitch, d_es);
    } else {
      es_scan_kernel<false><<<grid * 2, ES_SCAN_THREADS, 0, c->stream>>>(d_sabs, G, C, c->set_info.p, c->S,
                                                              c->es_lists.p, c->list_pitch, d_es);
    }
    c->launches++;
    FGS_CUDA(cudaGetLastError());
  }
  if (c->n_slow > 0)
    FGS_TRY(launch_es_wks_generic(c, d_rank16, d_sabs, G, C, c->slow_ids.p, c->n_slow, d_es));
  return FGS_OK;
}

int launch_es_mean(fgs_ctx *c, const double *d_scores, int G, long long C, int es_kind,
                   int abs_flag, double *d_es) {
  if (C <= 0 || c->S <= 0) return FGS_OK;
  long long units = (long long)c->S * C;
  long long blocks = (units + 7) / 8;
  long long cap = (long long)c->num_sms * 32;
  if (blocks > cap) blocks = cap;
  es_mean_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(d_scores, G, C, c->set_off.p,
                                                          c->set_idx.p, c->S, es_kind, abs_flag,
                                                          d_es);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// ---------------------------------------------------------------------------
// Observed pass with every extra of flexgsea_weighted_ks_ (:841-908): ES,
// max_es_at, le_prop, running_es_pos/neg, running_es_at and the leading-edge
// span.  S x R units only, one warp each.  The running sums are accumulated in
// double-double so each prefix rounds like R's long-double cumsum, and the
// divisions follow the reference's order (w / sum(w), (r - k) / (G - m)), so
// which.max / which.min land on the same index.
// ---------------------------------------------------------------------------
struct dd { double hi, lo; };
__device__ __forceinline__ dd dd_add_d(dd a, double b) {
  double s = __dadd_rn(a.hi, b);
  double bb = __dsub_rn(s, a.hi);
  double e = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(b, bb));
  e = __dadd_rn(e, a.lo);
  double hi = __dadd_rn(s, e);
  double lo = __dsub_rn(e, __dsub_rn(hi, s));
  return {hi, lo};
}

__global__ void __launch_bounds__(128)
observed_wks_kernel(const uint16_t *__restrict__ rank16, const double *__restrict__ sabs, int G,
                    int R, const int *__restrict__ set_off, const int *__restrict__ set_idx, int S,
                    long long nnz, double *__restrict__ es_out, double *__restrict__ max_es_at,
                    double *__restrict__ le_prop, int *__restrict__ at, double *__restrict__ rpos,
                    double *__restrict__ rneg, int *__restrict__ le_first,
                    int *__restrict__ le_count) {
  const int lane = threadIdx.x & 31;
  const long long unit = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (unit >= (long long)S * R) return;
  const int col = (int)(unit / S), s = (int)(unit % S);
  const size_t Gp = pitch_of(G);
  const uint16_t *rk = rank16 + (size_t)col * Gp;
  const double *sa = sabs + (size_t)col * Gp;
  const int off = set_off[s], m = set_off[s + 1] - off;
  int *at_c = at + (size_t)col * nnz + off;
  double *pos_c = rpos + (size_t)col * nnz + off, *neg_c = rneg + (size_t)col * nnz + off;
  // sum(w) in member order, double-double (R: long double rsum)
  dd tot = {0.0, 0.0};
  if (lane == 0)
    for (int j = 0; j < m; j++) tot = dd_add_d(tot, sa[rk[set_idx[off + j]] - 1]);
  const double W = __shfl_sync(0xffffffffu, tot.hi, 0);
  const double denom = (double)(G - m);
  for (int j = lane; j < m; j += 32) {
    const int gj = set_idx[off + j];
    const int rj = rk[gj];
    int kord = 0;
    for (int t = 0; t < m; t++) {
      const int rt = rk[set_idx[off + t]];
      kord += (rt < rj || (rt == rj && t <= j));
    }
    at_c[kord - 1] = gj + 1;  // running_es_at = gene.set[o], 1-based (:883-885)
  }
  __syncwarp();
  // cumsum over the ordered hits; a single lane keeps R's left-to-right order
  if (lane == 0) {
    dd cum = {0.0, 0.0};
    for (int k = 0; k < m; k++) {
      const int r = rk[at_c[k] - 1];
      const double wo = sa[r - 1] / W;                        // :857
      cum = dd_add_d(cum, wo);                                // :861
      const double d_miss = (double)(r - (k + 1)) / denom;    // :862-863
      const double ps = cum.hi - d_miss;                      // :865
      pos_c[k] = ps;
      neg_c[k] = ps - wo;                                     // :866
    }
  }
  __syncwarp();
  // max / min with first-occurrence index (which.max / which.min, :887, :892)
  double mx = -INFINITY, mn = INFINITY;
  int imx = 0x7fffffff, imn = 0x7fffffff;
  bool has_nan = false;
  for (int k = lane; k < m; k += 32) {
    double ps = pos_c[k], ng = neg_c[k];
    if (isnan(ps) || isnan(ng)) has_nan = true;
    if (ps > mx) { mx = ps; imx = k; }
    if (ng < mn) { mn = ng; imn = k; }
  }
  for (int o = 16; o; o >>= 1) {
    double omx = __shfl_xor_sync(0xffffffffu, mx, o);
    int oimx = __shfl_xor_sync(0xffffffffu, imx, o);
    if (omx > mx || (omx == mx && oimx < imx)) { mx = omx; imx = oimx; }
    double omn = __shfl_xor_sync(0xffffffffu, mn, o);
    int oimn = __shfl_xor_sync(0xffffffffu, imn, o);
    if (omn < mn || (omn == mn && oimn < imn)) { mn = omn; imn = oimn; }
  }
  has_nan = __any_sync(0xffffffffu, has_nan);
  if (lane == 0) {
    const size_t u = (size_t)col * S + s;
    bool do_pos = mx > -mn;  // :870
    double es = do_pos ? mx : mn;
    int wmes = do_pos ? imx : imn;
    if (has_nan || m == 0) { es = nan(""); do_pos = false; wmes = 0; }
    if (wmes == 0x7fffffff) wmes = 0;
    es_out[u] = es;
    max_es_at[u] = m > 0 ? (double)rk[at_c[wmes] - 1] : nan("");   // :897-899
    const int first = do_pos ? 0 : wmes, count = do_pos ? wmes + 1 : m - wmes;  // :889, :894
    le_first[u] = first;
    le_count[u] = count;
    le_prop[u] = (double)count / (double)m;                        // :900-902
  }
}

int launch_observed_wks(fgs_ctx *c, const uint16_t *d_rank16, const double *d_sabs, int G, int R,
                        double *d_es, double *d_max_es_at, double *d_le_prop, int *d_at,
                        double *d_pos, double *d_neg, int *d_le_first, int *d_le_count) {
  long long units = (long long)c->S * R;
  if (units <= 0) return FGS_OK;
  observed_wks_kernel<<<(unsigned)((units + 3) / 4), 128, 0, c->stream>>>(
      d_rank16, d_sabs, G, R, c->set_off.p, c->set_idx.p, c->S, c->nnz, d_es, d_max_es_at,
      d_le_prop, d_at, d_pos, d_neg, d_le_first, d_le_count);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

}  // namespace fgs
