// K4 / K5: enrichment scores per (gene set, score column).
//
// K4 replaces the fast branch of flexgsea_weighted_ks_ (R/flexgsea.R:816-840):
//   w = |s[set]|^p / sum(.);  o = order(rank[set]);  d_hit = cumsum(w[o]);
//   d_miss = (rank[set][o] - 1..m) / (G - m);  pos = d_hit - d_miss;
//   neg = pos - w[o];  es = max(pos) > -min(neg) ? max(pos) : min(neg)
// K5 replaces flexgsea_maxmean_ (:725-741) and flexgsea_mean_ (:755-767).
//
// es_wks_kernel: one persistent CTA per SM, a column at a time, EIGHT LANES per
// gene set (four sets per warp).  The column's rank array (uint16, gene -> rank)
// and its rank-ordered hit weights (fp64, rank -> |score|^p) are staged in
// shared memory (10 G bytes).  The lanes of a set gather its ranks, sort them IN
// REGISTERS with a bitonic network on packed uint16 pairs (no shared-memory
// scratch), and run the running sum as lane-local sums + one scan of the lane
// totals.  Sets are handled in descending size in compile-time size classes
// (8 lanes x {4, 8, 16, 32, 64} ranks, a whole warp x {32, 64}).
//
// Sets with duplicated members (R keeps them; a pad-free distinct-rank sort
// cannot) or more than ES_MAX_SET members go through es_wks_generic_kernel, an
// O(m^2 / 32) warp kernel that needs no ordering at all: max/min over the
// members of pos/neg is order independent once each member knows its ordinal
// and prefix weight, which it gets by comparing against every other member.
#include <algorithm>

#include "fgs_common.cuh"

namespace fgs {

// ES_MAX_SET (fgs_common.cuh): largest set the in-register kernel takes

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

// ---- in-register bitonic sort of a set's ranks --------------------------------
// N = LANES * NB uint16 keys held by LANES consecutive lanes, NB per lane, two
// per 32-bit register (VIMNMX.U16x2 compares both at once).  Blocked layout:
// element e = lane * NB + slot, so after the sort lane l holds the hits with
// ordinals [l*NB, l*NB + NB) -- exactly the distribution the running-sum scan
// wants.  Pads carry the key G.
//
// LANES = 8 puts four sets in one warp: a comparator stage between lanes costs
// five instructions per register (SHFL, two predicated VIMNMX, two moves) but
// one between registers only one, so the fewer lanes a set spans the cheaper its
// sort -- 6 cross-lane stages instead of 15 -- and every fixed cost (scan,
// reduction, set bookkeeping) is shared by four sets.
//
// The network is the direction-free form of bitonic sort: level k first
// compares e with its mirror e ^ (k - 1), then e with e ^ j for j = k/4 .. 1,
// and EVERY comparator sends the minimum to the lower index.  So all in-lane
// comparators have a compile-time direction (one VIMNMX each), and only the
// cross-lane ones depend on the lane's side of the exchange.
__device__ __forceinline__ uint32_t swap_halves(uint32_t v) { return __byte_perm(v, v, 0x1032); }
// comparator between the two halves of one register: as 32-bit numbers, the
// larger of (hi:lo) and (lo:hi) is the one with the larger key on top and the
// smaller below -- one PRMT and one 32-bit max
__device__ __forceinline__ uint32_t sort_halves(uint32_t v) { return max(v, swap_halves(v)); }
// cross-lane comparator: `hi_mask` is 0 on the lane that keeps the minimum and
// ~0 on the lane that keeps the maximum
__device__ __forceinline__ uint32_t minmax2_side(uint32_t a, uint32_t o, uint32_t hi_mask) {
#ifdef FGS_ES_XORCMP
  // experiment r2-E1 (profiles/README.md): min, then max = a ^ o ^ min on the lanes that keep the
  // maximum -- VIMNMX + two LOP3 instead of two predicated VIMNMX + moves
  const uint32_t t = __vminu2(a, o);
  return t ^ ((a ^ o) & hi_mask);
#else
  return hi_mask ? __vmaxu2(a, o) : __vminu2(a, o);
#endif
}

//
// A set need not fill its eight lanes: one of L < 8 lanes (m <= L * NB) is sorted by the same
// 8-lane network with lanes L..7 imagined full of pads.  Every comparator sends the minimum
// down, so one whose upper end is an imagined lane changes nothing; the lane on its lower end
// reads an all-pad register from lane 31 instead (kept idle, and so all pads, whenever L < 8).
// `base` = first lane of the set, `Lp` = L on the set's lanes and 0 on idle ones, which then
// read lane 31 throughout and stay all pads.  Sets of L lanes sit back to back in the warp
// (floor(31 / L) of them), which is where the gain is: 6 sets of 5 lanes instead of 4 of 8.
// source lane of a cross-lane comparator (see above)
template <int LANES>
__device__ __forceinline__ int bitonic_src(int lane, int mask, int base, int Lp) {
  if constexpr (LANES == 32) return lane ^ mask;
  const int pl = lane ^ mask;
  return pl < Lp ? base + pl : 31;
}

// one e <-> e ^ J stage; the stages of a level follow each other by template recursion
// (J, J/2, .. 1) so that the whole network is straight-line code whatever its size -- a
// `#pragma unroll` over the levels gives up at 2 048 keys and leaves the keys in local memory
template <int LANES, int NB, int J>
__device__ __forceinline__ void bitonic_xor_stages(uint32_t (&v)[NB / 2], int lane, int base, int Lp) {
  constexpr int NR = NB / 2;
  if constexpr (J > 0) {
    if constexpr (J >= NB) {  // partner lane ^ (J / NB), same slot
      const uint32_t hi_mask = 0u - ((lane / (J / NB)) & 1u);
      const int src = bitonic_src<LANES>(lane, J / NB, base, Lp);
#pragma unroll
      for (int q = 0; q < NR; q++)
        v[q] = minmax2_side(v[q], __shfl_sync(0xffffffffu, v[q], src), hi_mask);
    } else if constexpr (J == NR) {  // the two halves of one register
#pragma unroll
      for (int q = 0; q < NR; q++) v[q] = sort_halves(v[q]);
    } else {  // partner register q ^ J, same half
#pragma unroll
      for (int q = 0; q < NR; q++) {
        const int qb = q ^ J;
        if (qb > q) {
          const uint32_t a = v[q], b2 = v[qb];
          v[q] = __vminu2(a, b2);
          v[qb] = __vmaxu2(a, b2);
        }
      }
    }
    bitonic_xor_stages<LANES, NB, J / 2>(v, lane, base, Lp);
  }
}

// level K: the mirror stage e <-> e ^ (K - 1), the stages K/4 .. 1, then level 2K
template <int LANES, int NB, int K>
__device__ __forceinline__ void bitonic_levels(uint32_t (&v)[NB / 2], int lane, int base, int Lp) {
  constexpr int N = LANES * NB;
  constexpr int NR = NB / 2;
  if constexpr (K <= N) {
    if constexpr (K <= NR) {  // only register bits flip: same half
#pragma unroll
      for (int q = 0; q < NR; q++) {
        const int qb = q ^ (K - 1);
        if (qb > q) {
          const uint32_t a = v[q], t = v[qb];
          v[q] = __vminu2(a, t);
          v[qb] = __vmaxu2(a, t);
        }
      }
    } else if constexpr (K == NB) {  // every slot bit flips: (half 0, q) <-> (half 1, NR-1-q) and (1, q) <-> (0, NR-1-q)
#pragma unroll
      for (int q = 0; q < NR / 2; q++) {
        const int qb = NR - 1 - q;
        const uint32_t a = v[q], bs = swap_halves(v[qb]);
        const uint32_t mn = __vminu2(a, bs), mx = __vmaxu2(a, bs);
        v[q] = __byte_perm(mn, mx, 0x7610);   // low: min(a.lo, b.hi) stays below; high: max(a.hi, b.lo)
        v[qb] = __byte_perm(mn, mx, 0x5432);  // low: min(a.hi, b.lo); high: max(a.lo, b.hi)
      }
    } else {  // lane ^ (K / NB - 1), slots reversed: register NR-1-q, halves swapped
      const uint32_t hi_mask = 0u - ((lane / ((K / NB) >> 1)) & 1u);
      const int src = bitonic_src<LANES>(lane, (K / NB) - 1, base, Lp);
#pragma unroll
      for (int q = 0; q < NR / 2; q++) {  // registers q and NR-1-q exchange with each other's mirror
        const uint32_t a = v[q], b = v[NR - 1 - q];
        const uint32_t oa = swap_halves(__shfl_sync(0xffffffffu, b, src));
        const uint32_t ob = swap_halves(__shfl_sync(0xffffffffu, a, src));
        v[q] = minmax2_side(a, oa, hi_mask);
        v[NR - 1 - q] = minmax2_side(b, ob, hi_mask);
      }
    }
    bitonic_xor_stages<LANES, NB, K / 4>(v, lane, base, Lp);
    bitonic_levels<LANES, NB, K * 2>(v, lane, base, Lp);
  }
}

// Slot s of a lane lives in register s % NR, half s / NR: the TOP slot bit selects the
// half.  A comparator stage on the half bit costs two instructions per register (PRMT +
// 32-bit max), one between registers a single VIMNMX; the top slot bit is the in-lane
// bit the network touches least often (log2(LANES) + 1 times, the lowest bit log2(N)).
template <int LANES, int NB>
__device__ __forceinline__ void bitonic_sort_ranks(uint32_t (&v)[NB / 2], int lane, int base = 0, int Lp = LANES) {
  static_assert(NB >= 4 && LANES <= 32, "packed pairs, at least two registers");
  bitonic_levels<LANES, NB, 2>(v, lane, base, Lp);
}

// Exact warp maximum of a double through two integer warp reductions (REDUX) on
// the order-preserving key (sign-flip transform): ~12 instructions instead of a
// five-step shuffle butterfly of 64-bit compares (~45).  Returns one of the
// inputs bit for bit, so the strict tie rule of the reference is unaffected.
__device__ __forceinline__ double warp_max_exact(double x) {
  const int hi = __double2hiint(x);
  const uint32_t lo = (uint32_t)__double2loint(x);
  const uint32_t m = (uint32_t)(hi >> 31);                 // all ones for negatives
  const uint32_t khi = (uint32_t)hi ^ (m | 0x80000000u);   // unsigned order == double order
  const uint32_t klo = lo ^ m;
  const uint32_t mh = __reduce_max_sync(0xffffffffu, khi);
  const uint32_t ml = __reduce_max_sync(0xffffffffu, khi == mh ? klo : 0u);
  const uint32_t inv = (mh & 0x80000000u) ? 0u : 0xffffffffu;
  return __hiloint2double((int)(mh ^ (inv | 0x80000000u)), (int)(ml ^ inv));
}

constexpr double ES_OBS_REL = 1e-9;  // |null - observed| below this fraction of the observed ES: redo exactly
// a null ES within rounding of the set's observed ES (label-preserving permutations,
// small universes): whether the two compare equal is the reference's rounding, so the
// unit is redone exactly like a tied one
__device__ __forceinline__ bool near_observed(const double *__restrict__ obs, const int *__restrict__ ident,
                                              int S, int R, long long col, int set_id, double es) {
  if (!obs) return false;
  if (ident && __ldg(ident + col / R)) return false;  // label-preserving permutation: the column is copied
  const double e = __ldg(obs + (size_t)(col % R) * S + set_id);
  return fabs(es - e) <= ES_OBS_REL * fabs(e);
}
constexpr double ES_TIE_REL = 1e-9;  // |es_pos + es_neg| below this fraction of their size: redo exactly
__device__ __forceinline__ void tie_push(int2 *__restrict__ list, int *__restrict__ ctr, int cap, int col,
                                         int set_id) {
  const int at = atomicAdd(ctr, 1);  // a count above cap makes es_wks_tie_kernel redo every unit
  if (at < cap) list[at] = make_int2(col, set_id);
}

// Two members' ranks (0-based, from the staged column) packed in one register;
// their hit weights (rank-indexed |score|^p; the pad rank G has weight 0) are
// added to the lane's share of the set's weight total.
template <bool W_SMEM>
__device__ __forceinline__ uint32_t gather2(const uint16_t *__restrict__ s_rank,
                                            const double *__restrict__ wsrc, uint32_t G, uint32_t pair,
                                            double &w0, double &w1) {
  const uint32_t lo = s_rank[pair & 0xffffu], hi = s_rank[pair >> 16];
  w0 += W_SMEM ? wsrc[lo] : (lo < G ? wsrc[lo] : 0.0);
  w1 += W_SMEM ? wsrc[hi] : (hi < G ? wsrc[hi] : 0.0);
  return __byte_perm(lo, hi, 0x5410);
}

// One set on LANES consecutive lanes (`ls` = lane within them), SL members per
// lane.  Gather the members' ranks (and, order-free, the set's total weight W),
// sort the ranks in registers, then ONE pass over the ordered hits per lane with
// the lane-local running sum: the weight of the lanes below is a per-lane
// constant, so it is added to the lane's maximum / minimum afterwards (one scan
// of the lane totals).  Returns the ES (valid in every lane of the set).
//   idx16: the set's members as uint16 gene ids in 16-byte blocks of eight,
//   padded with the pad gene G, whose staged rank is the pad key G.
template <int LANES, int SL, bool W_SMEM>
__device__ __forceinline__ double wks_sets(const uint16_t *__restrict__ s_rank,
                                           const double *__restrict__ wsrc,
                                           const uint16_t *__restrict__ idx16, int zblk, int m, int G,
                                           double invGm, int ls, bool &tie, int base = 0, int L = LANES,
                                           int Lp = LANES) {
  constexpr int NR = SL / 2;
  const uint32_t padpair = (uint32_t)G * 0x10001u;
  uint32_t v[NR];
  double w0 = 0.0, w1 = 0.0;
  if constexpr (SL == 4) {  // four members per lane: one 8-byte load
    uint2 w = make_uint2(padpair, padpair);
    if (ls * 4 < m) w = __ldg((const uint2 *)(idx16 + (size_t)zblk * 8) + ls);
    v[0] = gather2<W_SMEM>(s_rank, wsrc, G, w.x, w0, w1);
    v[1] = gather2<W_SMEM>(s_rank, wsrc, G, w.y, w0, w1);
  } else {
#pragma unroll
    for (int b = 0; b < SL / 8; b++) {  // eight members per 16-byte load
      const int blk = b * L + ls;
      uint4 w = make_uint4(padpair, padpair, padpair, padpair);
      if (blk * 8 < m) w = __ldg((const uint4 *)(idx16 + (size_t)zblk * 8) + blk);
      v[4 * b + 0] = gather2<W_SMEM>(s_rank, wsrc, G, w.x, w0, w1);
      v[4 * b + 1] = gather2<W_SMEM>(s_rank, wsrc, G, w.y, w0, w1);
      v[4 * b + 2] = gather2<W_SMEM>(s_rank, wsrc, G, w.z, w0, w1);
      v[4 * b + 3] = gather2<W_SMEM>(s_rank, wsrc, G, w.w, w0, w1);
    }
  }
  double W = w0 + w1;  // every lane of the set ends with the same bits
  if constexpr (LANES == 32) {
#pragma unroll
    for (int o = LANES / 2; o; o >>= 1) W += __shfl_xor_sync(0xffffffffu, W, o);
  } else {  // L <= 8 lanes from `base`: fold towards the first lane, then hand its total out
#pragma unroll
    for (int d = 1; d < LANES; d <<= 1) {
      const double t = __shfl_down_sync(0xffffffffu, W, d);
      if (ls + d < L) W += t;
    }
    W = __shfl_sync(0xffffffffu, W, base);
  }
  const double invW = 1.0 / W;

  bitonic_sort_ranks<LANES, SL>(v, ls, base, Lp);

  const int k0 = ls * SL;
  // (r - k) as a double without integer->double conversions: 2^52 + r0 has the
  // integer in its low mantissa word; subtracting 2^52 + k0 and then t is exact
  const double kbase = __hiloint2double(0x43300000, k0);
  const int nvalid = m - k0;  // slots [0, nvalid) of this lane hold members, the rest pads
  double cl = 0.0, mx = -INFINITY, mn = INFINITY;
#pragma unroll
  for (int t = 0; t < SL; t++) {
    const uint32_t r0 = (v[t % NR] >> (16 * (t / NR))) & 0xffffu;  // slot t: register t % NR, half t / NR
    const double w = (W_SMEM || t < nvalid) ? wsrc[r0] : 0.0;
    cl += w;
    // d_miss = (r - k) / (G - m) with r = r0 + 1, k = k0 + t + 1 (R/flexgsea.R:825-826)
    const double rmk = (__hiloint2double(0x43300000, (int)r0) - kbase) - (double)t;
    const double ps = cl * invW - rmk * invGm;
    const double ng = ps - w * invW;
    // No branch around the pads (key G, weight 0, ordinals >= m): there
    // pos = neg = (k - m) / (G - m) >= 0 up to rounding, which can never lower the
    // minimum (neg of the first hit is <= 0), so only the maximum needs the guard.
    mx = (t < nvalid && ps > mx) ? ps : mx;
    mn = (ng < mn) ? ng : mn;
  }
  // weight of the lanes below: inclusive scan of the lane totals over the set's
  // lanes; the shuffle's own in-range predicate guards the add (no select)
  double incw = cl;
  if constexpr (LANES == 32) {
#pragma unroll
    for (int d = 1; d < LANES; d <<= 1) {
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 lo, hi, tlo, thi;\n\t.reg .f64 t;\n\t"
                   "mov.b64 {lo, hi}, %0;\n\t"
                   "shfl.sync.up.b32 tlo|p, lo, %1, %2, 0xffffffff;\n\t"
                   "shfl.sync.up.b32 thi|p, hi, %1, %2, 0xffffffff;\n\t"
                   "mov.b64 t, {tlo, thi};\n\t"
                   "@p add.f64 %0, %0, t;\n\t}"
                   : "+d"(incw) : "r"(d), "n"((32 - LANES) << 8));
    }
  } else {  // the set's lanes start at any lane: the lane's own position guards the add
#pragma unroll
    for (int d = 1; d < LANES; d <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, incw, d);
      if (ls >= d) incw += t;
    }
  }
  const double below = (incw - cl) * invW;  // one more rounding than a single running sum; ES tolerance is 1e-10
  mx += below;
  mn += below;
  if constexpr (LANES == 32) {
    mx = warp_max_exact(mx);
    mn = -warp_max_exact(-mn);
  } else {  // folded towards the set's first lane, which is the one that reports
#pragma unroll
    for (int d = 1; d < LANES; d <<= 1) {
      const double a = __shfl_down_sync(0xffffffffu, mx, d), b = __shfl_down_sync(0xffffffffu, mn, d);
      const bool in = ls + d < L;
      mx = (in && a > mx) ? a : mx;
      mn = (in && b < mn) ? b : mn;
    }
  }
  double es = (mx > -mn) ? mx : mn;                      // strict, :833
  // es_pos and -es_neg equal up to rounding (a structural tie: e.g. the unweighted
  // statistic, p = 0): which side wins is decided by the reference's own rounding,
  // so the unit is redone in the reference's operation order (es_wks_tie_kernel)
  tie = fabs(mx + mn) <= ES_TIE_REL * fmax(fabs(mx), fabs(mn));
  if (W == 0.0 || m == G || m == 0) { es = nan(""); tie = false; }  // 0/0 in the reference
  return es;
}

constexpr int ES_THREADS = 1024;  // 32 warps at 64 registers: better than 24 warps at 80 (measured)
constexpr int ES_LONG_THREADS = 768;  // columns too long for on-chip weights: gathers through L1 / L2, 80 registers
constexpr int ES_WIDE_THREADS = 512;  // sets on a whole warp keep up to 32 packed registers: 128 registers per thread
// WIDE = false: the groups of sets up to 512 members (eight-lane network, L lanes per set);
// WIDE = true: the groups that are one set of 513 .. ES_MAX_SET members on a whole warp.  Two
// instantiations because the whole-warp classes need twice the registers; the second one is
// launched only when such sets exist.
template <bool W_SMEM, bool WIDE>
__global__ void __launch_bounds__(WIDE ? ES_WIDE_THREADS : (W_SMEM ? ES_THREADS : ES_LONG_THREADS), 1)
es_wks_kernel(const uint16_t *__restrict__ rank16, const double *__restrict__ sabs, int G,
              long long C, const int4 *__restrict__ set_info, const uint16_t *__restrict__ idx16,
              const double *__restrict__ set_invgm, int S, double *__restrict__ es_out,
              int2 *__restrict__ tie_list, int *__restrict__ tie_ctr, int tie_cap,
              const double *__restrict__ obs, const int *__restrict__ ident, int R,
              const int2 *__restrict__ groups, int n_groups, int tail_split, int bulk) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const size_t Gp = pitch_of(G);
  uint16_t *s_rank = (uint16_t *)smem;                    // [Gp + 8]: 0-based ranks, pad gene at [G]
  double *s_abs = (double *)(smem + (Gp + 8) * 2);
  __shared__ int s_next;  // next group of the column (see below)
  __shared__ __align__(8) uint64_t s_bar;  // counts the bytes of the weight column in flight (bulk copy)
  if (W_SMEM && tid < 8) s_abs[G + tid] = 0.0;  // zero slot(s) behind the weights: the pads' weight
  __shared__ uint32_t s_phase;             // its phase: flips once per column of this CTA
  if (W_SMEM && bulk && tid == 0) { bulk_bar_init(&s_bar); s_phase = 0; }

  // Whole rounds of one column per CTA, then the tail round: the C % gridDim.x columns left over
  // are split `tail_split` ways (every CTA stages the column, the groups go round by `part`), so
  // that the last round keeps the SMs busy for 1 / tail_split of a column instead of idling most
  // of them for a whole one -- what a rank's share of a sharded null (1 250 columns = 8.45 rounds
  // at C2 over eight GPUs) would otherwise pay.
  // (a launch holds one block of columns: int arithmetic; part / nparts live in shared memory so
  // that the group loop keeps its registers)
  const int C_full = ((int)C / (int)gridDim.x) * (int)gridDim.x;
  __shared__ int s_part[2];
  for (int w = blockIdx.x;; w += gridDim.x) {
    int col = w;
    if (w >= C_full) {
      const int t = w - C_full;
      if (t >= ((int)C - C_full) * tail_split) break;
      col = C_full + t / tail_split;
    }
    __syncthreads();
    {  // stage the column.  The weights are a plain copy: the copy engine takes them (one bulk copy of
      // the G & ~1 doubles, rows are 64-byte aligned) while the threads turn the ranks 0-based
      // with 16-byte vector copies; without `bulk` the threads copy the weights too.
      if (W_SMEM && bulk && tid == 0) bulk_g2s(s_abs, sabs + (size_t)col * Gp, (uint32_t)(G & ~1) * 8u, &s_bar);
      const uint4 *src = (const uint4 *)(rank16 + (size_t)col * Gp);
      uint4 *dst = (uint4 *)s_rank;
      for (int i = tid; i < (int)(Gp / 8); i += blockDim.x) {
        uint4 r = __ldg(src + i);
        r.x = __vsub2(r.x, 0x00010001u); r.y = __vsub2(r.y, 0x00010001u);
        r.z = __vsub2(r.z, 0x00010001u); r.w = __vsub2(r.w, 0x00010001u);
        dst[i] = r;
      }
      if (W_SMEM) {
        if (bulk) {
          if ((G & 1) && tid == 32) s_abs[G - 1] = __ldg(sabs + (size_t)col * Gp + G - 1);
          bulk_wait(&s_bar, s_phase);  // (s_phase is flipped between the two barriers below)
        } else {
          for (int i = tid; i < G; i += blockDim.x) s_abs[i] = __ldg(sabs + (size_t)col * Gp + i);
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      s_rank[G] = (uint16_t)G; s_next = 0;  // the pad gene's key (inside the staged row when G < Gp)
      if (W_SMEM && bulk) s_phase ^= 1u;
      const bool in_tail = w >= C_full;
      s_part[0] = in_tail ? (w - C_full) % tail_split : 0;
      s_part[1] = in_tail ? tail_split : 1;
    }
    __syncthreads();
    const double *wsrc = W_SMEM ? s_abs : sabs + (size_t)col * Gp;
    // observed ES of this column's response; none for a label-preserving permutation (copied later)
    const double *col_obs = (obs && !(ident && __ldg(ident + col / R))) ? obs + (size_t)(col % R) * S : nullptr;
    double *es_col = es_out + (size_t)col * S;
    // sets in descending size; a group (fgs_set_gene_sets) is the sets one warp takes at once:
    // `cnt` sets of one size class on L lanes each, back to back.  A warp fetches its next group
    // from a counter in shared memory: largest first, whoever is free -- the groups of a column
    // differ 16-fold in cost, and a fixed round-robin leaves the first warps with more work than
    // the last in every round.
    for (;;) {
      int g = 0;
      if (lane == 0) g = atomicAdd(&s_next, 1);
      g = s_part[0] + s_part[1] * __shfl_sync(0xffffffffu, g, 0);
      if (g >= n_groups) break;
      const int2 ge = __ldg(groups + g);  // {first set (processing order), cnt | L << 8 | log2(SL) << 16}
      const int cnt = ge.y & 0xff, L = (ge.y >> 8) & 0xff, sl = ge.y >> 16;
      if constexpr (!WIDE) {
        const int seg = lane / L, ls = lane - seg * L;
        const bool on = seg < cnt;
        const int k = ge.x + seg, base = lane - ls, Lp = on ? L : 0;
        int4 info = make_int4(0, (int)0x80000000, 0, 0);  // {-, m | slow << 31, member block offset, set id}
        double ig = 0.0;
        if (on) { info = __ldg(set_info + k); ig = __ldg(set_invgm + k); }  // ig = 1 / (G - m)
        // the set's observed ES, fetched now so that the comparison at the end costs nothing
        const double e_obs = (col_obs && on) ? __ldg(col_obs + info.w) : nan("");
        const bool mine = info.y >= 0;              // slow sets: es_wks_generic_kernel does these
        const int m = mine ? info.y : 0;
        double es;
        bool tie = false;
        if (sl == 2) es = wks_sets<8, 4, W_SMEM>(s_rank, wsrc, idx16, info.z, m, G, ig, ls, tie, base, L, Lp);
        else if (sl == 3) es = wks_sets<8, 8, W_SMEM>(s_rank, wsrc, idx16, info.z, m, G, ig, ls, tie, base, L, Lp);
        else if (sl == 4) es = wks_sets<8, 16, W_SMEM>(s_rank, wsrc, idx16, info.z, m, G, ig, ls, tie, base, L, Lp);
        else if (sl == 5) es = wks_sets<8, 32, W_SMEM>(s_rank, wsrc, idx16, info.z, m, G, ig, ls, tie, base, L, Lp);
        else es = wks_sets<8, 64, W_SMEM>(s_rank, wsrc, idx16, info.z, m, G, ig, ls, tie, base, L, Lp);  // 257 .. 512 members
        if (mine && ls == 0) {
          es_col[info.w] = es;
          if (tie || fabs(es - e_obs) <= ES_OBS_REL * fabs(e_obs)) tie_push(tie_list, tie_ctr, tie_cap, (int)col, info.w);
        }
      } else {  // a set of 513 .. ES_MAX_SET members on the whole warp
        (void)cnt; (void)L;
        const int4 info = __ldg(set_info + ge.x);
        if (info.y < 0) continue;
        const double ig = __ldg(set_invgm + ge.x);
        const double e_obs = col_obs ? __ldg(col_obs + info.w) : nan("");
        double es;
        bool tie = false;
        if (sl == 5) es = wks_sets<32, 32, W_SMEM>(s_rank, wsrc, idx16, info.z, info.y, G, ig, lane, tie);
        else es = wks_sets<32, 64, W_SMEM>(s_rank, wsrc, idx16, info.z, info.y, G, ig, lane, tie);
        if (lane == 0) {
          es_col[info.w] = es;
          if (tie || fabs(es - e_obs) <= ES_OBS_REL * fabs(e_obs)) tie_push(tie_list, tie_ctr, tie_cap, (int)col, info.w);
        }
      }
    }
    if (w >= C_full) break;  // the tail round is the last
  }
}

// Members of every set as uint16 gene ids, in processing order, each set in
// 16-byte blocks of eight padded with the pad gene G (a warp per set).
//
// The ES kernels read these lists eight lanes at a time with a stride of eight
// members (lane l takes block l of a 64-member chunk), and use the gene id as an
// index into shared memory.  The order of a set's members is free (sums and
// sorts do not care), so every 64-member chunk is stored sorted by gene id mod 16:
// members eight apart then fall into different bank pairs of the fp64 score
// column (mean / maxmean kernel) instead of colliding at random.
__global__ void build_idx16_kernel(const int4 *__restrict__ set_info, const int *__restrict__ set_idx,
                                   int S, int G, uint16_t *__restrict__ idx16) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= S) return;
  const int4 info = set_info[k];
  const int m = info.y & 0x7fffffff, mpad = (m + 7) & ~7;
  uint16_t *out = idx16 + (size_t)info.z * 8;
  for (int c0 = 0; c0 < mpad; c0 += 64) {
    // two members per lane; key = (class, position) is unique, pads sort last
    const int j0 = c0 + lane, j1 = c0 + 32 + lane;
    const int g0 = j0 < m ? set_idx[info.x + j0] : G, g1 = j1 < m ? set_idx[info.x + j1] : G;
#ifdef FGS_ES_RANKBANK
    // experiment r2-E2: order by the bank of the uint16 rank array ((g >> 1) & 31) instead of the
    // bank pair of the fp64 score column (what the maxmean / mean kernel gathers)
    const int key0 = ((j0 < m ? ((g0 >> 1) & 31) : 32) << 8) | lane, key1 = ((j1 < m ? ((g1 >> 1) & 31) : 32) << 8) | (32 + lane);
#else
    const int key0 = ((j0 < m ? (g0 & 15) : 16) << 8) | lane, key1 = ((j1 < m ? (g1 & 15) : 16) << 8) | (32 + lane);
#endif
    int p0 = 0, p1 = 0;
    for (int f = 0; f < 32; f++) {
      const int a = __shfl_sync(0xffffffffu, key0, f), b = __shfl_sync(0xffffffffu, key1, f);
      p0 += (a < key0) + (b < key0);
      p1 += (a < key1) + (b < key1);
    }
    if (c0 + p0 < mpad) out[c0 + p0] = (uint16_t)g0;
    if (c0 + p1 < mpad) out[c0 + p1] = (uint16_t)g1;
  }
}

// The same member blocks, arranged for the bank structure of es_mean_smem_kernel's gathers.
//
// That kernel reads the fp64 score column by gene id, eight lanes per set and two sets per half
// warp: the load of step t = 8 b + j takes, from lane ls of a set, the member at position
// (8 b + ls) * 8 + j.  A half-warp's sixteen 8-byte reads are one shared-memory wavefront only if
// they fall into sixteen different bank pairs, i.e. the sixteen gene ids differ mod 16.  The
// order of a set's members is free (sums do not care) and fixed for all columns, so it is chosen
// here, for the two sets of a pair together: at every step the sixteen residues are dealt out,
// eight to each set -- a residue goes to the set that has more members of it left -- and each
// lane takes an unused member of its residue.  That covers the rounds in which all sixteen lanes
// hold members (64 members of each set per round); what is left over (unequal class sizes, the
// ragged end, pads) fills the remaining positions in any order.  A warp per pair of sets
// (2 u, 2 u + 1 in processing order: the pairs es_mean_smem_kernel puts on one half-warp).
constexpr int ARR_MAX = 4096;  // members of a pair the scratch holds; larger pairs keep build_idx16_kernel's order
__global__ void __launch_bounds__(256)
arrange_idx16_kernel(const int4 *__restrict__ set_info, const int *__restrict__ set_idx, int S, int G,
                     uint16_t *__restrict__ idx16) {
  extern __shared__ uint16_t arr_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * (blockDim.x >> 5) + warp;
  if (2 * u >= S) return;
  uint16_t *cls = arr_smem + (size_t)warp * ARR_MAX;   // members grouped by residue: set A's classes, then set B's
  const int kA = 2 * u, kB = 2 * u + 1;
  const int4 iA = set_info[kA];
  const int4 iB = kB < S ? set_info[kB] : make_int4(0, 0, 0, 0);
  const int mA = iA.y & 0x7fffffff, mB = iB.y & 0x7fffffff;
  // rounds in which every lane of both sets holds members (a lone last set: its own rounds)
  const int rounds = kB < S ? min(mA, mB) / 64 : mA / 64;
  // the scratch holds both sets' members and, per set, the positions it could not fill (at most
  // the dealt region); pairs that do not fit keep build_idx16_kernel's order
  if (mA + mB + 2 * rounds * 64 > ARR_MAX) return;
  // this lane's class: lanes 0..15 = residues of A, 16..31 = residues of B
  const bool isB = lane >= 16;
  const int r = lane & 15;
  const int m_mine = isB ? mB : mA, off_mine = isB ? iB.x : iA.x;
  // class sizes and starts (each lane counts its own class; sets are a few hundred members)
  int cnt = 0;
  for (int j = 0; j < m_mine; j++) cnt += (set_idx[off_mine + j] & 15) == r;
  int start = cnt;  // exclusive prefix over the 16 classes of the lane's set
  for (int d = 1; d < 16; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, start, d, 16);
    if (r >= d) start += t;
  }
  start -= cnt;
  if (isB) start += mA;
  for (int j = 0, at = start; j < m_mine; j++) {
    const int g = set_idx[off_mine + j];
    if ((g & 15) == r) cls[at++] = (uint16_t)g;
  }
  __syncwarp();
  uint16_t *outA = idx16 + (size_t)iA.z * 8, *outB = idx16 + (size_t)iB.z * 8;
  uint16_t *out_mine = isB ? outB : outA;
  int used = 0;                      // members of this lane's class handed out so far
  // holes: positions of the dealt region a set could not fill (its class had run out)
  __shared__ int s_nhole[8][2];
  int *nhole = s_nhole[warp];
  if (lane < 2) nhole[lane] = 0;
  __syncwarp();
  uint16_t *holesA = cls + mA + mB;  // positions (uint16), kept behind the class lists
  const int hole_cap = rounds * 64;
  uint16_t *holesB = holesA + hole_cap;
  for (int t = 0; t < rounds * 8; t++) {
    const int b = t >> 3, j = t & 7;
    const int other = __shfl_xor_sync(0xffffffffu, cnt - used, 16);  // the partner set's supply of this residue
    const int d = isB ? other - (cnt - used) : (cnt - used) - other; // A's surplus of residue r (same on lanes r and r + 16)
    int rank = 0;                                                     // rank of the residue by surplus, ties by index
    for (int o = 0; o < 16; o++) {
      const int dq = __shfl_sync(0xffffffffu, d, o);
      rank += (dq > d) || (dq == d && o < r);
    }
    // A takes the eight residues it has most surplus of (its lane `rank`), B the others (its lane `rank - 8`)
    const bool mine = isB ? rank >= 8 : rank < 8;
    if (mine) {
      const int ls = isB ? rank - 8 : rank;
      const int pos = (b * 8 + ls) * 8 + j;
      if (used < cnt) out_mine[pos] = cls[start + used++];
      else if (m_mine > 0) (isB ? holesB : holesA)[atomicAdd(&nhole[isB], 1)] = (uint16_t)pos;
    }
  }
  __syncwarp();
  // the rest, in class order: first the holes of the dealt region, then the positions behind it; pads at the end
  // exclusive prefix of the leftovers over the classes of each set
  const int left = cnt - used;
  int lstart = left;
  for (int dd = 1; dd < 16; dd <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, lstart, dd, 16);
    if (r >= dd) lstart += t;
  }
  lstart -= left;
  const int nh = nhole[isB];
  const uint16_t *holes_mine = isB ? holesB : holesA;
  for (int i = 0; i < left; i++) {
    const int q = lstart + i;          // index among the set's leftovers
    const int pos = q < nh ? holes_mine[q] : rounds * 64 + (q - nh);
    out_mine[pos] = cls[start + used + i];
  }
  __syncwarp();
  // pads behind the last member
  const int mpad = (m_mine + 7) & ~7;
  if (r < mpad - m_mine) out_mine[m_mine + r] = (uint16_t)G;
}

// Order-free weighted KS for sets the bitonic kernel cannot take (duplicates,
// > ES_MAX_SET members).  One warp per (set, column).  Ties in rank (duplicate
// members) are ordered by member position, like R's stable order() (:822).
__global__ void __launch_bounds__(256)
es_wks_generic_kernel(const uint16_t *__restrict__ rank16, const double *__restrict__ sabs, int G,
                      long long C, const int *__restrict__ set_off, const int *__restrict__ set_idx,
                      const int *__restrict__ set_ids, int n_ids, int S, double *__restrict__ es_out,
                      int2 *__restrict__ tie_list, int *__restrict__ tie_ctr, int tie_cap,
                      const double *__restrict__ obs, const int *__restrict__ ident, int R) {
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
  bool has_nan = false;
  for (int j = lane; j < m; j += 32) {
    const int rj = rk[set_idx[off + j]];
    const double wj = sa[rj - 1];
    int kord = 0;
    double cum = 0.0;
    for (int t = 0; t < m; t++) {
      const int rt = rk[set_idx[off + t]];
      if (rt < rj || (rt == rj && t <= j)) { kord++; cum += sa[rt - 1]; }
    }
    // G == m (possible with FEWER than G distinct genes when members repeat): the
    // reference divides by zero, NaN where r == k and +-Inf elsewhere, and max() / min()
    // return NaN as soon as one element is; 0 * Inf and n * Inf behave the same here
    const double ps = cum * invW - (double)(rj - kord) * invGm;
    const double ng = ps - wj * invW;
    has_nan |= isnan(ps) || isnan(ng);
    mx = fmax(mx, ps);
    mn = fmin(mn, ng);
  }
  mx = warp_max(mx);
  mn = warp_min(mn);
  has_nan = __any_sync(0xffffffffu, has_nan);
  if (lane == 0) {
    double es = (mx > -mn) ? mx : mn;
    const bool degenerate = W == 0.0 || m == 0 || has_nan;
    if (degenerate) es = nan("");
    es_out[(size_t)col * S + s] = es;
    if (!degenerate && (fabs(mx + mn) <= ES_TIE_REL * fmax(fabs(mx), fabs(mn)) || near_observed(obs, ident, S, R, col, s, es)))
      tie_push(tie_list, tie_ctr, tie_cap, (int)col, s);
  }
}

// es_pos == -es_neg up to rounding: the sign of the result is the reference's
// rounding, so these units are redone in the reference's own operation order
// (the arithmetic of observed_wks_kernel / R/flexgsea.R:819-838): w / sum(w) per
// member, a running sum that rounds like R's long-double cumsum (double-double
// here: exact whenever the tie is structural, e.g. equal weights), then
// d_hit - d_miss with d_miss = (r - k) / (G - m) as a division.  A warp per
// listed (column, set); every unit when the list overflowed.  O(m^2 / 32) with
// no scratch: each member sums the weights of the members ranked before it.
struct dd2 { double hi, lo; };
__device__ __forceinline__ dd2 dd2_add(dd2 a, double b) {
  const double s = __dadd_rn(a.hi, b);
  const double bb = __dsub_rn(s, a.hi);
  double e = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(b, bb));
  e = __dadd_rn(e, a.lo);
  const double hi = __dadd_rn(s, e);
  return {hi, __dsub_rn(e, __dsub_rn(hi, s))};
}
__global__ void __launch_bounds__(256)
es_wks_tie_kernel(const uint16_t *__restrict__ rank16, const double *__restrict__ sabs, int G,
                  long long C, const int *__restrict__ set_off, const int *__restrict__ set_idx, int S,
                  const int2 *__restrict__ list, const int *__restrict__ ctr, int cap,
                  double *__restrict__ es_out) {
  const int lane = threadIdx.x & 31;
  const int listed = *ctr;
  const bool all = listed > cap;
  const long long nunit = all ? (long long)S * C : listed;
  const size_t Gp = pitch_of(G);
  for (long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < nunit;
       u += (long long)gridDim.x * (blockDim.x >> 5)) {
    const long long col = all ? u / S : list[u].x;
    const int s = all ? (int)(u % S) : list[u].y;
    const uint16_t *rk = rank16 + (size_t)col * Gp;
    const double *sa = sabs + (size_t)col * Gp;
    const int off = set_off[s], m = set_off[s + 1] - off;
    dd2 tot = {0.0, 0.0};  // sum(w) in member order (R: long double)
    if (lane == 0)
      for (int j = 0; j < m; j++) tot = dd2_add(tot, sa[rk[set_idx[off + j]] - 1]);
    const double W = __shfl_sync(0xffffffffu, tot.hi, 0);
    const double denom = (double)(G - m);
    double mx = -INFINITY, mn = INFINITY;
    bool has_nan = false;
    for (int j = lane; j < m; j += 32) {
      const int rj = rk[set_idx[off + j]];
      int kord = 0;
      dd2 cum = {0.0, 0.0};
      for (int t = 0; t < m; t++) {
        const int rt = rk[set_idx[off + t]];
        if (rt < rj || (rt == rj && t <= j)) { kord++; cum = dd2_add(cum, __ddiv_rn(sa[rt - 1], W)); }
      }
      const double wo = __ddiv_rn(sa[rj - 1], W);
      const double d_miss = __ddiv_rn((double)(rj - kord), denom);
      const double ps = __dsub_rn(cum.hi, d_miss);
      const double ng = __dsub_rn(ps, wo);
      if (isnan(ps) || isnan(ng)) has_nan = true;
      if (ps > mx) mx = ps;
      if (ng < mn) mn = ng;
    }
    mx = warp_max(mx);
    mn = warp_min(mn);
    has_nan = __any_sync(0xffffffffu, has_nan);
    if (lane == 0) {
      double es = (mx > -mn) ? mx : mn;  // :833
      if (has_nan || m == 0) es = nan("");
      es_out[(size_t)col * S + s] = es;
    }
  }
}

// K5: mean / maxmean.  A warp per (set, column); scores read through L1/L2
// (a column is 8 G bytes and is re-read by every set of that column).
__global__ void __launch_bounds__(256)
es_mean_kernel(const double *__restrict__ scores, int G, long long C,
               const int *__restrict__ set_off, const int *__restrict__ set_idx, int S, int es_kind,
               int abs_flag, double *__restrict__ es_out, int2 *__restrict__ tie_list,
               int *__restrict__ tie_ctr, int tie_cap, const double *__restrict__ obs,
               const int *__restrict__ ident, int R) {
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
      if (es_kind == FGS_ES_MEAN) { sp += v; sn += fabs(v); }
      else { double a = fabs(v); sp += (v + a) / 2; sn += (-v + a) / 2; }  // :734-735
    }
    sp = warp_sum(sp);
    sn = warp_sum(sn);
    if (lane == 0) {
      double es;
      bool shaky;  // a sum that cancelled (mean) or a tie between the two sides (maxmean): redo exactly
      if (es_kind == FGS_ES_MEAN) { es = sp / m; shaky = fabs(sp) <= ES_OBS_REL * sn; }  // :763
      else {
        double ps = sp / m, ng = sn / m;
        es = fmax(ps, ng);                                     // :736
        if (isnan(ps) || isnan(ng)) es = nan("");
        else if (ng > ps) es = -es;                            // :737, strict
        shaky = fabs(ps - ng) <= ES_OBS_REL * fmax(ps, ng);
      }
      es_out[(size_t)col * S + s] = es;
      if (shaky || near_observed(obs, ident, S, R, col, s, es)) tie_push(tie_list, tie_ctr, tie_cap, (int)col, s);
    }
  }
}

// mean / maxmean of listed (column, set) units -- or of every unit when the list
// overflowed or the problem is small -- the way the reference gets them: R sums in long
// double, so the sum is exact to the last bit of the result; here double-double partial
// sums per lane, combined in double-double, one rounding at the division.  Matters where
// ES values tie exactly in the reference (rounded data, small universes, permutations that
// reproduce the observed labelling) and a plain parallel sum would break the tie at random.
__device__ __forceinline__ dd2 dd2_add_dd(dd2 a, dd2 b) { return dd2_add(dd2_add(a, b.hi), b.lo); }
__device__ __forceinline__ double dd2_div_n(dd2 a, double n) {
  const double q = a.hi / n;
  const double r = fma(-q, n, a.hi) + a.lo;
  return q + r / n;
}
__global__ void __launch_bounds__(256)
es_mean_exact_kernel(const double *__restrict__ scores, int G, long long C,
                     const int *__restrict__ set_off, const int *__restrict__ set_idx, int S,
                     int es_kind, int abs_flag, const int2 *__restrict__ list,
                     const int *__restrict__ ctr, int cap, double *__restrict__ es_out) {
  const int lane = threadIdx.x & 31;
  const int listed = *ctr;
  const bool all = listed > cap;
  const long long nunit = all ? (long long)S * C : listed;
  for (long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < nunit;
       u += (long long)gridDim.x * (blockDim.x >> 5)) {
    const long long col = all ? u / S : list[u].x;
    const int s = all ? (int)(u % S) : list[u].y;
    const double *sc = scores + (size_t)col * G;
    const int off = set_off[s], m = set_off[s + 1] - off;
    dd2 sp = {0.0, 0.0}, sn = {0.0, 0.0};
    for (int j = lane; j < m; j += 32) {
      double v = sc[set_idx[off + j]];
      if (abs_flag) v = fabs(v);
      if (es_kind == FGS_ES_MEAN) sp = dd2_add(sp, v);
      else { const double a = fabs(v); sp = dd2_add(sp, (v + a) / 2); sn = dd2_add(sn, (-v + a) / 2); }  // :734-735
    }
    for (int o = 16; o; o >>= 1) {
      dd2 t = {__shfl_xor_sync(0xffffffffu, sp.hi, o), __shfl_xor_sync(0xffffffffu, sp.lo, o)};
      sp = dd2_add_dd(sp, t);
      dd2 t2 = {__shfl_xor_sync(0xffffffffu, sn.hi, o), __shfl_xor_sync(0xffffffffu, sn.lo, o)};
      sn = dd2_add_dd(sn, t2);
    }
    if (lane == 0) {
      double es;
      if (es_kind == FGS_ES_MEAN) es = dd2_div_n(sp, (double)m);  // :763
      else {
        const double ps = dd2_div_n(sp, (double)m), ng = dd2_div_n(sn, (double)m);
        es = fmax(ps, ng);                                       // :736
        if (isnan(ps) || isnan(ng)) es = nan("");
        else if (ng > ps) es = -es;                              // :737, strict
      }
      es_out[(size_t)col * S + s] = es;
    }
  }
}
int launch_es_mean_exact(fgs_ctx *c, const double *d_scores, int G, long long C, int es_kind,
                         int abs_flag, double *d_es) {
  if (C <= 0 || c->S <= 0) return FGS_OK;
  es_mean_exact_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(d_scores, G, C, c->set_off.p, c->set_idx.p, c->S,
                                                           es_kind, abs_flag, c->tie_list.p, c->tie_ctr.p,
                                                           c->tie_cap, d_es);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

int launch_es_wks_generic(fgs_ctx *c, const uint16_t *d_rank16, const double *d_sabs, int G,
                          long long C, const int *d_set_ids, int n_ids, double *d_es) {
  if (C <= 0 || n_ids <= 0) return FGS_OK;
  long long units = (long long)n_ids * C;
  long long blocks = (units + 7) / 8;
  if (blocks > 0x7fffffffLL) { set_error("too many (set, column) units"); return FGS_ERR_ARG; }
  es_wks_generic_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(
      d_rank16, d_sabs, G, C, c->set_off.p, c->set_idx.p, d_set_ids, n_ids, c->S, d_es,
      c->tie_list.p, c->tie_ctr.p, c->tie_cap, c->cur_obs, c->cur_ident, c->cur_obs_R);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// 1 / (G - m) per set in processing order (G is a launch parameter: user-score
// calls may rank a different number of genes than the resident x)
__global__ void fill_invgm_kernel(const int4 *__restrict__ set_info, int S, int G,
                                  double *__restrict__ invgm) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < S) invgm[k] = 1.0 / (double)(G - (set_info[k].y & 0x7fffffff));
}

// Gene-set intake (fgs_set_gene_sets): a warp per set turns the 1-based ids that
// match(set, gene.names) produced into 0-based ones in place, reports the first
// set holding an id < 1 and the largest id, and flags the sets the in-register
// kernel cannot take: duplicated members (found by comparing every member with
// its predecessors, O(m^2 / 32)) or more than ES_MAX_SET members.
__global__ void __launch_bounds__(256)
prep_sets_kernel(const int *__restrict__ set_off, int *__restrict__ set_idx, int S,
                 unsigned char *__restrict__ slow, int *__restrict__ status) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= S) return;
  const int a = set_off[s], m = set_off[s + 1] - a;
  int vmax = 0, bad = 0, dup = 0;
  for (int j = lane; j < m; j += 32) {
    const int v = set_idx[a + j];
    bad |= v < 1;
    vmax = max(vmax, v);
  }
  if (m <= ES_MAX_SET) {
    for (int j = lane; j < m; j += 32) {
      const int v = set_idx[a + j];
      for (int t = 0; t < j; t++) dup |= set_idx[a + t] == v;
    }
  }
  __syncwarp();
  for (int j = lane; j < m; j += 32) set_idx[a + j] -= 1;
  bad = __any_sync(0xffffffffu, bad);
  dup = __any_sync(0xffffffffu, dup);
  vmax = __reduce_max_sync(0xffffffffu, vmax);
  if (lane == 0) {
    slow[s] = (unsigned char)(dup || m > ES_MAX_SET);
    if (bad) atomicMin(&status[0], s);
    atomicMax(&status[1], vmax);
  }
}
int launch_prep_sets(fgs_ctx *c, int S, int *h_status, unsigned char *h_slow) {
  int *status = (int *)(c->set_slow.p + (((size_t)S + 7) & ~(size_t)7));
  const int init[2] = {S, 0};
  FGS_CUDA(cudaMemcpyAsync(status, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  prep_sets_kernel<<<ceil_div(S, 8), 256, 0, c->stream>>>(c->set_off.p, c->set_idx.p, S, c->set_slow.p, status);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  FGS_CUDA(cudaMemcpyAsync(h_status, status, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  FGS_CUDA(cudaMemcpyAsync(h_slow, c->set_slow.p, (size_t)S, cudaMemcpyDeviceToHost, c->stream));
  FGS_CUDA(cudaStreamSynchronize(c->stream));
  return FGS_OK;
}

// 1 / (G - m) per set and the uint16 member blocks, both functions of G: once per (gene sets, G)
static int ensure_set_blocks(fgs_ctx *c, int G) {
  if (c->invgm_G == G) return FGS_OK;
  FGS_TRY(c->set_invgm.ensure(c->S));
  fill_invgm_kernel<<<ceil_div(c->S, 256), 256, 0, c->stream>>>(c->set_info.p, c->S, G, c->set_invgm.p);
  FGS_TRY(c->set_idx16.ensure((size_t)std::max<long long>(c->idx16_blocks, 1) * 8));
  build_idx16_kernel<<<ceil_div(c->S, 8), 256, 0, c->stream>>>(c->set_info.p, c->set_idx.p, c->S, G,
                                                               c->set_idx16.p);
  c->launches += 2;
  if (c->opt_arrange) {  // the bank-aware order for the maxmean / mean gathers (pairs of sets)
    const size_t smem = (size_t)8 * ARR_MAX * sizeof(uint16_t);
    FGS_CUDA(cudaFuncSetAttribute(arrange_idx16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    arrange_idx16_kernel<<<ceil_div((c->S + 1) / 2, 8), 256, smem, c->stream>>>(c->set_info.p, c->set_idx.p, c->S, G,
                                                                                c->set_idx16.p);
    c->launches++;
  }
  FGS_CUDA(cudaGetLastError());
  c->invgm_G = G;
  return FGS_OK;
}

static int launch_es_wks_inner(fgs_ctx *c, const uint16_t *d_rank16, const double *d_sabs, int G,
                               long long C, double *d_es) {
  if (c->opt_force_generic)
    return launch_es_wks_generic(c, d_rank16, d_sabs, G, C, nullptr, c->S, d_es);
  const size_t Gp = pitch_of(G);
  const size_t budget = (size_t)c->smem_optin;
  if (Gp * 2 + 16 + 48 > budget)  // (48: the kernel's static shared memory) the rank array does not fit on chip: the order-free kernel does everything
    return launch_es_wks_generic(c, d_rank16, d_sabs, G, C, nullptr, c->S, d_es);
  if (G > 65534)  // the pad key G must fit a uint16 next to the ranks
    return launch_es_wks_generic(c, d_rank16, d_sabs, G, C, nullptr, c->S, d_es);
  const int grid = c->num_sms;
  const int tail = (int)(C % grid);
  const int tail_split = (tail && c->opt_tail_split) ? std::max(1, std::min(8, grid / tail)) : 1;
  // rank (2 B) + weight (8 B) per gene on chip, + pad gene, + zero weight slots
  const bool w_smem = Gp * 10 + 16 + 64 + 48 <= budget;  // 48: the kernel's static shared memory
  const size_t smem = w_smem ? Gp * 10 + 16 + 64 : Gp * 2 + 16;
  FGS_TRY(ensure_set_blocks(c, G));
  // groups [0, n_wide) are single sets on a whole warp, the rest packed sets of up to 512 members
  auto go = [&](auto kern, int threads, const int2 *groups, int n) -> int {
    if (n <= 0) return FGS_OK;
    FGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, c->stream>>>(d_rank16, d_sabs, G, C, c->set_info.p, c->set_idx16.p,
                                             c->set_invgm.p, c->S, d_es, c->tie_list.p, c->tie_ctr.p,
                                             c->tie_cap, c->cur_obs, c->cur_ident, c->cur_obs_R, groups, n,
                                             tail_split, (c->opt_bulk_stage && G >= 2) ? 1 : 0);
    c->launches++;
    FGS_CUDA(cudaGetLastError());
    return FGS_OK;
  };
  const int2 *gw = c->set_groups.p, *gp = c->set_groups.p + c->n_wide_groups;
  const int n_packed = c->n_groups - c->n_wide_groups;
  if (w_smem) {
    FGS_TRY(go(es_wks_kernel<true, false>, ES_THREADS, gp, n_packed));
    FGS_TRY(go(es_wks_kernel<true, true>, ES_WIDE_THREADS, gw, c->n_wide_groups));
  } else {
    FGS_TRY(go(es_wks_kernel<false, false>, ES_LONG_THREADS, gp, n_packed));
    FGS_TRY(go(es_wks_kernel<false, true>, ES_WIDE_THREADS, gw, c->n_wide_groups));
  }
  if (c->n_slow > 0)
    FGS_TRY(launch_es_wks_generic(c, d_rank16, d_sabs, G, C, c->slow_ids.p, c->n_slow, d_es));
  return FGS_OK;
}

// The member genes of the tied units as (column, gene) entries, for
// s2n_exact_entries_kernel: with tensor-core scores the weights are only ~1e-13
// from the reference's, and at a structural tie the sign of the result hangs on
// their last bit.  Overflow (of this list or of the tie list) raises the flag
// that makes the caller rerun the call with the bit-faithful score kernel.
__global__ void __launch_bounds__(256)
tie_members_kernel(const int2 *__restrict__ tie_list, const int *__restrict__ tie_ctr, int tie_cap,
                   const int *__restrict__ set_off, const int *__restrict__ set_idx,
                   int2 *__restrict__ out, int *__restrict__ out_ctr, int out_cap,
                   int *__restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const int listed = *tie_ctr;
  if (listed > tie_cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(overflow, 16);
    return;
  }
  for (int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < listed;
       u += gridDim.x * (blockDim.x >> 5)) {
    const int col = tie_list[u].x, s = tie_list[u].y;
    const int off = set_off[s], m = set_off[s + 1] - off;
    int base = 0;
    if (lane == 0) base = atomicAdd(out_ctr, m);
    base = __shfl_sync(0xffffffffu, base, 0);
    if ((long long)base + m > out_cap) {
      if (lane == 0) atomicOr(overflow, 16);
      continue;
    }
    for (int j = lane; j < m; j += 32) out[base + j] = make_int2(col, set_idx[off + j]);
  }
}
// ... and their hit weights rewritten from the exact scores (ranks are certified, so
// the rank-ordered weight array only changes in value)
__global__ void tie_refresh_weights_kernel(const int2 *__restrict__ list, const int *__restrict__ ctr,
                                           int cap, const double *__restrict__ scores,
                                           const uint16_t *__restrict__ rank16, double *__restrict__ sabs,
                                           int G, double p_exp) {
  const int n = min(*ctr, cap);
  const size_t Gp = pitch_of(G);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int col = list[i].x, g = list[i].y;
    double w = fabs(scores[(size_t)col * G + g]);
    w = hit_weight(w, p_exp);
    sabs[(size_t)col * Gp + rank16[(size_t)col * Gp + g] - 1] = w;
  }
}
int launch_tie_members(fgs_ctx *c, int2 *d_out, int *d_out_ctr, int out_cap, int *d_overflow) {
  tie_members_kernel<<<c->num_sms * 2, 256, 0, c->stream>>>(c->tie_list.p, c->tie_ctr.p, c->tie_cap,
                                                            c->set_off.p, c->set_idx.p, d_out, d_out_ctr,
                                                            out_cap, d_overflow);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}
int launch_tie_refresh_weights(fgs_ctx *c, const int2 *d_list, const int *d_ctr, int cap,
                               const double *d_scores, const uint16_t *d_rank16, double *d_sabs, int G,
                               double p_exp) {
  tie_refresh_weights_kernel<<<c->num_sms * 2, 256, 0, c->stream>>>(d_list, d_ctr, cap, d_scores, d_rank16,
                                                                    d_sabs, G, p_exp);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// es_null[, r, p] = observed ES[, r] for the label-preserving permutations of a block
// (a block row per permutation: the others leave at once)
__global__ void copy_observed_kernel(const double *__restrict__ obs, const int *__restrict__ ident, int S,
                                     int R, double *__restrict__ es_out) {
  const int p = blockIdx.y;
  if (!ident[p]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (set, response) of the observed matrix
  if (i < S * R) es_out[(size_t)p * R * S + i] = obs[i];
}
int launch_copy_observed(fgs_ctx *c, const double *d_obs, const int *d_ident, int R, long long C,
                         double *d_es) {
  const long long Pb = C / R;
  if (Pb <= 0 || c->S <= 0) return FGS_OK;
  for (long long p0 = 0; p0 < Pb; p0 += 65535) {  // grid.y limit
    const int np = (int)std::min<long long>(65535, Pb - p0);
    dim3 grid(ceil_div((long long)c->S * R, 256), np);
    copy_observed_kernel<<<grid, 256, 0, c->stream>>>(d_obs, d_ident + p0, c->S, R, d_es + (size_t)p0 * R * c->S);
  }
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// the exact redo of the units whose two extremes tied (see es_wks_tie_kernel)
int launch_es_wks_ties(fgs_ctx *c, const uint16_t *d_rank16, const double *d_sabs, int G, long long C,
                       double *d_es) {
  if (C <= 0 || c->S <= 0) return FGS_OK;
  int grid = c->num_sms * 4;  // grid-stride over the device-side count (usually zero)
  es_wks_tie_kernel<<<grid, 256, 0, c->stream>>>(d_rank16, d_sabs, G, C, c->set_off.p, c->set_idx.p, c->S,
                                                 c->tie_list.p, c->tie_ctr.p, c->tie_cap, d_es);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

__global__ void set_int_kernel(int *p, int v) { *p = v; }

// Weighted KS of C score columns: the in-register kernel (or the order-free one),
// then -- unless the caller first wants to refresh the tied units' weights
// (defer_ties) -- the exact redo of the units whose two extremes tied.
int launch_es_wks(fgs_ctx *c, const uint16_t *d_rank16, const double *d_sabs, int G, long long C,
                  double *d_es, bool defer_ties) {
  if (C <= 0 || c->S <= 0) return FGS_OK;
  FGS_TRY(ensure_tie_list(c));
  if (c->cur_exact_all) {  // a small problem: every unit by es_wks_tie_kernel (a count above the cap means "all")
    set_int_kernel<<<1, 1, 0, c->stream>>>(c->tie_ctr.p, c->tie_cap + 1);
    c->launches++;
    return launch_es_wks_ties(c, d_rank16, d_sabs, G, C, d_es);
  }
  FGS_TRY(launch_es_wks_inner(c, d_rank16, d_sabs, G, C, d_es));
  if (!defer_ties) FGS_TRY(launch_es_wks_ties(c, d_rank16, d_sabs, G, C, d_es));
  return FGS_OK;
}

// K5 on chip: persistent CTA per SM, the score column (8 G bytes) staged in shared
// memory; eight lanes per set, four sets per warp (descending size, quads
// round-robin over the warps).  A lane takes eight members per 16-byte load of the
// uint16 member blocks (pads point at a zero slot behind the column).  Per member
// one or two FP64 adds: sum(v) and, for maxmean, sum(|v|); then
//   mean    = sum(v) / m                                            (:763)
//   maxmean : pos = (sum|v| + sum v) / 2m, neg = (sum|v| - sum v) / 2m (:734-735),
//             es = max(pos, neg), negated iff neg > pos               (:736-737)
constexpr int ESM_THREADS = 1024;
template <bool MAXMEAN>
__global__ void __launch_bounds__(ESM_THREADS, 1)
es_mean_smem_kernel(const double *__restrict__ scores, int G, long long C,
                    const int4 *__restrict__ set_info, const uint16_t *__restrict__ idx16, int S,
                    int abs_flag, double *__restrict__ es_out, int2 *__restrict__ tie_list,
                    int *__restrict__ tie_ctr, int tie_cap, const double *__restrict__ obs,
                    const int *__restrict__ ident, int R, int tail_split, int bulk) {
  extern __shared__ __align__(16) unsigned char smem[];
  double *s_sc = (double *)smem;  // [G + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int seg = lane >> 3, ls = lane & 7;
  const int nquads = (S + 3) >> 2;
  __shared__ __align__(8) uint64_t s_bar;
  // the column is a plain copy unless |.| is taken on the way in: one bulk copy by the copy engine
  // (columns have pitch G: 16-byte aligned when G is even)
  bulk = bulk && !abs_flag && !(G & 1);
  if (tid == 0) { s_sc[G] = 0.0; if (bulk) bulk_bar_init(&s_bar); }  // the pad gene
  uint32_t phase = 0;
  const long long C_full = (C / gridDim.x) * gridDim.x;  // then the tail round, split (see es_wks_kernel)
  for (long long w = blockIdx.x;; w += gridDim.x) {
    long long col = w;
    int part = 0, nparts = 1;
    if (w >= C_full) {
      const long long t = w - C_full;
      if (t >= (C - C_full) * tail_split) break;
      col = C_full + t / tail_split; part = (int)(t % tail_split); nparts = tail_split;
    }
    __syncthreads();
    const double *sc = scores + (size_t)col * G;
    if (bulk) {
      if (tid == 0) bulk_g2s(s_sc, sc, (uint32_t)G * 8u, &s_bar);
      bulk_wait(&s_bar, phase);
      phase ^= 1u;
    } else {
      for (int i = tid; i < G; i += blockDim.x) {
        const double v = __ldg(sc + i);
        s_sc[i] = abs_flag ? fabs(v) : v;
      }
    }
    __syncthreads();
    double *es_col = es_out + (size_t)col * S;
    // observed ES of this column's response; none for a label-preserving permutation (copied later)
    const double *col_obs = (obs && !(ident && __ldg(ident + col / R))) ? obs + (size_t)(col % R) * S : nullptr;
    for (int q = warp * nparts + part; q < nquads; q += nwarps * nparts) {
      const int k = 4 * q + seg;
      int4 info = make_int4(0, 0, 0, 0);
      if (k < S) info = __ldg(set_info + k);
      const double e_obs = (col_obs && k < S) ? __ldg(col_obs + info.w) : nan("");
      const int m = info.y & 0x7fffffff;  // duplicates are just members here
      const int nblk = (m + 7) >> 3;
      int nb = (nblk + 7) >> 3;           // 16-byte loads per lane, the same count for the whole warp
      nb = max(nb, __shfl_xor_sync(0xffffffffu, nb, 8));
      nb = max(nb, __shfl_xor_sync(0xffffffffu, nb, 16));
      const uint4 *blocks = (const uint4 *)(idx16 + (size_t)info.z * 8);
      double sv0 = 0.0, sv1 = 0.0, sa0 = 0.0, sa1 = 0.0;
      for (int b = 0; b < nb; b++) {
        const int blk = b * 8 + ls;
        if (blk < nblk) {
          const uint4 w = __ldg(blocks + blk);
          const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const double v0 = s_sc[ww[i] & 0xffffu], v1 = s_sc[ww[i] >> 16];
            sv0 += v0; sv1 += v1;
            sa0 += fabs(v0); sa1 += fabs(v1);  // (mean: only to notice a sum that cancelled)
          }
        }
      }
      double sv = sv0 + sv1, sa = sa0 + sa1;
#pragma unroll
      for (int o = 4; o; o >>= 1) {
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
      }
      if (ls == 0 && k < S) {
        double es;
        bool shaky;  // a sum that cancelled (mean) or a tie between the two sides (maxmean): redo exactly
        if (!MAXMEAN) { es = sv / m; shaky = fabs(sv) <= ES_OBS_REL * sa; }
        else {
          const double ps = (sa + sv) / 2 / m, ng = (sa - sv) / 2 / m;
          es = fmax(ps, ng);
          if (isnan(ps) || isnan(ng)) es = nan("");
          else if (ng > ps) es = -es;
          shaky = fabs(ps - ng) <= ES_OBS_REL * fmax(ps, ng);
        }
        es_col[info.w] = es;
        if (shaky || fabs(es - e_obs) <= ES_OBS_REL * fabs(e_obs)) tie_push(tie_list, tie_ctr, tie_cap, (int)col, info.w);
      }
    }
    if (w >= C_full) break;
  }
}

int ensure_tie_list(fgs_ctx *c) {
  if (c->tie_cap == 0) {
    c->tie_cap = 1 << 22;  // 32 MB of (column, set) pairs per block; beyond that every unit is redone
    FGS_TRY(c->tie_list.ensure((size_t)c->tie_cap));
    FGS_TRY(c->tie_ctr.ensure(4));
  }
  FGS_CUDA(cudaMemsetAsync(c->tie_ctr.p, 0, 4 * sizeof(int), c->stream));
  return FGS_OK;
}

// maxmean / mean of C score columns.  Units that land within rounding of the observed ES are
// listed (the caller redoes them with launch_es_mean_exact); a small problem is done exactly
// from the start.
int launch_es_mean(fgs_ctx *c, const double *d_scores, int G, long long C, int es_kind,
                   int abs_flag, double *d_es) {
  if (C <= 0 || c->S <= 0) return FGS_OK;
  FGS_TRY(ensure_tie_list(c));
  if (c->cur_exact_all) {
    set_int_kernel<<<1, 1, 0, c->stream>>>(c->tie_ctr.p, c->tie_cap + 1);
    c->launches++;
    return FGS_OK;  // launch_es_mean_exact (all units) follows
  }
  const size_t smem = ((size_t)G + 2) * 8;
  if (smem + 32 <= (size_t)c->smem_optin && G <= 65534) {  // (32: the kernel's static shared memory)
    const int grid = c->num_sms;
    const int tail = (int)(C % grid);
    const int tail_split = (tail && c->opt_tail_split) ? std::max(1, std::min(8, grid / tail)) : 1;
    FGS_TRY(ensure_set_blocks(c, G));
    if (es_kind == FGS_ES_MAXMEAN) {
      auto k = es_mean_smem_kernel<true>;
      FGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, ESM_THREADS, smem, c->stream>>>(d_scores, G, C, c->set_info.p, c->set_idx16.p, c->S, abs_flag, d_es,
                                                c->tie_list.p, c->tie_ctr.p, c->tie_cap, c->cur_obs, c->cur_ident, c->cur_obs_R, tail_split,
                                                (c->opt_bulk_stage && G >= 2) ? 1 : 0);
    } else {
      auto k = es_mean_smem_kernel<false>;
      FGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, ESM_THREADS, smem, c->stream>>>(d_scores, G, C, c->set_info.p, c->set_idx16.p, c->S, abs_flag, d_es,
                                                c->tie_list.p, c->tie_ctr.p, c->tie_cap, c->cur_obs, c->cur_ident, c->cur_obs_R, tail_split,
                                                (c->opt_bulk_stage && G >= 2) ? 1 : 0);
    }
  } else {  // column does not fit on chip: gather through L1/L2
    long long units = (long long)c->S * C;
    long long blocks = (units + 7) / 8;
    long long cap = (long long)c->num_sms * 32;
    if (blocks > cap) blocks = cap;
    es_mean_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(d_scores, G, C, c->set_off.p,
                                                            c->set_idx.p, c->S, es_kind, abs_flag,
                                                            d_es, c->tie_list.p, c->tie_ctr.p, c->tie_cap,
                                                            c->cur_obs, c->cur_ident, c->cur_obs_R);
  }
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
