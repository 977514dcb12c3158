// K6 / K7: significance of the observed ES against the resident null.
//
// Replaces flexgsea_calc_sig (R/flexgsea.R:564-629), calc_fdr_nes (:500-533),
// adj_fdr_nes (:474-498) and flexgsea_calc_sig_simple (:544-547).
//
// The reference's pooled FDR is O(S^2 P): every set rescans the whole S x P
// null-NES matrix (:512-528).  Here the null is streamed ONCE: each null value
// is normalised on the fly (:619-621) and binary-searched into the sorted
// observed NES; exact integer histogram counts + a suffix/prefix sum give every
// set's exceedance count.  All counts are integers, so the result is
// independent of thread order.  Per-set sums of the null (for the NES means,
// :602-616) are accumulated in double-double and combined in a fixed order, so
// the means round like R's long-double mean() and are reproducible across any
// split of the permutations over chunks or GPUs.
#include <algorithm>

#include "fgs_common.cuh"

namespace fgs {

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
__device__ __forceinline__ dd dd_add_dd(dd a, dd b) {
  dd r = dd_add_d(a, b.hi);
  return dd_add_d(r, b.lo);
}
// (hi + lo) / n rounded to double
__device__ __forceinline__ double dd_div_n(dd a, double n) {
  double q = a.hi / n;
  double r = fma(-q, n, a.hi) + a.lo;
  return q + r / n;
}

// A null ES within rounding of the set's observed ES.  Such values come from
// permutations that reproduce the observed labelling (common with few samples) or,
// in small gene universes, from the few values the statistic can take; the reference
// computes the observed and the null ES with the same code, so whether they compare
// equal is decided by ITS rounding.  When the observed ES was known during the null
// loop (fgs_observed_es was called first, as flexgsea() does), the ES kernels
// recomputed every such unit in the reference's operation order and nothing is done
// here (snap_rel = 0).  Otherwise the null value is taken as equal to the observed one.
constexpr double SIG_SNAP_REL = 1e-11;
__device__ __forceinline__ double snap_to_observed(double v, double e, double snap_rel) {
  return fabs(v - e) <= snap_rel * fabs(e) ? e : v;
}

constexpr int SIG_NF = 4;  // sums per set: pos hi, pos lo, neg hi, neg lo
constexpr int SIG_NC = 6;  // counts per set

// ---- stage 1: per-set partials over a chunk of permutations -----------------
__global__ void __launch_bounds__(128)
sig_partial_kernel(const double *__restrict__ null, int S, int R, int P, int resp,
                   const double *__restrict__ es, int split_p, int abs_flag, int chunk_len,
                   double snap_rel, double *__restrict__ psums /* [nchunk][S][4] */,
                   long long *__restrict__ pcounts /* [nchunk][S][6] */) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int chunk = blockIdx.y;
  const int p0 = chunk * chunk_len, p1 = min(P, p0 + chunk_len);
  const double e = es[s];
  const bool pos_side = (e >= 0.0) || abs_flag;
  dd sp = {0.0, 0.0}, sn = {0.0, 0.0};
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;  // a chunk is far below 2^31 permutations
  const double *base = null + (size_t)s + (size_t)S * resp;
  const size_t stride = (size_t)S * R;
  // one streaming read of the null (8 S P bytes): SIG_U loads in flight per thread, and ONE
  // double-double add per value -- the accumulator of the value's sign is selected, updated and
  // put back (written as two guarded adds the compiler computes both and selects: twice the FP64 work)
  constexpr int SIG_U = 8;
  for (int pb = p0; pb < p1; pb += SIG_U) {
    double vv[SIG_U];
#pragma unroll
    for (int u = 0; u < SIG_U; u++) vv[u] = (pb + u < p1) ? __ldg(base + stride * (pb + u)) : nan("");
#pragma unroll
    for (int u = 0; u < SIG_U; u++) {
      const double v = snap_to_observed(vv[u], e, snap_rel);
      const bool nonneg = v >= 0.0, neg = v < 0.0;  // NaN (and the padding of the last group): neither
      c0 += nonneg; c1 += neg; c2 += (v <= 0.0);
      dd acc = nonneg ? sp : sn;
      acc = dd_add_d(acc, (nonneg || neg) ? v : 0.0);
      if (nonneg) sp = acc;
      if (neg) sn = acc;
      if (split_p) {  // :579-587
        c3 += pos_side && nonneg && e <= v;
        c4 += !pos_side && (v <= 0.0) && e >= v;
      } else {        // :589-597
        c3 += (e <= v);
        c4 += (e >= v);
      }
    }
  }
  double *ps = psums + ((size_t)chunk * S + s) * SIG_NF;
  ps[0] = sp.hi; ps[1] = sp.lo; ps[2] = sn.hi; ps[3] = sn.lo;
  long long *pc = pcounts + ((size_t)chunk * S + s) * SIG_NC;
  pc[0] = c0; pc[1] = c1; pc[2] = c2; pc[3] = c3; pc[4] = c4; pc[5] = 0;
}

__global__ void sig_combine_kernel(const double *__restrict__ psums,
                                   const long long *__restrict__ pcounts, int S, int nchunk,
                                   double *__restrict__ sums, long long *__restrict__ counts) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  dd sp = {0.0, 0.0}, sn = {0.0, 0.0};
  long long c[SIG_NC] = {0, 0, 0, 0, 0, 0};
  for (int k = 0; k < nchunk; k++) {  // fixed order
    const double *ps = psums + ((size_t)k * S + s) * SIG_NF;
    sp = dd_add_dd(sp, {ps[0], ps[1]});
    sn = dd_add_dd(sn, {ps[2], ps[3]});
    const long long *pc = pcounts + ((size_t)k * S + s) * SIG_NC;
    for (int i = 0; i < SIG_NC; i++) c[i] += pc[i];
  }
  if (sums) { double *o = sums + (size_t)s * SIG_NF; o[0] = sp.hi; o[1] = sp.lo; o[2] = sn.hi; o[3] = sn.lo; }
  if (counts) for (int i = 0; i < SIG_NC; i++) counts[(size_t)s * SIG_NC + i] = c[i];
}

int launch_sig_stage1(fgs_ctx *c, const double *d_null, int S, int R, int P, int response,
                      const double *d_es, int split_p, int abs_flag, double *d_sums,
                      long long *d_counts) {
  int nchunk = P / 64;
  int want = (c->num_sms * 16 * 128) / (S > 0 ? S : 1) + 1;
  if (nchunk > want) nchunk = want;
  if (nchunk < 1) nchunk = 1;
  if (nchunk > 1024) nchunk = 1024;
  int chunk_len = (P + nchunk - 1) / nchunk;
  if (chunk_len < 1) chunk_len = 1;
  nchunk = P > 0 ? (P + chunk_len - 1) / chunk_len : 1;
  FGS_TRY(c->sig_d.ensure((size_t)nchunk * S * SIG_NF));
  FGS_TRY(c->sig_i.ensure((size_t)nchunk * S * SIG_NC));
  dim3 grid(ceil_div(S, 128), nchunk);
  sig_partial_kernel<<<grid, 128, 0, c->stream>>>(d_null, S, R, P, response, d_es, split_p,
                                                  abs_flag, chunk_len,
                                                  c->null_exact_near_obs ? 0.0 : SIG_SNAP_REL, c->sig_d.p,
                                                  c->sig_i.p);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  sig_combine_kernel<<<ceil_div(S, 128), 128, 0, c->stream>>>(c->sig_d.p, c->sig_i.p, S, nchunk,
                                                              d_sums, d_counts);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// ---- means, p-values, NES, FWER from combined sums / counts ----------------
// sums_parts [nparts][S][4] (one part per GPU, combined here in part order),
// counts [S][6] (already summed).  Outputs [S].
// counts[s][k] = sum over the parts of the counts every rank appended to its sums (one all-gather
// carries both: part layout [S x 4 doubles | S x 6 int64], `stride` doubles apart)
__global__ void sig_sum_counts_kernel(const double *__restrict__ parts, int nparts, size_t stride, int S,
                                      long long *__restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * SIG_NC) return;
  long long t = 0;
  for (int k = 0; k < nparts; k++) t += ((const long long *)(parts + (size_t)k * stride + (size_t)S * SIG_NF))[i];
  counts[i] = t;
}
int launch_sig_sum_counts(fgs_ctx *c, const double *d_parts, int nparts, size_t stride, int S, long long *d_counts) {
  sig_sum_counts_kernel<<<ceil_div((long long)S * SIG_NC, 256), 256, 0, c->stream>>>(d_parts, nparts, stride, S, d_counts);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

__global__ void sig_means_kernel(const double *__restrict__ sums_parts, int nparts, size_t stride,
                                 const long long *__restrict__ counts,
                                 const double *__restrict__ es, int S, long long P_total,
                                 int split_p, int calc_nes, int abs_flag, double *__restrict__ p_out,
                                 double *__restrict__ p_high, double *__restrict__ p_low,
                                 double *__restrict__ mpos_out, double *__restrict__ mneg_out,
                                 double *__restrict__ nes_out, double *__restrict__ fwer_out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const double e = es[s];
  const long long *c = counts + (size_t)s * SIG_NC;
  double p;
  if (split_p) {
    if ((e >= 0.0) || abs_flag) p = (double)(c[3] + 1) / (double)(c[0] + 1);  // :580-582
    else p = (double)(c[4] + 1) / (double)(c[2] + 1);                         // :584-585
    p_high[s] = nan(""); p_low[s] = nan("");
  } else {
    const double ph = (double)(c[3] + 1) / (double)(P_total + 1);             // :590
    const double pl = (double)(c[4] + 1) / (double)(P_total + 1);             // :596
    p_high[s] = ph;
    if (abs_flag) { p = ph; p_low[s] = nan(""); }
    else { p_low[s] = pl; p = fmin(fmin(pl, ph) * 2, 1.0); }                  // :598
  }
  p_out[s] = p;
  fwer_out[s] = fmin(p * (double)S, 1.0);                                     // :627
  if (calc_nes) {
    dd sp = {0.0, 0.0}, sn = {0.0, 0.0};
    for (int k = 0; k < nparts; k++) {
      const double *ps = sums_parts + (size_t)k * stride + (size_t)s * SIG_NF;
      sp = dd_add_dd(sp, {ps[0], ps[1]});
      sn = dd_add_dd(sn, {ps[2], ps[3]});
    }
    double mp = c[0] > 0 ? dd_div_n(sp, (double)c[0]) : nan("");              // :602
    if (isnan(mp) && e >= 0.0) mp = e;                                        // :603-604
    double mg = c[1] > 0 ? dd_div_n(sn, (double)c[1]) : nan("");              // :612-614
    if (isnan(mg) && e < 0.0) mg = e;                                         // :615-616
    mpos_out[s] = mp; mneg_out[s] = mg;
    nes_out[s] = abs_flag ? e / mp : e / (e >= 0.0 ? mp : -mg);               // :607, :617-618
  } else {
    mpos_out[s] = nan(""); mneg_out[s] = nan(""); nes_out[s] = nan("");
  }
}

// ---- ordering helper: stable position of every element within its group ----
// group(i) in {0 = none, 1, 2}; position among same-group elements by
// (key ascending, index ascending).  S^2 compares, tiled: a block owns 128
// elements and one chunk of the partners (staged in shared memory), partial
// positions are added with integer atomics.  pos[] and group_n[] start at zero.
constexpr int RANK_TI = 128, RANK_TJ = 512;
__global__ void __launch_bounds__(RANK_TI)
group_rank_kernel(const double *__restrict__ key, const int *__restrict__ group, int n, int chunk,
                  int *__restrict__ pos, int *__restrict__ group_n) {
  __shared__ double sk[RANK_TJ];
  __shared__ int sg[RANK_TJ];
  const int i = blockIdx.x * RANK_TI + threadIdx.x;
  const bool valid = i < n;
  const int gi = valid ? group[i] : -1;
  const double ki = valid ? key[i] : 0.0;
  const int j0 = blockIdx.y * chunk, j1 = min(n, j0 + chunk);
  int r = 0;
  for (int jt = j0; jt < j1; jt += RANK_TJ) {
    const int len = min(RANK_TJ, j1 - jt);
    __syncthreads();
    for (int t = threadIdx.x; t < len; t += RANK_TI) { sk[t] = key[jt + t]; sg[t] = group[jt + t]; }
    __syncthreads();
    for (int t = 0; t < len; t++) {
      const double kj = sk[t];
      r += (sg[t] == gi) && ((kj < ki) || (kj == ki && jt + t < i));
    }
  }
  if (valid && r) atomicAdd(&pos[i], r);
  if (valid && blockIdx.y == 0 && gi > 0) atomicAdd(&group_n[gi], 1);
}
static void launch_group_rank(fgs_ctx *c, const double *key, const int *group, int n, int *pos, int *gn) {
  cudaMemsetAsync(pos, 0, (size_t)n * sizeof(int), c->stream);
  int gy = ceil_div(n, RANK_TJ);
  if (gy > 32) gy = 32;
  const int chunk = ceil_div(ceil_div(n, gy), RANK_TJ) * RANK_TJ;
  gy = ceil_div(n, chunk);
  dim3 grid(ceil_div(n, RANK_TI), gy);
  group_rank_kernel<<<grid, RANK_TI, 0, c->stream>>>(key, group, n, chunk, pos, gn);
}

// thresholds for the pooled-FDR count: group 1 = sets compared with '>'
// (nes >= 0, or abs; :518), group 2 = sets compared with '<'; NaN -> none
__global__ void fdr_groups_kernel(const double *__restrict__ nes, int S, int abs_flag,
                                  int *__restrict__ group) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  const double v = nes[i];
  group[i] = isnan(v) ? 0 : (((v >= 0.0) || abs_flag) ? 1 : 2);
}
__global__ void scatter_sorted_kernel(const double *__restrict__ key, const int *__restrict__ group,
                                      const int *__restrict__ pos, int n,
                                      double *__restrict__ sorted1, int *__restrict__ id1,
                                      double *__restrict__ sorted2, int *__restrict__ id2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (group[i] == 1) { sorted1[pos[i]] = key[i]; id1[pos[i]] = i; }
  else if (group[i] == 2) { sorted2[pos[i]] = key[i]; id2[pos[i]] = i; }
}

// ---- stage 2: one streaming pass over the local null -------------------------
// hist1[b] += 1 where b = #{thresholds1 <  v'};  set at sorted position j gets
//   #{v' > t_j} = sum_{b > j} hist1[b]
// hist2[b] += 1 where b = #{thresholds2 <= v'}; set at sorted position j gets
//   #{v' < t_j} = sum_{b <= j} hist2[b]
// Cell index over the sorted threshold list: `nc` equal cells between its smallest and largest
// value, lut[c] = #{t < edge(c)}, so a value's count is bracketed by lut[c], lut[c + 1].  With
// cells far finer than the typical gap between thresholds (65 536 cells for 8 000 thresholds)
// nearly every cell holds at most one threshold and the count is one comparison; fuller cells
// finish with a bisection over the cell's population.  uint16 entries (n < 65 536).
constexpr int FDR_NC_MAX = 65536, FDR_NC_MIN = 4096;
__device__ __forceinline__ double lut_edge(double t0, double w, int c) { return __fma_rn((double)c, w, t0); }
__global__ void fdr_lut_kernel(const double *__restrict__ t, int n, int nc, uint16_t *__restrict__ lut) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nc) return;
  if (n <= 0) { lut[c] = 0; return; }
  if (c == nc) { lut[c] = (uint16_t)n; return; }
  const double t0 = t[0], w = (t[n - 1] - t[0]) / nc;
  const double x = lut_edge(t0, w, c);
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (t[mid] < x) lo = mid + 1; else hi = mid;
  }
  lut[c] = (uint16_t)lo;
}
// #{t < v} (nonstrict = false) or #{t <= v} (nonstrict = true) through the cell index
__device__ __forceinline__ int lut_count(const double *__restrict__ a, int n, const uint16_t *__restrict__ lut,
                                         int nc, double t0, double w, double invw, double tmax, double v,
                                         bool nonstrict) {
  if (n == 0 || v < t0) return 0;
  if (v > tmax) return n;
  int lo = 0, hi = n;
  if (w > 0.0 && w < INFINITY) {  // (a single distinct threshold or an infinite one: plain bisection)
    int c = min(max((int)((v - t0) * invw), 0), nc - 1);
    while (c > 0 && v < lut_edge(t0, w, c)) c--;                  // rounding of the cell guess
    while (c < nc - 1 && v >= lut_edge(t0, w, c + 1)) c++;
    lo = lut[c]; hi = lut[c + 1];    // #{t < edge(c)} <= either count <= #{t < edge(c + 1)}
  }
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const double t = a[mid];
    if ((t < v) || (nonstrict && t == v)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// The count for a value known only approximately: `va` is within FDR_APPROX_REL of the exact
// normalised value.  The count is a step function of the value with steps at the thresholds, so it
// is the exact value's count unless a threshold lies within that distance of `va`; `sure` says so.
constexpr double FDR_APPROX_REL = 1e-15;  // v * fl(1 / m) is within 2 ulp (4.5e-16) of fl(v / m)

// ONE sorted list serves both sides: T = [thresholds of group 2 (all negative), ascending] followed
// by [thresholds of group 1 (all >= 0; every set in abs mode), ascending].  A non-negative (or abs)
// value nv needs #{t1 < nv} = #{T < nv} - n2; a negative one #{t2 <= nv} = #{T <= nv}.  Both land in
// one histogram H[n + 2]: a value with count B goes to H[B] (negative side: hist2[b] = H[b]) or
// H[B + 1] (positive side: hist1[b] = H[n2 + 1 + b]) -- no branch on the sign of the value, so a
// warp whose lanes hold values of both signs runs one instruction stream, not two.
constexpr int FDR_T_SMEM = 1024, FDR_T_GLOBAL = 256;  // threads per CTA of the two variants
template <bool T_SMEM>
__global__ void __launch_bounds__(T_SMEM ? FDR_T_SMEM : FDR_T_GLOBAL)
fdr_count_kernel(const double *__restrict__ null, int S, int R, int P, int resp, int abs_flag,
                 const double *__restrict__ es, const double *__restrict__ mpos,
                 const double *__restrict__ mneg, const double *__restrict__ T, int n, int n2,
                 const uint16_t *__restrict__ lut_g, int nc, int chunk_len, double snap_rel,
                 unsigned long long *__restrict__ hist, unsigned long long *__restrict__ glob) {
  extern __shared__ __align__(16) unsigned char smem[];
  double *st = (double *)smem;
  unsigned int *h = (unsigned int *)(st + (T_SMEM ? n : 0));
  uint16_t *lut = (uint16_t *)(h + (n + 2));
  for (int i = threadIdx.x; i < n + 2; i += blockDim.x) h[i] = 0;
  for (int i = threadIdx.x; i < nc + 1; i += blockDim.x) lut[i] = lut_g[i];
  if (T_SMEM)
    for (int i = threadIdx.x; i < n; i += blockDim.x) st[i] = T[i];
  __syncthreads();
  const double *a = T_SMEM ? st : T;
  const double t0 = n > 0 ? T[0] : 0.0, tmax = n > 0 ? T[n - 1] : 0.0, w = (tmax - t0) / nc;
  const double iw = w > 0.0 ? 1.0 / w : 0.0;
  const int p0 = blockIdx.y * chunk_len, p1 = min(P, p0 + chunk_len);
  // One streaming read of the null (8 S P bytes): U permutations in flight per thread.  The
  // normalisation nes_null = v / mean (:608-610, :619-621) is an exact division in the
  // reference and decides `>` / `<` against the observed NES; here the value is first
  // located with v * (1 / mean) -- one multiply instead of a division per value -- and only
  // when a threshold lies within the error of that product is the division done and the value
  // located again (a null NES that ties with an observed one takes that path: counts exact).
  constexpr int U = 8;
  unsigned long long gpos = 0, gneg = 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    const double mp = mpos[s], mg = mneg[s], e = es[s];
    const double dpos = mp, dneg = abs_flag ? mp : -mg;  // the divisor of a value by its sign
    const double rpos = 1.0 / dpos, rneg = 1.0 / dneg;
    const double *base = null + (size_t)s + (size_t)S * resp;
    const size_t stride = (size_t)S * R;
    unsigned int lpos = 0, lneg = 0;
    for (int pb = p0; pb < p1; pb += U) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = (pb + u < p1) ? __ldg(base + stride * (pb + u)) : nan("");
#pragma unroll
      for (int u = 0; u < U; u++) {
        const double vs = snap_rel != 0.0 ? snap_to_observed(v[u], e, snap_rel) : v[u];
        const bool vneg = !(vs >= 0.0);
        double nv = vs * (vneg ? rneg : rpos);
        bool sure = isfinite(nv) && fabs(nv) > 1e-290;
        bool nonstrict = !abs_flag && nv < 0.0;   // a negative value undercuts: `<`, i.e. count T <= nv
        int B = 0;
        if (sure) {
          B = lut_count(a, n, lut, nc, t0, w, iw, tmax, nv, nonstrict);
          const double d = FDR_APPROX_REL * fabs(nv);
          sure = (B == 0 || nv - a[B - 1] > d) && (B == n || a[B] - nv > d);
        }
        if (!sure) {  // the reference's own arithmetic
          nv = vs / (vneg ? dneg : dpos);
          if (isnan(nv)) continue;
          nonstrict = !abs_flag && nv < 0.0;
          B = lut_count(a, n, lut, nc, t0, w, iw, tmax, nv, nonstrict);
        }
        lpos += nv > 0.0;                                                          // :503
        lneg += nv < 0.0;                                                          // :504
        // bins that no set's count reads are skipped: they are also the crowded ones
        if (nonstrict ? (B < n2) : (B > n2)) atomicAdd(&h[nonstrict ? B : B + 1], 1u);
      }
    }
    gpos += lpos; gneg += lneg;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n + 2; i += blockDim.x) if (h[i]) atomicAdd(&hist[i], (unsigned long long)h[i]);
  for (int o = 16; o; o >>= 1) {
    gpos += __shfl_xor_sync(0xffffffffu, gpos, o);
    gneg += __shfl_xor_sync(0xffffffffu, gneg, o);
  }
  if ((threadIdx.x & 31) == 0) { if (gpos) atomicAdd(&glob[0], gpos); if (gneg) atomicAdd(&glob[1], gneg); }
}

// histogram -> per-set exceedance counts: a running sum over the sorted
// positions (descending for group 1, ascending for group 2); single CTA, every
// thread owns a contiguous chunk, chunk totals scanned in shared memory
__device__ void running_sum_scatter(const unsigned long long *__restrict__ hist, int first, int step,
                                    const int *__restrict__ id, int id_first, int n,
                                    long long *__restrict__ out, unsigned long long *__restrict__ part) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (n + nt - 1) / nt, t0 = min(n, tid * per), t1 = min(n, t0 + per);
  unsigned long long sum = 0;
  for (int t = t0; t < t1; t++) sum += hist[first + step * t];
  __syncthreads();
  part[tid] = sum;
  __syncthreads();
  unsigned long long run = 0;
  for (int u = 0; u < tid; u++) run += part[u];
  for (int t = t0; t < t1; t++) {
    run += hist[first + step * t];
    out[id[id_first + step * t]] = (long long)run;
  }
}
__global__ void __launch_bounds__(256)
hist_to_counts_kernel(const unsigned long long *__restrict__ hist1, const int *id1, int n1,
                      const unsigned long long *__restrict__ hist2, const int *id2, int n2, int S,
                      long long *__restrict__ null_counts) {
  __shared__ unsigned long long part[256];
  // group 1: position j = n1-1 .. 0 accumulates hist1[j + 1]
  running_sum_scatter(hist1, n1, -1, id1, n1 - 1, n1, null_counts, part);
  // group 2: position j = 0 .. n2-1 accumulates hist2[j]
  running_sum_scatter(hist2, 0, 1, id2, 0, n2, null_counts, part);
}

// ---- finish: raw FDR + monotone adjustment ----------------------------------
// a_i = #{nes > nes_i} (positive side, :519) or #{nes < nes_i} (:523) over all sets,
// and the sign counts cp = #{nes > 0}, cn = #{nes < 0} (:501-502): S^2 compares, tiled
// like group_rank_kernel.  acount[] and sign_n[0..1] start at zero.
__global__ void __launch_bounds__(RANK_TI)
fdr_exceed_kernel(const double *__restrict__ nes, int S, int abs_flag, int chunk,
                  int *__restrict__ acount, int *__restrict__ sign_n) {
  __shared__ double sk[RANK_TJ];
  const int i = blockIdx.x * RANK_TI + threadIdx.x;
  const bool valid = i < S;
  const double v = valid ? nes[i] : 0.0;
  const bool pos_side = (v >= 0.0) || abs_flag;
  const int j0 = blockIdx.y * chunk, j1 = min(S, j0 + chunk);
  int a = 0;
  for (int jt = j0; jt < j1; jt += RANK_TJ) {
    const int len = min(RANK_TJ, j1 - jt);
    __syncthreads();
    for (int t = threadIdx.x; t < len; t += RANK_TI) sk[t] = nes[jt + t];
    __syncthreads();
    for (int t = 0; t < len; t++) {
      const double w = sk[t];
      a += pos_side ? (w > v) : (w < v);
    }
  }
  if (valid && a) atomicAdd(&acount[i], a);
  if (valid && blockIdx.y == 0) {
    if (v > 0.0) atomicAdd(&sign_n[0], 1);
    if (v < 0.0) atomicAdd(&sign_n[1], 1);
  }
}
__global__ void fdr_raw_kernel(const double *__restrict__ nes, const long long *__restrict__ null_counts,
                               const long long *__restrict__ glob, const int *__restrict__ acount,
                               const int *__restrict__ sign_n, int S, int abs_flag,
                               double *__restrict__ fdr, long long *__restrict__ fdr_counts,
                               long long *__restrict__ glob_counts, int *__restrict__ adj_group) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  const double v = nes[i];
  const long long cp = sign_n[0], cn = sign_n[1], a = acount[i];
  const bool pos_side = (v >= 0.0) || abs_flag;
  const long long b = null_counts[i];
  double rel, rel_null;
  if (pos_side) { rel = (double)(a + 1) / (double)cp; rel_null = (double)(b + 1) / (double)glob[0]; }
  else { rel = (double)(a + 1) / (double)cn; rel_null = (double)(b + 1) / (double)glob[1]; }
  if (isnan(v)) rel = nan("");
  fdr[i] = rel_null / rel;                            // :527
  if (fdr_counts) { fdr_counts[2 * i] = a; fdr_counts[2 * i + 1] = b; }
  if (i == 0 && glob_counts) { glob_counts[0] = cp; glob_counts[1] = cn; glob_counts[2] = glob[0]; glob_counts[3] = glob[1]; }
  adj_group[i] = isnan(v) ? 0 : (v >= 0.0 ? 1 : 2);   // adj_fdr_nes splits on nes >= 0 only (:477, :488)
}
__global__ void negate_group2_kernel(const double *__restrict__ nes, const int *__restrict__ group,
                                     int S, double *__restrict__ key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  // group 1: increasing NES (:478); group 2: decreasing NES (:489) == increasing -NES
  key[i] = group[i] == 2 ? -nes[i] : nes[i];
}
// inclusive prefix-min in the given order, start value `init`; single CTA.
// vals[id[k]] is visited for k = 0..n-1 and overwritten with the running min.
__global__ void __launch_bounds__(1024)
ordered_prefix_min_kernel(double *__restrict__ vals, const int *__restrict__ group,
                          const int *__restrict__ pos, int S, int which, int n, double init,
                          int *__restrict__ order_scratch) {
  __shared__ double cmin[1024];
  const int tid = threadIdx.x;
  for (int i = tid; i < S; i += blockDim.x) if (group[i] == which) order_scratch[pos[i]] = i;
  __syncthreads();
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int k0 = tid * per, k1 = min(n, k0 + per);
  double m = INFINITY;
  for (int k = k0; k < k1; k++) m = fmin(m, vals[order_scratch[k]]);
  cmin[tid] = m;
  __syncthreads();
  double run = init;
  for (int t = 0; t < tid; t++) run = fmin(run, cmin[t]);
  for (int k = k0; k < k1; k++) {
    const int i = order_scratch[k];
    const double v = vals[i];
    if (v <= run) run = v; else vals[i] = run;  // :479-483 (NaN never passes '<=')
  }
}

// BH (p.adjust(p, 'BH'), R/flexgsea.R:625): o = order(p, decreasing=TRUE);
// pmin(1, cummin(n / (n:1) * p[o]))[order(o)]
__global__ void bh_prepare_kernel(const double *__restrict__ p, int n, double *__restrict__ key,
                                  int *__restrict__ group) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  key[i] = -p[i];  // decreasing p == increasing -p, ties by index
  group[i] = 1;
}
__global__ void bh_scale_kernel(const double *__restrict__ p, const int *__restrict__ pos, int n,
                                double *__restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  vals[i] = ((double)n / (double)(n - pos[i])) * p[i];
}
__global__ void clamp1_kernel(double *__restrict__ v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = fmin(v[i], 1.0);
}

int launch_sig_means(fgs_ctx *c, const double *d_sums_parts, int nparts, size_t stride, const long long *d_counts,
                     const double *d_es, int S, long long P_total, int split_p, int calc_nes,
                     int abs_flag, double *d_p, double *d_ph, double *d_pl, double *d_mpos,
                     double *d_mneg, double *d_nes, double *d_fwer) {
  sig_means_kernel<<<ceil_div(S, 128), 128, 0, c->stream>>>(d_sums_parts, nparts, stride, d_counts, d_es, S,
                                                            P_total, split_p, calc_nes, abs_flag,
                                                            d_p, d_ph, d_pl, d_mpos, d_mneg, d_nes,
                                                            d_fwer);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

// work layout (ints / doubles / u64) carved by the caller
int launch_sig_stage2(fgs_ctx *c, const double *d_null, int S, int R, int P, int response,
                      const double *d_es, int abs_flag, const double *d_mpos, const double *d_mneg,
                      double *d_nes, long long *d_null_counts, long long *d_glob) {
  // scratch: group[S], pos[S], gn[4], id1[S], id2[S] (ints); t1[S], t2[S] (doubles);
  // hist1[S+1], hist2[S+1] (u64)
  FGS_TRY(c->counters.ensure((size_t)4 * S + 8));
  int *group = c->counters.p, *pos = group + S, *id1 = pos + S, *id2 = id1 + S, *gn = id2 + S;
  FGS_TRY(c->sig_d.ensure((size_t)2 * S + 2));
  double *t1 = c->sig_d.p, *t2 = t1 + S;
  FGS_TRY(c->sig_i.ensure((size_t)2 * S + 4));
  unsigned long long *h1 = (unsigned long long *)c->sig_i.p, *h2 = h1 + S + 1;
  FGS_CUDA(cudaMemsetAsync(gn, 0, 4 * sizeof(int), c->stream));
  FGS_CUDA(cudaMemsetAsync(h1, 0, (size_t)(2 * S + 2) * sizeof(unsigned long long), c->stream));
  FGS_CUDA(cudaMemsetAsync(d_glob, 0, 2 * sizeof(long long), c->stream));
  FGS_CUDA(cudaMemsetAsync(d_null_counts, 0, (size_t)S * sizeof(long long), c->stream));
  const int nb = ceil_div(S, 128);
  fdr_groups_kernel<<<nb, 128, 0, c->stream>>>(d_nes, S, abs_flag, group);
  launch_group_rank(c, d_nes, group, S, pos, gn);
  scatter_sorted_kernel<<<nb, 128, 0, c->stream>>>(d_nes, group, pos, S, t1, id1, t2, id2);
  c->launches += 3;
  FGS_CUDA(cudaGetLastError());
  int h_gn[4];
  FGS_CUDA(cudaMemcpyAsync(h_gn, gn, sizeof(h_gn), cudaMemcpyDeviceToHost, c->stream));
  FGS_CUDA(cudaStreamSynchronize(c->stream));
  const int n1 = h_gn[1], n2 = h_gn[2];
  const int n12 = n1 + n2;
  // T = [group-2 thresholds][group-1 thresholds], one ascending list (see fdr_count_kernel)
  FGS_TRY(c->sig_t.ensure((size_t)std::max(n12, 1)));
  if (n2 > 0) FGS_CUDA(cudaMemcpyAsync(c->sig_t.p, t2, (size_t)n2 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (n1 > 0) FGS_CUDA(cudaMemcpyAsync(c->sig_t.p + n2, t1, (size_t)n1 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  unsigned long long *H = h1;   // [n12 + 2]: hist2 = H[0 .. n2], hist1 = H[n2 + 1 .. n12 + 1]
  if (P > 0) {
    if (n12 >= 65536) { set_error("pooled FDR over %d gene sets: more than the 65535 the cell index addresses", S); return FGS_ERR_UNSUPPORTED; }
    // the finest cell index that still fits on chip next to the histogram (and the thresholds, when they fit too)
    const size_t smem_hist = (size_t)(n12 + 2) * 4 + 16, smem_thr = (size_t)n12 * 8;
    int nc = FDR_NC_MAX;
    while (nc > FDR_NC_MIN && smem_hist + smem_thr + (size_t)(nc + 1) * 2 > (size_t)c->smem_optin) nc >>= 1;
    const size_t smem_lut = (size_t)(nc + 1) * 2;
    const size_t smem_h = smem_hist + smem_lut;                  // histogram + cell index
    const size_t smem_t = smem_thr + smem_h;                     // + the thresholds themselves
    FGS_TRY(c->sig_lut.ensure((size_t)(FDR_NC_MAX + 2) / 2 + 1));
    uint16_t *d_lut = (uint16_t *)c->sig_lut.p;
    fdr_lut_kernel<<<ceil_div(nc + 1, 256), 256, 0, c->stream>>>(c->sig_t.p, n12, nc, d_lut);
    c->launches += 1;
    if (smem_h > (size_t)c->smem_optin) {
      // (about 50,000 gene sets: the per-CTA histogram no longer fits on chip)
      set_error("pooled FDR over %d gene sets needs %zu bytes of shared memory per CTA, %d available", S, smem_h,
                c->smem_optin);
      return FGS_ERR_UNSUPPORTED;
    }
    const bool t_smem = smem_t <= (size_t)c->smem_optin;
    const int threads = t_smem ? FDR_T_SMEM : FDR_T_GLOBAL;
    const int per_sm = t_smem ? 1 : std::max(1, std::min(8, (int)((size_t)c->smem_optin / smem_h)));
    int gx = ceil_div(S, threads);
    int gy = std::max(1, (c->num_sms * per_sm) / gx);    // one wave of CTAs, each streaming its chunk
    if (gy > P) gy = P;
    int chunk_len = (P + gy - 1) / gy;
    gy = (P + chunk_len - 1) / chunk_len;
    dim3 grid(gx, gy);
    const double snap = c->null_exact_near_obs ? 0.0 : SIG_SNAP_REL;
    if (t_smem) {
      auto k = fdr_count_kernel<true>;
      FGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      k<<<grid, threads, smem_t, c->stream>>>(d_null, S, R, P, response, abs_flag, d_es, d_mpos, d_mneg, c->sig_t.p,
                                              n12, n2, d_lut, nc, chunk_len, snap, H, (unsigned long long *)d_glob);
    } else {
      auto k = fdr_count_kernel<false>;
      FGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h));
      k<<<grid, threads, smem_h, c->stream>>>(d_null, S, R, P, response, abs_flag, d_es, d_mpos, d_mneg, c->sig_t.p,
                                              n12, n2, d_lut, nc, chunk_len, snap, H, (unsigned long long *)d_glob);
    }
    c->launches++;
    FGS_CUDA(cudaGetLastError());
  }
  h2 = H; h1 = H + n2 + 1;
  hist_to_counts_kernel<<<1, 256, 0, c->stream>>>(h1, id1, n1, h2, id2, n2, S, d_null_counts);
  c->launches++;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

int launch_sig_finish(fgs_ctx *c, const double *d_nes, const long long *d_null_counts,
                      const long long *d_glob, int S, int abs_flag, double *d_fdr,
                      long long *d_fdr_counts, long long *d_glob_counts) {
  FGS_TRY(c->counters.ensure((size_t)4 * S + 8));
  int *group = c->counters.p, *pos = group + S, *order = pos + S, *gn = order + S + S;
  FGS_TRY(c->sig_d.ensure((size_t)2 * S + 2));
  double *key = c->sig_d.p;
  const int nb = ceil_div(S, 128);
  int *acount = order + S;  // the second half of the order scratch
  FGS_CUDA(cudaMemsetAsync(gn, 0, 8 * sizeof(int), c->stream));
  FGS_CUDA(cudaMemsetAsync(acount, 0, (size_t)S * sizeof(int), c->stream));
  {
    int gy = ceil_div(S, RANK_TJ);
    if (gy > 32) gy = 32;
    const int chunk = ceil_div(ceil_div(S, gy), RANK_TJ) * RANK_TJ;
    gy = ceil_div(S, chunk);
    dim3 grid(ceil_div(S, RANK_TI), gy);
    fdr_exceed_kernel<<<grid, RANK_TI, 0, c->stream>>>(d_nes, S, abs_flag, chunk, acount, gn + 4);
  }
  fdr_raw_kernel<<<nb, 128, 0, c->stream>>>(d_nes, d_null_counts, d_glob, acount, gn + 4, S, abs_flag, d_fdr,
                                            d_fdr_counts, d_glob_counts, group);
  negate_group2_kernel<<<nb, 128, 0, c->stream>>>(d_nes, group, S, key);
  launch_group_rank(c, key, group, S, pos, gn);
  c->launches += 4;
  FGS_CUDA(cudaGetLastError());
  int h_gn[4];
  FGS_CUDA(cudaMemcpyAsync(h_gn, gn, sizeof(h_gn), cudaMemcpyDeviceToHost, c->stream));
  FGS_CUDA(cudaStreamSynchronize(c->stream));
  for (int which = 1; which <= 2; which++) {
    if (h_gn[which] <= 0) continue;
    ordered_prefix_min_kernel<<<1, 1024, 0, c->stream>>>(d_fdr, group, pos, S, which, h_gn[which],
                                                         1.0, order);
    c->launches++;
  }
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

int launch_bh(fgs_ctx *c, const double *d_p, int S, double *d_fdr) {
  FGS_TRY(c->counters.ensure((size_t)4 * S + 8));
  int *group = c->counters.p, *pos = group + S, *order = pos + S, *gn = order + S + S;
  FGS_TRY(c->sig_d.ensure((size_t)2 * S + 2));
  double *key = c->sig_d.p;
  const int nb = ceil_div(S, 128);
  FGS_CUDA(cudaMemsetAsync(gn, 0, 8 * sizeof(int), c->stream));
  bh_prepare_kernel<<<nb, 128, 0, c->stream>>>(d_p, S, key, group);
  launch_group_rank(c, key, group, S, pos, gn);
  bh_scale_kernel<<<nb, 128, 0, c->stream>>>(d_p, pos, S, d_fdr);
  ordered_prefix_min_kernel<<<1, 1024, 0, c->stream>>>(d_fdr, group, pos, S, 1, S, INFINITY, order);
  clamp1_kernel<<<nb, 128, 0, c->stream>>>(d_fdr, S);
  c->launches += 5;
  FGS_CUDA(cudaGetLastError());
  return FGS_OK;
}

}  // namespace fgs
