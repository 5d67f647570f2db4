// Tie-averaged pooled ranks -> normal scores, for the raw draws and for the folded
// draws, written back to the draws' original [C][N] positions.
//
//   k_param_prep      pooled median (statistics.median, diagnostics.py:97), the split
//                     point of the folded order, numpy-linear quantiles
//                     (backends_numpy.py:44) and their order statistics.
//   k_rank_codes      (LSD fallback path only) one tile of the sorted keys: tie runs
//                     by equality (diagnostics.py:113-122) -> k = i + j per sorted element
//                     (all marked tied).  The distribution-sort path gets k from k_finish directly.
//   k_merge_partition merge-path split points of the folded order.
//   k_fold            folded values |v - med| (diagnostics.py:98) are monotone on each
//                     side of the median, so their ascending order is the MERGE of the
//                     reversed left part and the right part of the already sorted
//                     sequence -- no second sort.  Tie runs of the folded order, then ONE
//                     16-byte scatter per draw of {z_bulk, z_folded} to its original
//                     position (a scattered store costs a full L1 wavefront whatever its
//                     width, so both scores travel together).
#pragma once
#include "as241.cuh"
#include "common.cuh"

namespace b200ref {

#ifndef B200REF_RANK_NT
#define B200REF_RANK_NT 128
#endif
constexpr int RANK_NT = B200REF_RANK_NT;
constexpr int RANK_IPT = 8;
constexpr int RANK_TILE = RANK_NT * RANK_IPT;

struct RankBufs {
  const uint64_t* key[2];  // sorted keys are in key[sel[p]]
  uint32_t* idx[2];
  const int* sel;
  uint32_t* kcode;     // [PB][M] k = i + j of the TIED sorted elements (flagged by TIED_BIT in idx)
  const uint8_t* fallback;  // [PB] 1 = parameter went through the LSD path
  const int* nfb;           // number of such parameters
  const int* plist;         // [*nfb] their indices
  double* med;         // [PB]
  uint32_t* m0;        // [PB] number of draws strictly below the median
  uint32_t* apos;      // [PB][ntiles+1] merge-path split (count taken from the left part)
  const double* ztab;  // [2M]; null when no scores are wanted (stats only)
  const double* zodd;  // [M]  zodd[t] = ztab[2t + 1]
  const int* flag;     // bit0: non-finite input seen -> the ordering is meaningless, kernels return
};

__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t* a, uint32_t lo, uint32_t hi, uint64_t v) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ uint32_t upper_bound_u64(const uint64_t* a, uint32_t lo, uint32_t hi, uint64_t v) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// numpy's linear quantile on a sorted sequence: virtual index (n-1)*q,
// _lerp(a, b, g) = a + (b-a)*g if g < 0.5 else b - (b-a)*(1-g); no FMA contraction.
__device__ __forceinline__ double quantile_lerp(const uint64_t* keys, uint32_t M, double q, double* os) {
  const double vi = __dmul_rn((double)(M - 1), q);
  double fl = floor(vi);
  uint32_t lo = (uint32_t)fl;
  if (lo > M - 1) lo = M - 1;
  const uint32_t hi = (lo + 1 < M) ? lo + 1 : M - 1;
  const double g = __dsub_rn(vi, fl);
  const double a = key2dbl(keys[lo]);
  const double b = key2dbl(keys[hi]);
  if (os) { os[0] = a; os[1] = b; }
  const double diff = __dsub_rn(b, a);
  return (g < 0.5) ? __dadd_rn(a, __dmul_rn(diff, g)) : __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, g)));
}

struct PrepOut {
  double* quant;   // [P][nq] or null
  double* ostats;  // [P][nq][2] or null
  double* median;  // [P] or null
};

struct QuantArgs {
  double q[8];
  int nq;
};

__global__ void k_param_prep(RankBufs b, uint32_t M, int PB, QuantArgs qa, PrepOut o) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= PB || (*b.flag & 1)) return;
  const uint64_t* keys = b.key[b.sel[p]] + (size_t)p * M;
  double med;
  if (M & 1u) med = key2dbl(keys[M / 2]);
  else med = __dmul_rn(__dadd_rn(key2dbl(keys[M / 2 - 1]), key2dbl(keys[M / 2])), 0.5);
  b.med[p] = med;
  b.m0[p] = lower_bound_u64(keys, 0, M, dbl2key(med));
  if (o.median) o.median[p] = med;
  for (int i = 0; i < qa.nq; ++i) {
    double os[2];
    const double v = quantile_lerp(keys, M, qa.q[i], os);
    if (o.quant) o.quant[(size_t)p * qa.nq + i] = v;
    if (o.ostats) {
      o.ostats[((size_t)p * qa.nq + i) * 2 + 0] = os[0];
      o.ostats[((size_t)p * qa.nq + i) * 2 + 1] = os[1];
    }
  }
}

// padded index for blocked (thread-major) access to 8-byte shared arrays: without it
// thread t's 8 consecutive elements start 64 B apart -> 16-way bank conflicts.
__device__ __forceinline__ uint32_t pidx8(uint32_t e) { return e + (e >> 3); }
constexpr int RANK_SV = RANK_TILE + 2 + (RANK_TILE + 2) / 8 + 2;

// sv[pidx8(1 + e)] = element e of the tile (ascending), sv[0] / sv[pidx8(n + 1)] = neighbours outside
// the tile (valid iff has_prev / has_next).  Returns k = i + j per item, where
// [i, j) is the element's tie run in the whole sequence.  ext_i / ext_j are the run
// bounds of runs that cross the tile's first / last element (only read if needed).
template <int NT, int IPT>
__device__ __forceinline__ void tie_codes(const uint64_t* sv, uint32_t n, uint32_t gbase, bool has_prev,
                                          bool has_next, const uint32_t* ext, uint32_t* scan_sm,
                                          uint32_t (&kout)[IPT]) {
  uint32_t hi[IPT], tj[IPT];
#pragma unroll
  for (int it = 0; it < IPT; ++it) {
    const uint32_t e = threadIdx.x * IPT + it;
    hi[it] = 0;
    tj[it] = 0xffffffffu;
    if (e < n) {
      const uint64_t cur = sv[pidx8(e + 1)];
      const bool head = (e == 0) ? !(has_prev && sv[0] == cur) : (sv[pidx8(e)] != cur);
      const bool tail = (e == n - 1) ? !(has_next && sv[pidx8(n + 1)] == cur) : (sv[pidx8(e + 2)] != cur);
      if (head) hi[it] = gbase + e + 1;
      if (tail) tj[it] = gbase + e + 1;
    }
  }
  block_scan_max_fwd<NT, IPT>(hi, scan_sm);
  block_scan_min_rev<NT, IPT>(tj, scan_sm);
#pragma unroll
  for (int it = 0; it < IPT; ++it) {
    const uint32_t i = hi[it] ? hi[it] - 1 : ext[0];
    const uint32_t j = (tj[it] != 0xffffffffu) ? tj[it] : ext[1];
    kout[it] = i + j;
  }
}

// LSD fallback path: k = i + j for every sorted element of the flagged parameters.
__global__ void __launch_bounds__(RANK_NT) k_rank_codes(RankBufs b, uint32_t M, double* rank_out) {
  __shared__ uint64_t sv[RANK_SV];
  __shared__ uint32_t ext[2];
  __shared__ uint32_t scan_sm[RANK_NT / 32];
  const int nlist = *b.nfb;
  const uint32_t tile = blockIdx.x;
  const uint32_t gbase = tile * RANK_TILE;
  const uint32_t n = min((uint32_t)RANK_TILE, M - gbase);
  const bool has_prev = gbase > 0, has_next = gbase + n < M;
  for (int q = blockIdx.y; q < nlist; q += gridDim.y) {
  const int p = b.plist[q];
  const int s = b.sel[p];
  const uint64_t* keys = b.key[s] + (size_t)p * M;
  uint32_t* idx = b.idx[s] + (size_t)p * M;
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < n; e += RANK_NT) sv[pidx8(e + 1)] = keys[gbase + e];
  if (threadIdx.x == 0) sv[0] = has_prev ? keys[gbase - 1] : 0;
  if (threadIdx.x == 32) sv[pidx8(n + 1)] = has_next ? keys[gbase + n] : 0;
  __syncthreads();
  if (threadIdx.x == 0)
    ext[0] = (has_prev && sv[0] == sv[pidx8(1)]) ? lower_bound_u64(keys, 0, gbase, sv[pidx8(1)]) : 0;
  if (threadIdx.x == 32)
    ext[1] = (has_next && sv[pidx8(n + 1)] == sv[pidx8(n)]) ? upper_bound_u64(keys, gbase + n, M, sv[pidx8(n)]) : 0;
  __syncthreads();

  uint32_t k[RANK_IPT];
  tie_codes<RANK_NT, RANK_IPT>(sv, n, gbase, has_prev, has_next, ext, scan_sm, k);
  // every element of a fallback parameter is marked tied: k_fold then reads its code from kcode
  uint32_t* kc = b.kcode + (size_t)p * M;
#pragma unroll
  for (int it = 0; it < RANK_IPT; ++it) {
    const uint32_t e = threadIdx.x * RANK_IPT + it;
    if (e < n) {
      kc[gbase + e] = k[it];
      const uint32_t pos = idx[gbase + e] & ~TIED_BIT;
      idx[gbase + e] = pos | TIED_BIT;
      if (rank_out) rank_out[(size_t)p * M + pos] = (double)(k[it] + 1u) * 0.5;
    }
  }
  }
}

// Folded views of the sorted sequence s[0..M): left part reversed and right part,
// both ascending in |v - med|.
struct FoldView {
  const uint64_t* keys;
  uint32_t m0, M;
  double med;
  __device__ __forceinline__ uint64_t fbits(uint32_t t) const {
    return (uint64_t)__double_as_longlong(fabs(__dsub_rn(key2dbl(keys[t]), med)));
  }
  __device__ __forceinline__ uint32_t nL() const { return m0; }
  __device__ __forceinline__ uint32_t nR() const { return M - m0; }
  __device__ __forceinline__ uint64_t L(uint32_t a) const { return fbits(m0 - 1 - a); }
  __device__ __forceinline__ uint64_t R(uint32_t c) const { return fbits(m0 + c); }
  // merge-path: number of left-part elements among the first o outputs (left first on ties)
  __device__ uint32_t split(uint32_t o) const {
    uint32_t lo = (o > nR()) ? o - nR() : 0, hi = min(o, nL());
    // The predicate "L(a) <= R(o - 1 - a)" is true below the split and false from it on.  The two parts hold about
    // the same share of every distance range, so the split lies near o * nL / M: bracket it from that guess with
    // doubling steps before bisecting (6-8 dependent loads instead of ~18; same result, the predicate is monotone).
    if (lo < hi) {
      const uint32_t g = min(hi, max(lo, (uint32_t)(((uint64_t)o * nL()) / max(M, 1u))));
      auto below = [&](uint32_t a) { return L(a) <= R(o - 1 - a); };  // a in [lo, hi)
      if (g < hi && below(g)) {
        uint32_t step = 16, a = g;  // the split is above g
        lo = g + 1;
        while (lo < hi) {
          a = min(hi - 1, a + step);
          if (below(a)) { lo = a + 1; step <<= 1; } else { hi = a; break; }
          if (a == hi - 1) break;
        }
      } else {
        uint32_t step = 16, a = g;  // the split is at or below g
        hi = min(hi, g);
        while (lo < hi) {
          a = (a > lo + step) ? a - step : lo;
          if (below(a)) { lo = a + 1; break; } else { hi = a; step <<= 1; }
          if (a == lo) break;
        }
      }
    }
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (L(mid) <= R(o - 1 - mid)) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
  __device__ uint32_t count_less(uint64_t d) const {  // elements of both parts with value < d
    uint32_t lo = 0, hi = nL();
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (L(mid) < d) lo = mid + 1; else hi = mid; }
    uint32_t c = lo;
    lo = 0; hi = nR();
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (R(mid) < d) lo = mid + 1; else hi = mid; }
    return c + lo;
  }
  __device__ uint32_t count_leq(uint64_t d) const {
    uint32_t lo = 0, hi = nL();
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (L(mid) <= d) lo = mid + 1; else hi = mid; }
    uint32_t c = lo;
    lo = 0; hi = nR();
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (R(mid) <= d) lo = mid + 1; else hi = mid; }
    return c + lo;
  }
};

__global__ void k_merge_partition(RankBufs b, uint32_t M, uint32_t ntiles, int PB) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (q > ntiles || p >= PB || (*b.flag & 1)) return;
  FoldView f{b.key[b.sel[p]] + (size_t)p * M, b.m0[p], M, b.med[p]};
  const uint32_t o = min(q * (uint32_t)RANK_TILE, M);
  b.apos[(size_t)p * (ntiles + 1) + q] = f.split(o);
}

struct FoldOut {
  double2* zz;         // [PB][M] {z_bulk, z_folded} in original positions (diagnostics path) or null
  double* z;           // separate outputs (ranks API), any may be null
  double* zf;
  double* rankf;
  // Long chains (a parameter's zz row does not fit L2): the scatter goes through per-destination-window record
  // segments (k_fold appends {position, z_bulk, z_folded}, k_unbin scatters window by window).  bin_shift = 0: direct.
  int bin_shift;       // window = positions >> bin_shift (at most 32 windows per parameter)
  uint32_t* cursor;    // [PB][32] records appended to each window's segment so far
  uint32_t* rpos;      // [PB][M] records: original position   (segment of window r starts at r << bin_shift)
  double* rzb;         // [PB][M]          z_bulk
  double* rzf;         // [PB][M]          z_folded
};

// Index into the folded-bits window: one pad word per 16 keeps the merge's reads (neighbouring threads
// walk the windows ~4 elements apart) off the same shared-memory banks, and a half-warp's 16 consecutive
// slots of the striped stores stay contiguous.
#ifndef B200REF_FOLD_PAD_SH
#define B200REF_FOLD_PAD_SH 4
#endif
constexpr int FOLD_PAD_SH = B200REF_FOLD_PAD_SH;
__device__ __forceinline__ uint32_t sdi(uint32_t i) { return i + (i >> FOLD_PAD_SH); }

#ifndef B200REF_FOLD_OCC
#define B200REF_FOLD_OCC 8
#endif

// first window slot in [lo, hi) whose folded bits are >= d (STRICT = false) or > d (STRICT = true)
template <bool STRICT>
__device__ __forceinline__ uint32_t window_bound(const uint64_t* sd, uint32_t lo, uint32_t hi, uint64_t d) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    const uint64_t v = sd[sdi(mid)];
    if (STRICT ? (v <= d) : (v < d)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One CTA = RANK_TILE consecutive outputs of the folded (merged) order.
//  phase 1  every global read of the tile is issued up front and coalesced: the left / right
//           windows of the sorted sequence (keys -> folded bits in shared memory; original positions
//           and bulk scores stay in registers), and the tile of untied folded scores;
//  phase 2  per-thread merge path + serial merge in shared memory: every window slot learns its
//           merged position; ties are detected on the fly (continuous draws: almost never);
//  phase 3  per window slot: the folded score of its merged position and ONE 16-byte scatter per draw.
//           Tiles with ties get their tie runs by binary searches in the two windows.
__global__ void __launch_bounds__(RANK_NT, B200REF_FOLD_OCC) k_fold(RankBufs b, uint32_t M, uint32_t ntiles, FoldOut o) {
  __shared__ uint64_t sd[RANK_TILE + (RANK_TILE >> FOLD_PAD_SH) + 2];  // left window then right window (folded bits, index sdi())
  __shared__ double szf[RANK_TILE];       // folded scores of the tile's merged positions (no ties)
  __shared__ uint16_t dest[RANK_TILE];    // window slot -> merged position inside the tile
  __shared__ uint64_t edge[2];            // merged values just before / after the tile
  __shared__ uint32_t ext[2];             // tie-run bounds reaching outside the tile
  const int p = blockIdx.y;
  const uint32_t tile = blockIdx.x;
  if (*b.flag & 1) return;
  const int s = b.sel[p];
  const uint32_t* idx = b.idx[s] + (size_t)p * M;
  const uint32_t* kc = b.kcode + (size_t)p * M;
  FoldView f{b.key[s] + (size_t)p * M, b.m0[p], M, b.med[p]};
  const uint32_t gbase = tile * RANK_TILE;
  const uint32_t n = min((uint32_t)RANK_TILE, M - gbase);
  const uint32_t a0 = b.apos[(size_t)p * (ntiles + 1) + tile];
  const uint32_t a1 = b.apos[(size_t)p * (ntiles + 1) + tile + 1];
  const uint32_t c0 = gbase - a0, c1 = gbase + n - a1;
  const uint32_t nl = a1 - a0, nr = c1 - c0;
  const bool has_prev = gbase > 0, has_next = gbase + n < M;
  const int tid = threadIdx.x;

  // loads are issued in two batches of RANK_IPT / 2 (clamped indices, no branches) so that every
  // warp has 12 independent requests in flight instead of one dependent chain per element
  uint32_t pi[RANK_IPT];
  double zb[RANK_IPT];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    constexpr int HB = RANK_IPT / 2;
    uint32_t tt[HB];
    uint64_t kv[HB];
#pragma unroll
    for (int j = 0; j < HB; ++j) {
      const uint32_t e = min((h * HB + j) * RANK_NT + tid, n - 1);
      tt[j] = (e < nl) ? (f.m0 - 1 - (a0 + e)) : (f.m0 + c0 + (e - nl));
    }
    // bulk score of an untied draw at sorted position t: T[2t + 1] = zodd[t] (read along the window);
    // tied draws (flagged by k_finish) fetch their run's code instead
#pragma unroll
    for (int j = 0; j < HB; ++j) {
      kv[j] = f.keys[tt[j]];
      pi[h * HB + j] = idx[tt[j]];
      zb[h * HB + j] = b.zodd[tt[j]];
    }
#pragma unroll
    for (int j = 0; j < HB; ++j) {
      if (pi[h * HB + j] & TIED_BIT) {
        zb[h * HB + j] = b.ztab[kc[tt[j]]];
        pi[h * HB + j] &= ~TIED_BIT;
      }
    }
    // (slots past the end of a short last tile repeat element n - 1: identical values, same address)
#pragma unroll
    for (int j = 0; j < HB; ++j) {
      const uint32_t e = min((h * HB + j) * RANK_NT + tid, n - 1);
      sd[sdi(e)] = (uint64_t)__double_as_longlong(fabs(__dsub_rn(key2dbl(kv[j]), f.med)));
    }
  }
  // scores of the tile's outputs if the folded order has no ties here (the usual case)
#pragma unroll
  for (int it = 0; it < RANK_IPT; ++it) {
    const uint32_t e = it * RANK_NT + tid;
    if (e < n) szf[e] = b.zodd[gbase + e];
  }
  if (tid == 0) {
    // merged element just before the tile: the larger of the two preceding heads
    uint64_t pv = 0;
    if (a0 > 0) pv = f.L(a0 - 1);
    if (c0 > 0) pv = max(pv, f.R(c0 - 1));
    edge[0] = pv;
  }
  if (tid == 32) {
    uint64_t nv = ~0ull;
    if (a1 < f.nL()) nv = f.L(a1);
    if (c1 < f.nR()) nv = min(nv, f.R(c1));
    edge[1] = nv;
  }
  __syncthreads();

  // per-thread merge path inside the tile; ties are detected by comparing each output with its predecessor
  bool tie = false;
  {
    const uint32_t od = min((uint32_t)tid * RANK_IPT, n);
    uint32_t lo = (od > nr) ? od - nr : 0, hi = min(od, nl);
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (sd[sdi(mid)] <= sd[sdi(nl + od - 1 - mid)]) lo = mid + 1; else hi = mid;
    }
    uint32_t la = lo, lc = od - lo;
    uint64_t vl = (la < nl) ? sd[sdi(la)] : ~0ull, vr = (lc < nr) ? sd[sdi(nl + lc)] : ~0ull;
    // merged value just before this thread's first output
    uint64_t prev = 0;
    bool have_prev = has_prev;
    if (od == 0) {
      prev = edge[0];
    } else {
      if (la > 0) prev = sd[sdi(la - 1)];
      if (lc > 0) prev = max(prev, sd[sdi(nl + lc - 1)]);
      have_prev = true;
    }
#pragma unroll
    for (int it = 0; it < RANK_IPT; ++it) {
      const uint32_t e = od + it;
      if (e < n) {
        const bool takeL = (la < nl) && (lc >= nr || vl <= vr);
        const uint64_t v = takeL ? vl : vr;
        tie |= have_prev && (v == prev);
        prev = v;
        have_prev = true;
        dest[takeL ? la : nl + lc] = (uint16_t)e;
        if (takeL) { ++la; vl = (la < nl) ? sd[sdi(la)] : ~0ull; }
        else       { ++lc; vr = (lc < nr) ? sd[sdi(nl + lc)] : ~0ull; }
        if (e == n - 1) tie |= has_next && (edge[1] == v);
      }
    }
  }
  const int any_tie = __syncthreads_or(tie);
  if (any_tie) {
    // smallest / largest folded value of the tile; a run touching them may reach outside the tile
    if (tid == 0) {
      const uint64_t v0 = min(nl ? sd[sdi(0)] : ~0ull, nr ? sd[sdi(nl)] : ~0ull);
      ext[0] = (has_prev && edge[0] == v0) ? f.count_less(v0) : gbase;
    }
    if (tid == 32) {
      const uint64_t v1 = max(nl ? sd[sdi(nl - 1)] : 0ull, nr ? sd[sdi(nl + nr - 1)] : 0ull);
      ext[1] = (has_next && edge[1] == v1) ? f.count_leq(v1) : gbase + n;
    }
    __syncthreads();
  }
  // per window slot (the order the loads came in): folded score of the slot's merged position, then
  // the only random access, the 16-byte scatter
  const size_t pbase = (size_t)p * M;
  // folded score (and code) of window slot w
  auto fold_score = [&](uint32_t w, uint32_t& kk, double& zf) {
    if (!any_tie) {
      const uint32_t e = dest[w];
      kk = 2u * (gbase + e) + 1u;
      zf = szf[e];
    } else {
      // tie run [i, j) of this slot's folded value in the whole folded order
      const uint64_t d = sd[sdi(w)];
      const uint32_t il = window_bound<false>(sd, 0, nl, d) + (window_bound<false>(sd, nl, nl + nr, d) - nl);
      const uint32_t jl = window_bound<true>(sd, 0, nl, d) + (window_bound<true>(sd, nl, nl + nr, d) - nl);
      const uint32_t i = (il == 0) ? ext[0] : gbase + il;
      const uint32_t j = (jl == n) ? ext[1] : gbase + jl;
      kk = i + j;
      zf = b.ztab[kk];
    }
  };
  if (o.bin_shift && o.zz && !b.fallback[p]) {
    // Long chains: a 16-byte store into a 128 MB row is a DRAM read-modify-write of its own (13 of config 4's
    // 18 ms in this kernel).  The tile's records are appended to the segment of their destination window
    // instead -- every window receives exactly as many records as it has positions, so segment r starts at
    // r << bin_shift -- in runs reserved with one global atomic per (tile, window); k_unbin then scatters one
    // window's segment at a time, inside what L2 can merge into full sectors.
    __shared__ uint32_t bcnt[32], bbase[32];
    if (tid < 32) bcnt[tid] = 0;
    __syncthreads();
    uint32_t rr[RANK_IPT];  // window | arrival rank in the tile's share of it << 5
#pragma unroll
    for (int it = 0; it < RANK_IPT; ++it) {
      const uint32_t w = it * RANK_NT + tid;
      if (w < n) {
        const uint32_t r = pi[it] >> o.bin_shift;
        rr[it] = r | (atomicAdd(&bcnt[r], 1u) << 5);
      }
    }
    __syncthreads();
    if (tid < 32 && bcnt[tid]) bbase[tid] = atomicAdd(&o.cursor[(size_t)p * 32 + tid], bcnt[tid]);
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RANK_IPT; ++it) {
      const uint32_t w = it * RANK_NT + tid;
      if (w < n) {
        uint32_t kk;
        double zf;
        fold_score(w, kk, zf);
        const uint32_t r = rr[it] & 31u;
        const size_t slot = pbase + ((size_t)r << o.bin_shift) + bbase[r] + (rr[it] >> 5);
        B200REF_CHECK(slot < pbase + M && bbase[r] + (rr[it] >> 5) < (1u << o.bin_shift) && pi[it] < M,
                      "k_fold: record leaves its window segment");
        o.rpos[slot] = pi[it];
        o.rzb[slot] = zb[it];
        o.rzf[slot] = zf;
      }
    }
    return;
  }
#pragma unroll
  for (int it = 0; it < RANK_IPT; ++it) {
    const uint32_t w = it * RANK_NT + tid;
    if (w < n) {
      uint32_t kk;
      double zf;
      fold_score(w, kk, zf);
      B200REF_CHECK(pi[it] < M, "k_fold: original position outside the parameter");
      if (o.zz) {
        __stcs(&o.zz[pbase + pi[it]], make_double2(zb[it], zf));  // streaming store (.cs): -1.5 % against the default policy
      } else {
        if (o.z) o.z[pbase + pi[it]] = zb[it];
        if (o.zf) o.zf[pbase + pi[it]] = zf;
        if (o.rankf) o.rankf[pbase + pi[it]] = (double)(kk + 1u) * 0.5;
      }
    }
  }
}

// Second half of the long-chain scatter: record s of a parameter (segments in window order) goes to its
// position.  Consecutive CTAs walk consecutive records, i.e. one destination window at a time.
constexpr int UNBIN_NT = 256, UNBIN_IPT = 8;
__global__ void __launch_bounds__(UNBIN_NT) k_unbin(FoldOut o, const uint8_t* __restrict__ fallback, const int* __restrict__ flag,
                                                    uint32_t M) {
  const int p = blockIdx.y;
  if ((*flag & 1) || fallback[p]) return;
  const size_t pbase = (size_t)p * M;
  const uint32_t s0 = blockIdx.x * (UNBIN_NT * UNBIN_IPT) + threadIdx.x;
  uint32_t pos[UNBIN_IPT];
  double zb[UNBIN_IPT], zf[UNBIN_IPT];
#pragma unroll
  for (int j = 0; j < UNBIN_IPT; ++j) {
    const uint32_t s = min(s0 + j * UNBIN_NT, M - 1);
    pos[j] = o.rpos[pbase + s];
    zb[j] = o.rzb[pbase + s];
    zf[j] = o.rzf[pbase + s];
  }
#pragma unroll
  for (int j = 0; j < UNBIN_IPT; ++j)
    if (s0 + j * UNBIN_NT < M) {
      B200REF_CHECK(pos[j] < M, "k_unbin: record position outside the parameter");
      __stcs(&o.zz[pbase + pos[j]], make_double2(zb[j], zf[j]));
    }
}

}  // namespace b200ref
