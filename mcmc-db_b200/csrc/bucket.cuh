// Distribution sort of one parameter's pooled draws: the fast ordering path.
//
// Replaces the 8-pass LSD radix sort for the ordering step of _rank_normalize
// (diagnostics.py:110) / np.quantile (backends_numpy.py:44) with ONE global
// partition pass plus an on-chip finish:
//
//   k_sample     a strided sample (16 per coarse bucket) is sorted on-chip; every
//                16th sample becomes a splitter, so coarse buckets hold ~M/NB1 draws
//                for ANY continuous distribution (heavy tails, multi-modal, ...).
//   k_bhist      reads the draws once: coarse-bucket counts per 4096-draw chunk,
//                finiteness, and shifted power sums for mean / std.
//   k_bscan      exclusive offsets per (chunk, bucket); mean / std; overflow flag.
//   k_partition  reads the draws again, stages each chunk in shared memory in bucket
//                order and writes (key64, position32) runs -- coalesced stores.
//   k_finish     one CTA per coarse bucket (<= 2048 draws, all in shared memory):
//                256 equal-width fine buckets, then each draw counts the smaller /
//                equal keys of its fine bucket: exact rank, exact tie run [i, j),
//                sorted position.  Writes the sorted (key, position, k = i + j).
//
// The bucket function only has to be MONOTONE for correctness (it is: floor of a
// monotone fp expression, and "number of splitters <= x"); its uniformity only
// affects speed.  A parameter whose data defeat it (a coarse bucket over capacity:
// one value repeated > 2048 times, ...) is flagged and sorted by the LSD path instead.
#pragma once
#include "common.cuh"

// Tuning knobs (defaults are the measured optimum on B200, see profiles/README.md):
//   B200REF_BHIST_UNROLL  draws in flight per thread in k_bhist (explicit groups)
//   B200REF_FIN_LOAD2     keys loaded per group before their bucket functions run in k_finish
#ifndef B200REF_BHIST_UNROLL
#define B200REF_BHIST_UNROLL 2
#endif
#ifndef B200REF_FIN_LOAD2
#define B200REF_FIN_LOAD2 2
#endif
#ifndef B200REF_FIN_TAIL_UNROLL1
#define B200REF_FIN_TAIL_UNROLL1 1
#endif

namespace b200ref {

constexpr int BHIST_UNROLL = B200REF_BHIST_UNROLL;  // draws in flight per thread in k_bhist
constexpr int NB1_MAX = 256;      // coarse buckets per parameter (power of two, <= 256)
constexpr int NCELL = 1024;       // linear cells of the splitter lookup table
constexpr int NFINE = 512;        // fine buckets per coarse bucket
constexpr int CAPBIG = 12288;     // oversize coarse buckets up to this many draws: k_finish_big
constexpr int NFINE_BIG = 2048;
constexpr int FINB_NT = 1024;
constexpr int BIGLIST_PER_PARAM = 256;  // every coarse bucket of a parameter may be oversize (heavy ties)
constexpr int CAP1 = 2048;        // draws a coarse bucket may hold (on-chip capacity of k_finish)
#ifndef B200REF_OVERSAMPLE
#define B200REF_OVERSAMPLE 16
#endif
constexpr int OVERSAMPLE = B200REF_OVERSAMPLE;    // samples per coarse bucket
constexpr int CHUNK_P = 4096;     // draws per partition chunk
constexpr int PART_NT = 256;
constexpr int PART_IPT = CHUNK_P / PART_NT;
constexpr int FIN_NT = 256;
constexpr int FIN_IPT = CAP1 / FIN_NT;
constexpr int BIG_FINE = 32;      // fine buckets larger than this are handled warp-cooperatively
constexpr int SUB_MAX = 32;       // equal-width sub-buckets per coarse bucket (long chains, M > ~330k)
constexpr int NBW_MAX = NB1_MAX * SUB_MAX;  // buckets per parameter on the wide path

// Per-parameter bucket function, written by k_sample.
struct BucketFn {
  double lo, hi;      // sample min / max (cell range, and outer bounds for the fine level)
  double cs;          // NCELL / (hi - lo), 0 if hi == lo
  double c0;          // sample median: centre for the shifted power sums
  double sp[NB1_MAX]; // sp[j], j = 0 .. nb1-2 : splitters (ascending); rest = +inf
  uint16_t lut[NCELL + 2];  // lut[c] = (splitters whose cell is < c) | (splitters whose cell is < c + 1) << 8: both ends of
                            // cell c's splitter range (each <= 255) in one load
  int nb1;
  int sub;            // sub-buckets per coarse bucket (1: the coarse buckets are the buckets)
  double fsub[NB1_MAX];  // sub / (upper - lower) of coarse bucket c, 0 if degenerate
  // single-pass partition (k_part1): coarse bucket c owns records [soff[c], soff[c + 1]) of the parameter's slab
  // region; slab_ok = 0: the capacities do not fit the region (the parameter goes to the LSD path)
  uint32_t soff[NB1_MAX + 1];
  int slab_ok;
};

__host__ __device__ inline int choose_nb1(int64_t M) {
  int nb = 1;
  while (nb < NB1_MAX && M > (int64_t)nb * 768) nb <<= 1;
  return nb;
}

// sub-buckets per coarse bucket so that a bucket holds ~1000 draws (<= 1280 on average) up to
// M = 10.5M pooled draws per parameter; beyond that buckets grow and the oversize paths take over.
__host__ __device__ inline int choose_sub(int64_t M) {
  const int nb1 = choose_nb1(M);
  int sub = 1;
  while (sub < SUB_MAX && M > (int64_t)nb1 * sub * 1280) sub <<= 1;
  return sub;
}

__device__ __forceinline__ int cell_of(double x, double lo, double cs) {
  // cvt.rmi.s32.f64 saturates and maps NaN (inf * 0) to 0; the integer clamp finishes the job
  return min(max(__double2int_rd((x - lo) * cs), 0), NCELL - 1);
}

// number of splitters <= x  (monotone in x)
__device__ __forceinline__ int coarse_of(double x, double lo, double cs, const double* __restrict__ sp,
                                         const uint16_t* __restrict__ lut) {
  const int c = cell_of(x, lo, cs);
  const uint32_t ae = lut[c];
  int a = (int)(ae & 0xffu);
  const int e = (int)(ae >> 8);
  while (a < e && sp[a] <= x) ++a;
  return a;
}

// sub-bucket of x inside coarse bucket c (lower bound lc, scale fsub = sub / width): monotone in x
__device__ __forceinline__ int sub_of(double x, double lc, double fsub, int sub) {
  const double t = (x - lc) * fsub;
  return (t >= 0.0) ? ((t < (double)sub) ? (int)t : sub - 1) : 0;
}

// fine bucket inside a coarse bucket with bounds [L, U): floor((x - L) * fs), clamped
__device__ __forceinline__ int fine_of(double x, double L, double fs) {
  const double t = (x - L) * fs;
  return (t >= 0.0) ? ((t < (double)NFINE) ? (int)t : NFINE - 1) : 0;
}

// A value repeated so often that it fills whole coarse buckets shows up as DUPLICATE splitters.  k_sample moves
// the last duplicate one ulp up, which makes the coarse bucket [v, nextafter(v)) hold the copies of v and nothing
// else: a "pure" bucket.  Its draws need no ranking -- they all share the run [B0, B0 + n) -- so it may have any
// size: k_finish streams it through, and it never sends its parameter to the oversize list or the LSD path.
// (Discrete and zero-inflated columns are normal in real posteriors.)
__device__ __forceinline__ bool coarse_is_pure(const double* sp, int nb1, int c) {
  if (c < 1 || c > nb1 - 2) return false;
  return dbl2key(sp[c]) == dbl2key(sp[c - 1]) + 1ull;
}

struct BucketBufs {
  BucketFn* fn;        // [PB]
  uint32_t* cnt;       // [PB][nchunks][NB1_MAX]  per-chunk bucket counts -> exclusive offsets
  uint32_t* boff;      // [PB][bstride]           bucket start offsets in the sorted order (nbt + 1 used)
  int bstride;         // NB1_MAX + 1 (narrow) or NBW_MAX + 1 (wide)
  uint32_t* cur;       // [PB][NBW_MAX]           wide path: bucket totals, then write cursors
  double* spart;       // [PB][nchunks][2]        shifted power sums per chunk
  uint8_t* fallback;   // [PB] 1 = sort this parameter with the LSD path
  int* nbigb;          // number of oversize coarse buckets in this batch
  int* biglist;        // [PB * BIGLIST_PER_PARAM][2] (parameter, coarse bucket)
  int biglist_cap;
  int* nfb;            // number of fallback parameters in this batch
  int* plist;          // [PB] their indices (compact, arbitrary order)
  int* flag;           // bit0 non-finite input
  uint64_t* pkey;      // [PB][pstride] partitioned keys
  uint32_t* pidx;      // [PB][pstride] partitioned positions
  size_t pstride;      // records per parameter in pkey / pidx: M (dense, two-pass partition) or the slab region
  int slab;            // 1: single-pass partition -- bucket c's records start at fn->soff[c] (narrow path only)
  uint32_t slab_cap;   // records of one parameter's slab region
  uint64_t* skey;      // [PB][M] sorted keys
  uint32_t* sidx;      // [PB][M] sorted positions
  uint32_t* bmeta;     // [PB][M] narrow path, between k_bhist and k_partition: bucket << 16 | arrival rank of the
                       //         draw among its chunk's draws of that bucket (shares storage with kcode)
  uint32_t* kcode;     // [PB][M] k = i + j of each TIED sorted element's run [i, j); such elements carry
                       //         TIED_BIT in sidx.  An untied element at sorted position t has k = 2t + 1.
  int* sel;            // [PB] (LSD path compatibility: 1 = sorted data are in skey / sidx)
  double* mean;        // [PB]
  double* sd;          // [PB]
};

// ---------------------------------------------------------------- k_sample
// Bitonic sort of 4096 keys held 8 per thread by 512 threads (element e = 8 * tid + r): strides below 8 are
// compare-exchanges inside the thread, strides 8 .. 128 shuffles inside the warp (a warp holds 256 consecutive
// elements), only the 10 stages with strides >= 256 go through shared memory -- the all-shared-memory network
// measured L1-bound (89 %).  `s` holds the keys on entry and the sorted keys on return.
__device__ __forceinline__ void sample_sort_4096(uint64_t* __restrict__ s) {
  const uint32_t tid = threadIdx.x;
  uint64_t k[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) k[r] = s[tid * 8 + r];
#pragma unroll
  for (uint32_t size = 2; size <= 4096; size <<= 1) {
#pragma unroll
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 8) {
        uint64_t q[8];
        if (stride >= 256) {
          __syncthreads();  // earlier partner reads are done
#pragma unroll
          for (int r = 0; r < 8; ++r) s[tid * 8 + r] = k[r];
          __syncthreads();
#pragma unroll
          for (int r = 0; r < 8; ++r) q[r] = s[((tid * 8) ^ stride) + r];
        } else {
#pragma unroll
          for (int r = 0; r < 8; ++r) q[r] = __shfl_xor_sync(FULL, k[r], stride >> 3);
        }
        // ascending blocks keep the smaller key at the lower index
        const bool keep_min = (((tid * 8) & stride) == 0) == (((tid * 8) & size) == 0);
#pragma unroll
        for (int r = 0; r < 8; ++r) k[r] = keep_min ? min(k[r], q[r]) : max(k[r], q[r]);
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int r2 = r ^ (int)stride;
          if (r2 > r) {
            const bool up = (((tid * 8 + r) & size) == 0);
            const uint64_t a = k[r], c = k[r2];
            if ((a > c) == up) { k[r] = c; k[r2] = a; }
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) s[tid * 8 + r] = k[r];
}

__global__ void __launch_bounds__(512) k_sample(const double* __restrict__ x, uint32_t M, BucketBufs b) {
  __shared__ uint64_t s[NB1_MAX * OVERSAMPLE];
  __shared__ int scell[NB1_MAX];
  __shared__ double spfix[NB1_MAX];
  __shared__ uint8_t fixed[NB1_MAX];
  const int p = blockIdx.x;
  const double* xp = x + (size_t)p * M;
  BucketFn* fn = b.fn + p;
  const int nb1 = choose_nb1(M);
  for (int j = threadIdx.x; j < NB1_MAX; j += 512) fixed[j] = 0;
  const uint32_t S = min((uint32_t)(nb1 * OVERSAMPLE), M);  // samples actually taken
  uint32_t S2 = 1;
  while (S2 < S) S2 <<= 1;                                   // padded to a power of two
  for (uint32_t j = threadIdx.x; j < S2; j += 512) {
    uint64_t k = ~0ull;
    if (j < S) {
      const uint32_t t = (uint32_t)(((uint64_t)(2 * j + 1) * M) / (2 * S));
      k = dbl2key(xp[t]);
    }
    s[j] = k;
  }
  __syncthreads();
  if (S2 == 4096u) {  // (uniform; the usual case: 256 coarse buckets x 16 samples)
    sample_sort_4096(s);
  } else {
  // bitonic network.  With stride <= 32 the 32 pairs of a warp (in each of its passes over t) stay
  // inside one 64-element block that only this warp touches, so consecutive such steps need only
  // a warp barrier; a block barrier separates them from the wider steps.
  bool prev_local = false;
  for (uint32_t size = 2; size <= S2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      const bool local = stride <= 32;
      if (local && prev_local) __syncwarp(); else __syncthreads();
      prev_local = local;
      for (uint32_t t = threadIdx.x; t < S2 / 2; t += 512) {
        const uint32_t i = 2 * t - (t & (stride - 1));
        const uint32_t j = i + stride;
        const bool up = ((i & size) == 0);
        const uint64_t a = s[i], c = s[j];
        if ((a > c) == up) { s[i] = c; s[j] = a; }
      }
    }
  }
  }
  __syncthreads();
  const double lo = key2dbl(s[0]), hi = key2dbl(s[S - 1]);
  const double cs = (hi > lo) ? (double)NCELL / (hi - lo) : 0.0;
  if (threadIdx.x == 0) {
    fn->lo = lo; fn->hi = hi; fn->cs = isfinite(cs) ? cs : 0.0;
    fn->c0 = key2dbl(s[S / 2]);
    fn->nb1 = nb1;
    fn->sub = choose_sub(M);
    b.fallback[p] = 0;
    b.sel[p] = 1;
    if (p == 0) { *b.nfb = 0; *b.nbigb = 0; }
  }
  const double csf = isfinite(cs) ? cs : 0.0;
  for (int j = threadIdx.x; j < NB1_MAX; j += 512) {
    double v = __longlong_as_double(0x7ff0000000000000ll);
    int c = NCELL;
    if (j < nb1 - 1) {
      v = key2dbl(s[(uint32_t)(((uint64_t)(j + 1) * S) / nb1)]);
      c = cell_of(v, lo, csf);
    }
    fn->sp[j] = v;
    scell[j] = c;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nb1 - 1; j += 512) {
    // last of a group of equal splitters -> one ulp up: the bucket below it becomes pure (see coarse_is_pure)
    const double v = fn->sp[j];
    if (j >= 1 && fn->sp[j - 1] == v && (j == nb1 - 2 || fn->sp[j + 1] != v)) {
      const double up1 = key2dbl(dbl2key(v) + 1ull);
      if (is_finite_bits(up1)) { spfix[j] = up1; fixed[j] = 1; }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nb1 - 1; j += 512)
    if (fixed[j]) { fn->sp[j] = spfix[j]; scell[j] = cell_of(spfix[j], lo, csf); }
  __syncthreads();
  for (int j = threadIdx.x; j < NB1_MAX; j += 512) {
    // scale of the sub-bucket / fine-bucket function of coarse bucket j
    double fsv = 0.0;
    if (j < nb1) {
      const double L = (j == 0) ? lo : fn->sp[j - 1];
      const double U = (j == nb1 - 1) ? hi : fn->sp[j];
      if (U > L) fsv = (double)choose_sub(M) / (U - L);
      if (!isfinite(fsv)) fsv = 0.0;
    }
    fn->fsub[j] = fsv;
  }
  for (int c = threadIdx.x; c < NCELL + 2; c += 512) {
    // number of splitters with cell < c and with cell < c + 1  (scell ascending over j < nb1-1, NCELL beyond)
    int a = 0, e = nb1 - 1;
    while (a < e) { const int mid = (a + e) >> 1; if (scell[mid] < c) a = mid + 1; else e = mid; }
    int a1 = a;
    e = nb1 - 1;
    while (a1 < e) { const int mid = (a1 + e) >> 1; if (scell[mid] < c + 1) a1 = mid + 1; else e = mid; }
    fn->lut[c] = (uint16_t)(a | (a1 << 8));
  }
  if (b.slab) {
    // Single-pass partition: capacity of every coarse bucket's slab.  A bucket holds M / nb1 draws on average with
    // a relative spread of 1 / sqrt(OVERSAMPLE) (its bounds are sample quantiles): three times the mean is beyond
    // 8 sigma, and also covers a repeated value too rare to show up as duplicate splitters (< 2 / nb1 of the
    // draws).  A pure bucket (copies of one value) is sized from the number of copies in the sample, + 6 sigma.
    __shared__ uint32_t capv[NB1_MAX];
    __syncthreads();  // the splitters in fn->sp are final
    for (int j = threadIdx.x; j < NB1_MAX; j += 512) {
      uint32_t c = 0;
      if (j < nb1) {
        c = min(M, 3u * ((M + (uint32_t)nb1 - 1u) / (uint32_t)nb1) + 64u);
        if (coarse_is_pure(fn->sp, nb1, j)) {
          const uint64_t kv = dbl2key(fn->sp[j - 1]);
          uint32_t l0 = 0, h0 = S;  // first sample >= v
          while (l0 < h0) { const uint32_t mid = (l0 + h0) >> 1; if (s[mid] < kv) l0 = mid + 1; else h0 = mid; }
          uint32_t l1 = l0, h1 = S;  // first sample > v
          while (l1 < h1) { const uint32_t mid = (l1 + h1) >> 1; if (s[mid] <= kv) l1 = mid + 1; else h1 = mid; }
          const double cs_ = (double)(l1 - l0);
          const double est = ceil((cs_ + 6.0 * sqrt(cs_) + 16.0) * ((double)M / (double)S));
          const uint32_t ce = (est < (double)M) ? (uint32_t)est : M;
          c = max(c, ce);
        }
      }
      capv[j] = c;
    }
    __syncthreads();
    {  // exclusive prefix of the capacities (each <= M < 2^31, 256 of them: 64-bit sums)
      __shared__ unsigned long long cw[NB1_MAX / 32];
      const int t = threadIdx.x, lane = t & 31, w = t >> 5;
      unsigned long long incl = 0, own = 0;
      if (t < NB1_MAX) {
        own = capv[t];
        incl = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long nn = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += nn;
        }
        if (lane == 31) cw[w] = incl;
      }
      __syncthreads();
      if (t < NB1_MAX) {
        unsigned long long pre = 0;
        for (int k = 0; k < w; ++k) pre += cw[k];
        const unsigned long long ex = pre + incl - own;
        fn->soff[t] = (uint32_t)min(ex, 0xffffffffull);
        if (t == NB1_MAX - 1) {
          const unsigned long long tot = ex + own;
          fn->soff[NB1_MAX] = (uint32_t)min(tot, 0xffffffffull);
          fn->slab_ok = (tot <= (unsigned long long)b.slab_cap) ? 1 : 0;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- k_bhist
#ifndef B200REF_BHIST_OCC
#define B200REF_BHIST_OCC 8
#endif
__global__ void __launch_bounds__(PART_NT, B200REF_BHIST_OCC) k_bhist(const double* __restrict__ x, uint32_t M, uint32_t nchunks,
                                                    BucketBufs b) {
  __shared__ double sp[NB1_MAX];
  __shared__ uint16_t lut[NCELL + 2];
  __shared__ uint32_t cnt[NB1_MAX];
  __shared__ double red[33 * 2];
  const int p = blockIdx.y;
  const uint32_t chunk = blockIdx.x;
  const BucketFn* fn = b.fn + p;
  const double* xp = x + (size_t)p * M;
  for (int i = threadIdx.x; i < NB1_MAX; i += PART_NT) { sp[i] = fn->sp[i]; cnt[i] = 0; }
  for (int i = threadIdx.x; i < NCELL + 2; i += PART_NT) lut[i] = fn->lut[i];
  const double lo = fn->lo, cs = fn->cs, c0 = fn->c0;
  __syncthreads();
  const uint32_t base = chunk * CHUNK_P;
  uint32_t* meta = b.bmeta + (size_t)p * M;
  double s1 = 0.0, s2 = 0.0;
  bool bad = false;
  // Every draw leaves its bucket and its arrival rank inside this chunk's share of the bucket (the returning
  // atomic): k_partition places the draw from that word alone -- no second evaluation of the bucket function,
  // no second counting pass.  Explicit groups: BHIST_UNROLL loads, then their bucket functions, then their
  // atomics, then the stores, so that several draws' memory round trips overlap per thread (the compiler does
  // not hoist loads across predicated blocks with shared-memory atomics).
#pragma unroll 1
  for (int it0 = 0; it0 < PART_IPT; it0 += BHIST_UNROLL) {
    double v[BHIST_UNROLL];
    uint32_t c[BHIST_UNROLL], r[BHIST_UNROLL];
#pragma unroll
    for (int j = 0; j < BHIST_UNROLL; ++j) {
      const uint32_t e = base + (it0 + j) * PART_NT + threadIdx.x;
      v[j] = (e < M) ? xp[e] : c0;
    }
#pragma unroll
    for (int j = 0; j < BHIST_UNROLL; ++j) c[j] = (uint32_t)coarse_of(v[j], lo, cs, sp, lut);
#pragma unroll
    for (int j = 0; j < BHIST_UNROLL; ++j) {
      const uint32_t e = base + (it0 + j) * PART_NT + threadIdx.x;
      r[j] = (e < M) ? atomicAdd(&cnt[c[j]], 1u) : 0u;
    }
#pragma unroll
    for (int j = 0; j < BHIST_UNROLL; ++j) {
      const uint32_t e = base + (it0 + j) * PART_NT + threadIdx.x;
      if (e < M) {
        bad |= !is_finite_bits(v[j]);
        const double d = v[j] - c0;
        s1 += d;
        s2 = fma(d, d, s2);
        meta[e] = (c[j] << 16) | r[j];
      }
    }
  }
  // deterministic block sums (fixed tree)
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[w] = s1; red[33 + w] = s2; }
  const int anybad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int k = 0; k < PART_NT / 32; ++k) { a += red[k]; c += red[33 + k]; }
    b.spart[((size_t)p * nchunks + chunk) * 2 + 0] = a;
    b.spart[((size_t)p * nchunks + chunk) * 2 + 1] = c;
    if (anybad) atomicOr(b.flag, 1);
  }
  uint32_t* out = b.cnt + ((size_t)p * nchunks + chunk) * NB1_MAX;
  for (int i = threadIdx.x; i < NB1_MAX; i += PART_NT) out[i] = cnt[i];
}

// ---------------------------------------------------------------- k_bscan
static_assert(NB1_MAX == 256 && PART_NT == NB1_MAX, "thread counts double as bin counts");

__global__ void __launch_bounds__(NB1_MAX) k_bscan(uint32_t M, uint32_t nchunks, BucketBufs b) {
  __shared__ uint32_t wtot[NB1_MAX / 32];
  __shared__ int over;
  const int p = blockIdx.x;
  const int t = threadIdx.x;
  if (t == 0) over = 0;
  __syncthreads();
  uint32_t* cnt = b.cnt + (size_t)p * nchunks * NB1_MAX;
  // (loads in groups of 8 ahead of the stores: the compiler cannot hoist them across stores to the same array)
  uint32_t run = 0;
  for (uint32_t c0 = 0; c0 < nchunks; c0 += 8) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < nchunks) ? cnt[(size_t)(c0 + j) * NB1_MAX + t] : 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c0 + j < nchunks) cnt[(size_t)(c0 + j) * NB1_MAX + t] = run;
      run += v[j];
    }
  }
  const bool pure = coarse_is_pure(b.fn[p].sp, b.fn[p].nb1, t);
  if (pure) {
    // any size: k_finish streams it
  } else if (run > (uint32_t)CAPBIG) {
    over = 1;
  } else if (run > (uint32_t)CAP1) {
    const int slot = atomicAdd(b.nbigb, 1);
    if (slot < b.biglist_cap) { b.biglist[2 * slot] = p; b.biglist[2 * slot + 1] = t; }
    else over = 1;  // list full: sort the whole parameter by the LSD path instead
  }
  // exclusive scan of the bucket totals over the 256 threads
  const int lane = t & 31, w = t >> 5;
  uint32_t incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) wtot[w] = incl;
  __syncthreads();
  uint32_t pre = (lane < w) ? wtot[lane] : 0u;  // prefix over the earlier warps (at most 32 of them)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(FULL, pre, o);
  const uint32_t excl = pre + incl - run;
  b.boff[(size_t)p * b.bstride + t] = excl;
  if (t == NB1_MAX - 1) b.boff[(size_t)p * b.bstride + NB1_MAX] = excl + run;
  for (uint32_t c0 = 0; c0 < nchunks; c0 += 8) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < nchunks) ? cnt[(size_t)(c0 + j) * NB1_MAX + t] : 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < nchunks) cnt[(size_t)(c0 + j) * NB1_MAX + t] = v[j] + excl;
  }
  if (t == 0) {
    // mean = c0 + S1/M ; var = (S2 - S1^2/M)/M  (sums about the sample median: no cancellation)
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t c = 0; c < nchunks; ++c) {
      s1 += b.spart[((size_t)p * nchunks + c) * 2 + 0];
      s2 += b.spart[((size_t)p * nchunks + c) * 2 + 1];
    }
    const double dm = (double)M;
    const double m1 = s1 / dm;
    b.mean[p] = b.fn[p].c0 + m1;
    const double var = (s2 - s1 * m1) / dm;
    b.sd[p] = sqrt(var > 0.0 ? var : 0.0);
  }
  __syncthreads();
  if (t == 0 && over) {
    b.fallback[p] = 1;
    b.plist[atomicAdd(b.nfb, 1)] = p;
  }
}

// ---------------------------------------------------------------- k_partition
struct PartSmem {
  uint64_t skey[CHUNK_P];   // the chunk's keys in bucket order
  uint32_t smeta[CHUNK_P];  // bucket << 16 | index inside the chunk
  uint32_t coff[NB1_MAX];   // start of each bucket inside the chunk
  uint32_t delta[NB1_MAX];  // (start of this chunk's run in the bucket's global range) - coff
  uint32_t wtot[PART_NT / 32];
};

#ifndef B200REF_PART_OCC
#define B200REF_PART_OCC 3
#endif
#ifndef B200REF_PART_NB
#define B200REF_PART_NB 2
#endif
constexpr int PART_NB = B200REF_PART_NB;          // load batches per chunk
constexpr int PART_HB = PART_IPT / PART_NB;  // draws per thread and load batch

// One CTA per 4096-draw chunk: the (bucket, arrival rank) word k_bhist left for every draw gives its slot
// in the chunk's bucket-ordered staging tile directly; the tile then leaves in runs (coalesced stores).
__global__ void __launch_bounds__(PART_NT, B200REF_PART_OCC) k_partition(const double* __restrict__ x, uint32_t M, uint32_t nchunks,
                                                        BucketBufs b) {
  extern __shared__ __align__(16) unsigned char part_smem_raw[];
  PartSmem& sm = *reinterpret_cast<PartSmem*>(part_smem_raw);
  const int p = blockIdx.y;
  if (b.fallback[p]) return;
  const uint32_t chunk = blockIdx.x;
  const double* xp = x + (size_t)p * M;
  const uint32_t* meta = b.bmeta + (size_t)p * M;
  const int tid = threadIdx.x;
  const uint32_t base = chunk * CHUNK_P;
  const uint32_t n = min((uint32_t)CHUNK_P, M - base);
  // first batch of loads in flight while the offsets are prepared
  double v[PART_HB];
  uint32_t m[PART_HB];
#pragma unroll
  for (int it = 0; it < PART_HB; ++it) {
    const uint32_t e = min(it * PART_NT + tid, n - 1);
    v[it] = xp[base + e];
    m[it] = meta[base + e];
  }
  {  // this chunk's count per bucket = next chunk's offset - this chunk's offset; exclusive scan -> coff
    const uint32_t* g = b.cnt + ((size_t)p * nchunks + chunk) * NB1_MAX;
    const uint32_t g0 = g[tid];
    const uint32_t g1 = (chunk + 1 < nchunks) ? g[NB1_MAX + tid] : b.boff[(size_t)p * b.bstride + tid + 1];
    const int lane = tid & 31, w = tid >> 5;
    const uint32_t c = g1 - g0;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t nn = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += nn;
    }
    if (lane == 31) sm.wtot[w] = incl;
    __syncthreads();
    uint32_t pre = (lane < w) ? sm.wtot[lane] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(FULL, pre, o);
    const uint32_t co = pre + incl - c;
    sm.coff[tid] = co;
    sm.delta[tid] = g0 - co;
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < PART_NB; ++h) {
    if (h >= 1) {
#pragma unroll
      for (int it = 0; it < PART_HB; ++it) {
        const uint32_t e = min((h * PART_HB + it) * PART_NT + tid, n - 1);
        v[it] = xp[base + e];
        m[it] = meta[base + e];
      }
    }
#pragma unroll
    for (int it = 0; it < PART_HB; ++it) {
      const uint32_t e = (h * PART_HB + it) * PART_NT + tid;
      if (e < n) {
        const uint32_t slot = sm.coff[m[it] >> 16] + (m[it] & 0xffffu);
        sm.skey[slot] = dbl2key(v[it]);
        sm.smeta[slot] = (m[it] & 0xffff0000u) | e;
      }
    }
  }
  __syncthreads();
  uint64_t* ok = b.pkey + (size_t)p * b.pstride;
  uint32_t* oi = b.pidx + (size_t)p * b.pstride;
#pragma unroll 4
  for (uint32_t s = tid; s < n; s += PART_NT) {
    const uint32_t mm = sm.smeta[s];
    const uint32_t pos = sm.delta[mm >> 16] + s;
    B200REF_CHECK(pos < M, "k_partition: run slot outside the parameter");
    ok[pos] = sm.skey[s];
    oi[pos] = base + (mm & 0xffffu);
  }
}

// ---------------------------------------------------------------- k_part1 / k_bscan1: single-pass partition
// The draws are read ONCE: bucket function, chunk histogram (the returning shared-memory atomic is the draw's place
// in its chunk's run), shifted power sums and the finiteness check as in k_bhist, then -- instead of a second pass
// over the draws after a global scan -- every (chunk, bucket) run is reserved in the bucket's slack-padded slab
// with one global atomic (capacities fn->soff, sized by k_sample) and the chunk leaves through the same
// bucket-ordered staging tile as in k_partition.  The order inside a bucket then depends on scheduling, the
// RESULTS do not: k_finish ranks by value and draws with equal values receive equal scores.  A run that does
// not fit its slab marks the parameter for the LSD path (fallback) and is not written.
struct Part1Smem {
  uint64_t skey[CHUNK_P];   // the chunk's keys in bucket order
  uint32_t smeta[CHUNK_P];  // bucket << 16 | index inside the chunk
  double sp[NB1_MAX];
  uint32_t cnt[NB1_MAX];    // the chunk's bucket counts
  uint32_t coff[NB1_MAX];   // start of each bucket inside the chunk
  uint32_t delta[NB1_MAX];  // (start of this chunk's run in the parameter's slab region) - coff; PART1_DEAD: not written
  uint32_t wtot[NB1_MAX / 32];
  double red[2 * 32];
  uint16_t lut[NCELL + 2];
};
constexpr uint32_t PART1_DEAD = 0x80000000u;  // + any slot of the chunk stays above every slab region (pstride < 2^31)

#ifndef B200REF_PART1_NT
#define B200REF_PART1_NT 256
#endif
#ifndef B200REF_PART1_OCC
#define B200REF_PART1_OCC 4
#endif
#ifndef B200REF_PART1_REREAD
#define B200REF_PART1_REREAD 1  // 1: the draws are read again (L2 hits) for the staging step instead of held in registers
#endif
constexpr int P1_NT = B200REF_PART1_NT;     // threads per chunk CTA (the first NB1_MAX of them own one bucket each)
constexpr int P1_IPT = CHUNK_P / P1_NT;
static_assert(P1_NT >= NB1_MAX && CHUNK_P % P1_NT == 0, "k_part1: one thread per bucket, whole draws per thread");

__global__ void __launch_bounds__(P1_NT, B200REF_PART1_OCC) k_part1(const double* __restrict__ x, uint32_t M, uint32_t nchunks,
                                                                    BucketBufs b) {
  extern __shared__ __align__(16) unsigned char part_smem_raw[];
  Part1Smem& sm = *reinterpret_cast<Part1Smem*>(part_smem_raw);
  const int p = blockIdx.y;
  const uint32_t chunk = blockIdx.x;
  const BucketFn* fn = b.fn + p;
  const double* xp = x + (size_t)p * M;
  const int tid = threadIdx.x;
  const uint32_t base = chunk * CHUNK_P;
  const uint32_t n = min((uint32_t)CHUNK_P, M - base);
  // the chunk's draws are in flight while the bucket function's tables arrive
  double v[P1_IPT];
#pragma unroll
  for (int it = 0; it < P1_IPT; ++it) v[it] = xp[base + min((uint32_t)(it * P1_NT + tid), n - 1)];
  for (int i = tid; i < NB1_MAX; i += P1_NT) { sm.sp[i] = fn->sp[i]; sm.cnt[i] = 0; }
  for (int i = tid; i < NCELL + 2; i += P1_NT) sm.lut[i] = fn->lut[i];
  const double lo = fn->lo, cs = fn->cs, c0 = fn->c0;
  const int slab_ok = fn->slab_ok;
  __syncthreads();
  uint32_t cr[P1_IPT];  // bucket << 16 | arrival rank in the chunk's share of the bucket
  double s1 = 0.0, s2 = 0.0;
  bool bad = false;
#pragma unroll
  for (int it = 0; it < P1_IPT; ++it) {
    const uint32_t e = it * P1_NT + tid;
    if (e < n) {
      const uint32_t c = (uint32_t)coarse_of(v[it], lo, cs, sm.sp, sm.lut);
      cr[it] = (c << 16) | atomicAdd(&sm.cnt[c], 1u);
      bad |= !is_finite_bits(v[it]);
      const double d = v[it] - c0;
      s1 += d;
      s2 = fma(d, d, s2);
    }
  }
  // deterministic block sums (fixed tree), as in k_bhist
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const int lane = tid & 31, w = tid >> 5;
  if (lane == 0) { sm.red[w] = s1; sm.red[32 + w] = s2; }
  const int anybad = __syncthreads_or(bad);
  if (tid == 0) {
    double a = 0.0, c = 0.0;
    for (int k = 0; k < P1_NT / 32; ++k) { a += sm.red[k]; c += sm.red[32 + k]; }
    b.spart[((size_t)p * nchunks + chunk) * 2 + 0] = a;
    b.spart[((size_t)p * nchunks + chunk) * 2 + 1] = c;
    if (anybad) atomicOr(b.flag, 1);
  }
  // exclusive scan of the chunk's counts -> coff; one global atomic per non-empty bucket reserves its run (its
  // round trip is hidden under the staging step: the run's start is only needed when the tile leaves)
  const bool owner = tid < NB1_MAX;  // thread t < 256 owns bucket t
  const uint32_t cnt_t = owner ? sm.cnt[tid] : 0u;
  const bool live = cnt_t != 0 && slab_ok;
  uint32_t g = 0, s0 = 0, cap = 0;
  if (live) {
    s0 = fn->soff[tid];
    cap = fn->soff[tid + 1] - s0;
    g = atomicAdd(&b.cur[(size_t)p * NB1_MAX + tid], cnt_t);
  }
  uint32_t co;
  {
    uint32_t incl = cnt_t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t nn = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += nn;
    }
    if (lane == 31 && owner) sm.wtot[w] = incl;
    __syncthreads();
    uint32_t pre = (lane < w && lane < NB1_MAX / 32) ? sm.wtot[lane] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(FULL, pre, o);
    co = pre + incl - cnt_t;
    if (owner) sm.coff[tid] = co;
  }
#if B200REF_PART1_REREAD
#pragma unroll
  for (int it = 0; it < P1_IPT; ++it) v[it] = xp[base + min((uint32_t)(it * P1_NT + tid), n - 1)];
#endif
  __syncthreads();
#pragma unroll
  for (int it = 0; it < P1_IPT; ++it) {
    const uint32_t e = it * P1_NT + tid;
    if (e < n) {
      const uint32_t slot = sm.coff[cr[it] >> 16] + (cr[it] & 0xffffu);
      sm.skey[slot] = dbl2key(v[it]);
      sm.smeta[slot] = (cr[it] & 0xffff0000u) | e;
    }
  }
  {
    uint32_t dl = PART1_DEAD;
    if (live) {
      if (g + cnt_t <= cap) dl = s0 + g - co;   // (mod 2^32; dl + slot is the record's place in the region)
      else b.fallback[p] = 1;                   // slab overflow: k_bscan1 lists the parameter for the LSD path
    }
    if (owner) sm.delta[tid] = dl;
  }
  __syncthreads();
  uint64_t* ok = b.pkey + (size_t)p * b.pstride;
  uint32_t* oi = b.pidx + (size_t)p * b.pstride;
  const uint32_t lim = (uint32_t)b.pstride;
#pragma unroll 4
  for (uint32_t s = tid; s < n; s += P1_NT) {
    const uint32_t mm = sm.smeta[s];
    const uint32_t pos = sm.delta[mm >> 16] + s;
    if (pos < lim) {  // dead runs (slab overflow, no slab) land above every region
      B200REF_CHECK(pos >= fn->soff[mm >> 16] && pos < fn->soff[(mm >> 16) + 1], "k_part1: record outside its bucket's slab");
      ok[pos] = sm.skey[s];
      oi[pos] = base + (mm & 0xffffu);
    }
  }
}

// Bucket totals (the slab cursors) -> sorted offsets, classification of the buckets, mean / std: k_bscan's job
// without the per-chunk offsets.
__global__ void __launch_bounds__(NB1_MAX) k_bscan1(uint32_t M, uint32_t nchunks, BucketBufs b) {
  __shared__ uint32_t wtot[NB1_MAX / 32];
  __shared__ int over;
  const int p = blockIdx.x;
  const int t = threadIdx.x;
  if (t == 0) over = 0;
  __syncthreads();
  const uint32_t run = b.cur[(size_t)p * NB1_MAX + t];
  const bool pure = coarse_is_pure(b.fn[p].sp, b.fn[p].nb1, t);
  if (pure) {
    // any size (inside its slab): k_finish streams it
  } else if (run > (uint32_t)CAPBIG) {
    over = 1;
  } else if (run > (uint32_t)CAP1) {
    const int slot = atomicAdd(b.nbigb, 1);
    if (slot < b.biglist_cap) { b.biglist[2 * slot] = p; b.biglist[2 * slot + 1] = t; }
    else over = 1;  // list full: sort the whole parameter by the LSD path instead
  }
  const int lane = t & 31, w = t >> 5;
  uint32_t incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) wtot[w] = incl;
  __syncthreads();
  uint32_t pre = (lane < w) ? wtot[lane] : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(FULL, pre, o);
  const uint32_t excl = pre + incl - run;
  b.boff[(size_t)p * b.bstride + t] = excl;
  if (t == NB1_MAX - 1) {
    b.boff[(size_t)p * b.bstride + NB1_MAX] = excl + run;
    if (excl + run != M) over = 1;  // a run was dropped (slab overflow / no slab): nothing of this partition is used
  }
  if (t == 0) {
    // mean = c0 + S1/M ; var = (S2 - S1^2/M)/M  (sums about the sample median: no cancellation)
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t c = 0; c < nchunks; ++c) {
      s1 += b.spart[((size_t)p * nchunks + c) * 2 + 0];
      s2 += b.spart[((size_t)p * nchunks + c) * 2 + 1];
    }
    const double dm = (double)M;
    const double m1 = s1 / dm;
    b.mean[p] = b.fn[p].c0 + m1;
    const double var = (s2 - s1 * m1) / dm;
    b.sd[p] = sqrt(var > 0.0 ? var : 0.0);
  }
  __syncthreads();
  if (t == 0 && (over || b.fallback[p] || !b.fn[p].slab_ok)) {
    b.fallback[p] = 1;
    b.plist[atomicAdd(b.nfb, 1)] = p;
  }
}

// ---------------------------------------------------------------- k_finish / k_finish_big
struct FinishOut {
  double* rank;   // optional [PB][M]: average ranks scattered to original positions (ranks API)
  // Statistics-only passes (mean / std / quantiles, no scores): the sorted positions the quantile step will
  // read; k_finish then ranks only the buckets that hold one of them.  npos == 0: every bucket is ranked.
  uint32_t pos[2 * 8 + 2];
  int npos;
};

// Shared-memory layout of one finishing CTA (dynamic).
template <int CAP, int NF>
struct FinSmem {
  double kd[CAP + 4];      // canonical values in fine-bucket order, then 4 x +inf
  uint32_t i2[CAP];        // their original positions
  uint32_t fh[NF];         // fine histogram, then running cursors (two-pass mode)
  uint32_t foff[NF + 1];   // exclusive offsets
  uint32_t wtot[32];
  uint32_t bigq[NF];       // fine buckets larger than BIG_FINE
  uint32_t nbig;
};

// Exclusive scan of fh[NF] -> foff[NF + 1] with NT threads (NF / NT consecutive bins each);
// also queues the fine buckets that need the warp-cooperative path.
template <int NT, int NF, int CAP>
__device__ __forceinline__ void fine_scan(FinSmem<CAP, NF>& sm) {
  constexpr int PER = NF / NT;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  uint32_t c[PER];
  uint32_t tsum = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) { c[k] = sm.fh[tid * PER + k]; tsum += c[k]; }
  uint32_t incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t nn = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += nn;
  }
  if (lane == 31) sm.wtot[w] = incl;
  __syncthreads();
  // exclusive prefix over the warps before this one: lane k < w holds warp k's total, then a warp sum
  uint32_t pre = (lane < w) ? sm.wtot[lane] : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(FULL, pre, o);
  uint32_t run = pre + incl - tsum;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    sm.foff[tid * PER + k] = run;
    if (c[k] > (uint32_t)BIG_FINE) sm.bigq[atomicAdd(&sm.nbig, 1u)] = tid * PER + k;
    run += c[k];
  }
  if (tid == NT - 1) sm.foff[NF] = run;
}

// Warp-cooperative in-place ascending sort of n (value, position) pairs in shared memory: a bitonic
// network whose compare-exchanges all point the same way, so a length that is not a power of two
// just skips the pairs reaching past the end (the missing tail behaves as +inf).
__device__ __forceinline__ void warp_sort_pairs(double* __restrict__ kd, uint32_t* __restrict__ iv, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t N = 2;
  while (N < n) N <<= 1;
  auto cx = [&](uint32_t i, uint32_t j) {
    if (j < n) {
      const double a = kd[i], c = kd[j];
      if (a > c) {
        kd[i] = c; kd[j] = a;
        const uint32_t t = iv[i]; iv[i] = iv[j]; iv[j] = t;
      }
    }
  };
  for (uint32_t size = 2; size <= N; size <<= 1) {
    const uint32_t hs = size >> 1;
    for (uint32_t t = lane; t < N / 2; t += 32) {
      const uint32_t blk = t / hs, r = t - blk * hs;
      cx(blk * size + r, blk * size + (size - 1 - r));
    }
    __syncwarp();
    for (uint32_t stride = size >> 2; stride > 0; stride >>= 1) {
      for (uint32_t t = lane; t < N / 2; t += 32) {
        const uint32_t i = 2 * t - (t & (stride - 1));
        cx(i, i + stride);
      }
      __syncwarp();
    }
  }
}

// Fine-bucket function of one bucket: floor(fma(x - L, A, B)) clamped to [0, NF).  Bucket cb is
// sub-bucket j of coarse bucket c (lower bound L, sub-bucket scale fs): A = fs * NF, B = -j * NF.
// Monotone in x (one subtraction, one fused multiply-add, A >= 0), which is all exactness needs.
struct FineFn {
  double L, A, B;
  // k_finish_big only: the two outermost buckets of a parameter also hold every draw beyond the sample's
  // extremes (~M / 4096 of them, far more for heavy tails).  Clamped, they would all share the first / last
  // fine bucket and be sorted by a single warp (2.9 ms of config 4's step).  edge: 0 none, 1 the bucket below
  // which there is nothing (x0 = sample minimum), 2 the bucket above which there is nothing (x0 = sample
  // maximum); w = 1 / (bucket width).
  int edge;
  double x0, w;
};

template <int NF>
__device__ __forceinline__ FineFn make_fine(const BucketFn* fn, int cb, int nsub) {
  const int sh = 31 - __clz(nsub);  // nsub is a power of two
  const int c = cb >> sh;
  FineFn f;
  f.L = (c == 0) ? fn->lo : fn->sp[c - 1];
  f.A = fn->fsub[c] * (double)NF;
  f.B = -(double)((cb - (c << sh)) * NF);  // 0 when the coarse buckets are the buckets (nsub == 1)
  f.edge = 0;
  f.x0 = 0.0;
  f.w = fn->fsub[c];
  const int nb1 = fn->nb1;
  if (cb == 0) { f.edge = 1; f.x0 = fn->lo; }
  else if (cb == nb1 * nsub - 1) { f.edge = 2; f.x0 = fn->hi; }
  return f;
}

// Edge buckets (k_finish_big): the bucket's own range takes half of the fine buckets, the draws beyond the
// sample's extreme the other half, spread by the logarithm of their distance in bucket widths (64 fine buckets
// per doubling at NF = 2048).  Monotone in x on both pieces and across x0.
template <int NF>
__device__ __forceinline__ int fine_edge(const FineFn& ff, double x) {
  constexpr double SPREAD = (double)(NF / 2) / 16.0;
  const int inner = min(max(__double2int_rd(fma(x - ff.L, ff.A, ff.B) * 0.5), 0), NF / 2 - 1);
  if (ff.edge == 1) {
    if (x >= ff.x0) return NF / 2 + inner;
    const int far = min(__double2int_rd(log2(1.0 + (ff.x0 - x) * ff.w) * SPREAD), NF / 2 - 1);
    return NF / 2 - 1 - max(far, 0);
  }
  if (x <= ff.x0) return inner;
  const int far = min(__double2int_rd(log2(1.0 + (x - ff.x0) * ff.w) * SPREAD), NF / 2 - 1);
  return NF / 2 + max(far, 0);
}

template <int NF, bool EDGE>
__device__ __forceinline__ int fine_bucket(const FineFn& ff, double x) {
  if (EDGE && ff.edge) return fine_edge<NF>(ff, x);  // uniform per bucket
  // cvt.rmi.s32.f64 saturates (NaN -> 0), the integer clamp finishes the job
  return min(max(__double2int_rd(fma(x - ff.L, ff.A, ff.B)), 0), NF - 1);
}

// Rank every draw of one bucket (n draws, sorted offset B0) exactly.
// TWO_PASS = false: the bucket is read once into registers (n <= NT * IPT);
// TWO_PASS = true : any n <= CAP, the bucket is read twice and slots come from cursors.
template <int NT, int NF, int CAP, int IPT, bool TWO_PASS, bool EDGE>
__device__ __forceinline__ void finish_bucket(FinSmem<CAP, NF>& sm, const FineFn ff, uint32_t B0, uint32_t S0, uint32_t n,
                                              uint32_t M, int p, const BucketBufs& b, const FinishOut& fo) {
  const int tid = threadIdx.x;
  auto fine = [&](double x) -> int { return fine_bucket<NF, EDGE>(ff, x); };
  for (int i = tid; i < NF; i += NT) sm.fh[i] = 0;
  if (tid == 0) sm.nbig = 0;
  __syncthreads();
  // B0: the bucket's offset in the sorted order; S0: where its partitioned records start (= B0 unless slabs)
  const uint64_t* pk = b.pkey + (size_t)p * b.pstride + S0;
  const uint32_t* pi = b.pidx + (size_t)p * b.pstride + S0;
  if constexpr (!TWO_PASS) {
    // Only the packed word (fine bucket | arrival rank in it << 16, both below 2^16) of every draw stays in
    // registers across the scan; key and position are read again afterwards (L1 / L2 hits).  Holding them too
    // spilled 3 words per draw at the 40 registers that 6 CTAs / SM allow (1.94 vs 1.70 ms on config 2).
    uint32_t fr[IPT];
    // keys in groups of LB: all loads of a group are in flight before the first bucket function needs its key
    constexpr int LB = (IPT % B200REF_FIN_LOAD2 == 0) ? B200REF_FIN_LOAD2 : 1;
#pragma unroll
    for (int it = 0; it < IPT; it += LB) {
      if ((uint32_t)(it * NT) >= n) break;  // uniform: typical buckets fill 2-4 of the IPT slots
      uint64_t kk[LB];
#pragma unroll
      for (int j = 0; j < LB; ++j) kk[j] = pk[min((uint32_t)((it + j) * NT + tid), n - 1)];
#pragma unroll
      for (int j = 0; j < LB; ++j) {
        if ((uint32_t)((it + j) * NT + tid) < n) {
          const uint32_t f = (uint32_t)fine(key2dbl(kk[j]));
          fr[it + j] = f | (atomicAdd(&sm.fh[f], 1u) << 16);
        }
      }
    }
    __syncthreads();
    fine_scan<NT, NF, CAP>(sm);
    __syncthreads();
#pragma unroll
    for (int it = 0; it < IPT; it += LB) {
      if ((uint32_t)(it * NT) >= n) break;
      uint64_t kk[LB];
      uint32_t ii[LB];
#pragma unroll
      for (int j = 0; j < LB; ++j) {
        const uint32_t e = min((uint32_t)((it + j) * NT + tid), n - 1);
        kk[j] = pk[e];
        ii[j] = pi[e];
      }
#pragma unroll
      for (int j = 0; j < LB; ++j) {
        if ((uint32_t)((it + j) * NT + tid) < n) {
          const uint32_t slot = sm.foff[fr[it + j] & 0xffffu] + (fr[it + j] >> 16);
          B200REF_CHECK(slot < n && n <= (uint32_t)CAP, "finish_bucket: staging slot outside the tile");
          sm.kd[slot] = key2dbl(kk[j]);
          sm.i2[slot] = ii[j];
        }
      }
    }
  } else {
    for (uint32_t e = tid; e < n; e += NT) atomicAdd(&sm.fh[fine(key2dbl(pk[e]))], 1u);
    __syncthreads();
    fine_scan<NT, NF, CAP>(sm);
    __syncthreads();
    for (int i = tid; i < NF; i += NT) sm.fh[i] = 0;  // now cursors
    __syncthreads();
    for (uint32_t e = tid; e < n; e += NT) {
      const double v = key2dbl(pk[e]);
      const int f = fine(v);
      const uint32_t slot = sm.foff[f] + atomicAdd(&sm.fh[f], 1u);
      sm.kd[slot] = v;
      sm.i2[slot] = pi[e];
    }
  }
  if (tid < 4) sm.kd[n + tid] = __longlong_as_double(0x7ff0000000000000ll);
  __syncthreads();
  uint64_t* ok = b.skey + (size_t)p * M;
  uint32_t* oi = b.sidx + (size_t)p * M;
  uint32_t* oc = b.kcode + (size_t)p * M;
  double* orank = fo.rank ? fo.rank + (size_t)p * M : nullptr;
  // small fine buckets: one thread per slot counts the smaller / equal values of its bucket
  for (uint32_t s = tid; s < n; s += NT) {
    const double k = sm.kd[s];
    const int fb = fine(k);
    const uint32_t a = sm.foff[fb], e = sm.foff[fb + 1];
    if (e - a > (uint32_t)BIG_FINE) continue;
    // The 4 slots from the bucket's start are read unconditionally: whatever lies past its end belongs
    // to later fine buckets (strictly larger values) or is the +inf padding, and counts for nothing.
    uint32_t less = 0, le = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const double kt = sm.kd[a + d];
      less += (kt < k);
      le += (kt <= k);
    }
#if B200REF_FIN_TAIL_UNROLL1
#pragma unroll 1  // fine buckets beyond 4 draws are rare: no unrolled prologue / epilogue on the common zero-trip path
#endif
    for (uint32_t t = a + 4; t < e; ++t) {
      const double kt = sm.kd[t];
      less += (kt < k);
      le += (kt <= k);
    }
    const uint32_t eq = le - less;
    uint32_t eqb = 0;  // equal values in earlier slots: only tied draws need it
    if (eq > 1u) {
#if B200REF_FIN_TAIL_UNROLL1
#pragma unroll 1
#endif
      for (uint32_t t = a; t < s; ++t) eqb += (sm.kd[t] == k);
    }
    const uint32_t i = B0 + a + less;
    const uint32_t pos = i + eqb;
    B200REF_CHECK(pos >= B0 && pos < B0 + n && (sm.i2[s] & ~TIED_BIT) < M, "finish_bucket: sorted position outside the bucket");
    ok[pos] = dbl2key(k);
    oi[pos] = (eq > 1u) ? (sm.i2[s] | TIED_BIT) : sm.i2[s];
    if (eq > 1u) oc[pos] = 2 * i + eq;  // i + j with j = i + eq
    if (orank) orank[sm.i2[s]] = (double)(2 * i + eq + 1u) * 0.5;
  }
  // big fine buckets (heavy ties or a density spike): one warp per bucket
  const uint32_t nb = sm.nbig;
  const int lane = tid & 31, w = tid >> 5;
  for (uint32_t q = w; q < nb; q += NT / 32) {
    const int fb = sm.bigq[q];
    const uint32_t a = sm.foff[fb], e = sm.foff[fb + 1];
    double mn = sm.kd[a], mx = mn;
    for (uint32_t t = a + lane; t < e; t += 32) { const double kt = sm.kd[t]; mn = fmin(mn, kt); mx = fmax(mx, kt); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fmin(mn, __shfl_xor_sync(FULL, mn, o));
      mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
    }
    // distinct values in one fine bucket (a density spike, or the draws beyond the sample's
    // extremes): sort the bucket in place, then tie runs by binary search
    if (mn != mx) warp_sort_pairs(sm.kd + a, sm.i2 + a, e - a);
    for (uint32_t s = a + lane; s < e; s += 32) {
      const double k = sm.kd[s];
      uint32_t less = 0, eq = e - a, eqb = s - a;
      if (mn != mx) {
        uint32_t lo = a, hi = s;          // first position holding k
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sm.kd[mid] < k) lo = mid + 1; else hi = mid; }
        uint32_t lo2 = s + 1, hi2 = e;    // first position holding a larger value
        while (lo2 < hi2) { const uint32_t mid = (lo2 + hi2) >> 1; if (sm.kd[mid] <= k) lo2 = mid + 1; else hi2 = mid; }
        less = lo - a; eq = lo2 - lo; eqb = s - lo;
      }
      const uint32_t i = B0 + a + less;
      const uint32_t pos = i + eqb;
      B200REF_CHECK(pos >= B0 && pos < B0 + n, "finish_bucket (warp path): sorted position outside the bucket");
      ok[pos] = dbl2key(k);
      oi[pos] = (eq > 1u) ? (sm.i2[s] | TIED_BIT) : sm.i2[s];
      if (eq > 1u) oc[pos] = 2 * i + eq;
      if (orank) orank[sm.i2[s]] = (double)(2 * i + eq + 1u) * 0.5;
    }
  }
}


// ================================================================ wide path (long chains)
// M above ~330k pooled draws: every coarse bucket is cut into `sub` equal-width sub-buckets
// (up to 8192 buckets per parameter).  Per-chunk counts would cost as much memory as the draws,
// so the chunk CTAs add their shared-memory histograms to per-parameter totals (k_bhist_w),
// and k_partition_w reserves a run inside each bucket with one global atomic per (chunk,
// bucket).  The order inside a bucket then depends on scheduling, the RESULTS do not: k_finish
// ranks by value, and draws with equal values receive equal scores.
struct WideSmem {
  double sp[NB1_MAX];
  double fsub[NB1_MAX];
  uint32_t cnt[NBW_MAX];
  uint16_t lut[NCELL + 2];
};

__device__ __forceinline__ int wide_of(double x, double lo, double cs, const WideSmem& sm, int sub) {
  const int c = coarse_of(x, lo, cs, sm.sp, sm.lut);
  const double lc = (c == 0) ? lo : sm.sp[c - 1];
  return c * sub + sub_of(x, lc, sm.fsub[c], sub);
}

__global__ void __launch_bounds__(PART_NT) k_bhist_w(const double* __restrict__ x, uint32_t M, uint32_t nchunks,
                                                      BucketBufs b) {
  extern __shared__ __align__(16) unsigned char wide_smem_raw[];
  WideSmem& sm = *reinterpret_cast<WideSmem*>(wide_smem_raw);
  __shared__ double red[33 * 2];
  const int p = blockIdx.y;
  const uint32_t chunk = blockIdx.x;
  const BucketFn* fn = b.fn + p;
  const double* xp = x + (size_t)p * M;
  const int sub = fn->sub, nbt = fn->nb1 * sub;
  for (int i = threadIdx.x; i < NB1_MAX; i += PART_NT) { sm.sp[i] = fn->sp[i]; sm.fsub[i] = fn->fsub[i]; }
  for (int i = threadIdx.x; i < NCELL + 2; i += PART_NT) sm.lut[i] = fn->lut[i];
  for (int i = threadIdx.x; i < nbt; i += PART_NT) sm.cnt[i] = 0;
  const double lo = fn->lo, cs = fn->cs, c0 = fn->c0;
  __syncthreads();
  const uint32_t base = chunk * CHUNK_P;
  double s1 = 0.0, s2 = 0.0;
  bool bad = false;
#pragma unroll 1
  for (int it0 = 0; it0 < PART_IPT; it0 += 2) {  // pairs of draws in flight, as in k_bhist
    double v[2];
    int c[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t e = base + (it0 + j) * PART_NT + threadIdx.x;
      v[j] = (e < M) ? xp[e] : c0;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) c[j] = wide_of(v[j], lo, cs, sm, sub);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t e = base + (it0 + j) * PART_NT + threadIdx.x;
      if (e < M) {
        bad |= !is_finite_bits(v[j]);
        const double d = v[j] - c0;
        s1 += d;
        s2 = fma(d, d, s2);
        atomicAdd(&sm.cnt[c[j]], 1u);
      }
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[w] = s1; red[33 + w] = s2; }
  const int anybad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int k = 0; k < PART_NT / 32; ++k) { a += red[k]; c += red[33 + k]; }
    b.spart[((size_t)p * nchunks + chunk) * 2 + 0] = a;
    b.spart[((size_t)p * nchunks + chunk) * 2 + 1] = c;
    if (anybad) atomicOr(b.flag, 1);
  }
  uint32_t* tot = b.cur + (size_t)p * NBW_MAX;
  for (int i = threadIdx.x; i < nbt; i += PART_NT) {
    const uint32_t c = sm.cnt[i];
    if (c) atomicAdd(&tot[i], c);
  }
}

constexpr int BSW_NT = 1024;
__global__ void __launch_bounds__(BSW_NT) k_bscan_w(uint32_t M, uint32_t nchunks, BucketBufs b) {
  __shared__ uint32_t wtot[BSW_NT / 32];
  __shared__ int over;
  constexpr int PER = NBW_MAX / BSW_NT;
  const int p = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) over = 0;
  __syncthreads();
  uint32_t* cur = b.cur + (size_t)p * NBW_MAX;
  const int nbt = b.fn[p].nb1 * b.fn[p].sub;
  const int wsh = 31 - __clz(b.fn[p].sub);  // sub is a power of two
  uint32_t c[PER], tsum = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = t * PER + k;
    c[k] = (i < nbt) ? cur[i] : 0u;
    tsum += c[k];
    if (c[k] > (uint32_t)CAP1 && coarse_is_pure(b.fn[p].sp, b.fn[p].nb1, i >> wsh)) {
      // pure coarse bucket (all of it in its sub-bucket 0): any size, k_finish streams it
    } else if (c[k] > (uint32_t)CAPBIG) {
      over = 1;
    } else if (c[k] > (uint32_t)CAP1) {
      const int slot = atomicAdd(b.nbigb, 1);
      if (slot < b.biglist_cap) { b.biglist[2 * slot] = p; b.biglist[2 * slot + 1] = i; }
      else over = 1;
    }
  }
  uint32_t incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) wtot[w] = incl;
  __syncthreads();
  uint32_t pre = (lane < w) ? wtot[lane] : 0u;  // prefix over the earlier warps (at most 32 of them)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(FULL, pre, o);
  uint32_t run = pre + incl - tsum;
  uint32_t* boff = b.boff + (size_t)p * b.bstride;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = t * PER + k;
    if (i < nbt) { boff[i] = run; cur[i] = run; }
    run += c[k];
  }
  if (t == BSW_NT - 1) boff[nbt] = run;
  if (t == 0) {
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t k = 0; k < nchunks; ++k) {
      s1 += b.spart[((size_t)p * nchunks + k) * 2 + 0];
      s2 += b.spart[((size_t)p * nchunks + k) * 2 + 1];
    }
    const double dm = (double)M;
    const double m1 = s1 / dm;
    b.mean[p] = b.fn[p].c0 + m1;
    const double var = (s2 - s1 * m1) / dm;
    b.sd[p] = sqrt(var > 0.0 ? var : 0.0);
  }
  __syncthreads();
  if (t == 0 && over) {
    b.fallback[p] = 1;
    b.plist[atomicAdd(b.nfb, 1)] = p;
  }
}

__global__ void __launch_bounds__(PART_NT) k_partition_w(const double* __restrict__ x, uint32_t M, uint32_t nchunks,
                                                          BucketBufs b) {
  extern __shared__ __align__(16) unsigned char wide_smem_raw[];
  WideSmem& sm = *reinterpret_cast<WideSmem*>(wide_smem_raw);
  const int p = blockIdx.y;
  if (b.fallback[p] || (*b.flag & 1)) return;
  const uint32_t chunk = blockIdx.x;
  const BucketFn* fn = b.fn + p;
  const double* xp = x + (size_t)p * M;
  const int tid = threadIdx.x;
  const int sub = fn->sub, nbt = fn->nb1 * sub;
  for (int i = tid; i < NB1_MAX; i += PART_NT) { sm.sp[i] = fn->sp[i]; sm.fsub[i] = fn->fsub[i]; }
  for (int i = tid; i < NCELL + 2; i += PART_NT) sm.lut[i] = fn->lut[i];
  for (int i = tid; i < nbt; i += PART_NT) sm.cnt[i] = 0;
  const double lo = fn->lo, cs = fn->cs;
  __syncthreads();
  const uint32_t base = chunk * CHUNK_P;
  const uint32_t n = min((uint32_t)CHUNK_P, M - base);
  double v[PART_IPT];
  uint32_t r[PART_IPT];
  uint16_t bk[PART_IPT];
#pragma unroll
  for (int it = 0; it < PART_IPT; ++it) {
    const uint32_t e = it * PART_NT + tid;
    v[it] = (e < n) ? xp[base + e] : 0.0;
  }
#pragma unroll
  for (int it = 0; it < PART_IPT; ++it) {
    const uint32_t e = it * PART_NT + tid;
    if (e < n) {
      const int c = wide_of(v[it], lo, cs, sm, sub);
      bk[it] = (uint16_t)c;
      r[it] = atomicAdd(&sm.cnt[c], 1u);
    }
  }
  __syncthreads();
  // reserve this chunk's run in every bucket it touches: one global atomic per (chunk, bucket)
  uint32_t* cur = b.cur + (size_t)p * NBW_MAX;
  // (groups of 8 returning atomics in flight per thread: one at a time, their round trips were the kernel)
  for (int i0 = tid; i0 < nbt; i0 += 8 * PART_NT) {
    uint32_t c[8], r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + j * PART_NT;
      c[j] = (i < nbt) ? sm.cnt[i] : 0u;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = c[j] ? atomicAdd(&cur[i0 + j * PART_NT], c[j]) : 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c[j]) sm.cnt[i0 + j * PART_NT] = r[j];
  }
  __syncthreads();
  uint64_t* ok = b.pkey + (size_t)p * b.pstride;
  uint32_t* oi = b.pidx + (size_t)p * b.pstride;
#pragma unroll
  for (int it = 0; it < PART_IPT; ++it) {
    const uint32_t e = it * PART_NT + tid;
    if (e < n) {
      const uint32_t pos = sm.cnt[bk[it]] + r[it];
      B200REF_CHECK(pos < M && pos >= b.boff[(size_t)p * b.bstride + bk[it]] && pos < b.boff[(size_t)p * b.bstride + bk[it] + 1],
                    "k_partition_w: reserved run leaves its bucket");
      ok[pos] = dbl2key(v[it]);
      oi[pos] = base + e;
    }
  }
}

using FinSmemSmall = FinSmem<CAP1, NFINE>;
using FinSmemBig = FinSmem<CAPBIG, NFINE_BIG>;

// FIN_BPC consecutive buckets (each at most CAP1 draws) of one parameter per CTA: the bucket offsets
// and the flags arrive in one round trip, and while a bucket is ranked the next one's keys, positions
// and bucket function are already on their way (L1 / L2 prefetch: no registers, no staging buffer).
#ifndef B200REF_FIN_OCC
#define B200REF_FIN_OCC 6
#endif
#ifndef B200REF_FIN_BPC
#define B200REF_FIN_BPC 8
#endif
#ifndef B200REF_FIN_PF
#define B200REF_FIN_PF 1  // 0 none, 1 prefetch.global.L1, 2 prefetch.global.L2
#endif
constexpr int FIN_BPC = B200REF_FIN_BPC;

__device__ __forceinline__ void prefetch_line(const void* p) {
#if B200REF_FIN_PF == 1
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#elif B200REF_FIN_PF == 2
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// nsub / nbt: sub-buckets per coarse bucket and buckets per parameter (choose_sub(M), choose_nb1(M) * nsub:
// what k_sample stored in fn), computed once on the host.
__global__ void __launch_bounds__(FIN_NT, B200REF_FIN_OCC) k_finish(uint32_t M, int nsub, int nbt, BucketBufs b, FinishOut fo) {
  extern __shared__ __align__(16) unsigned char fin_smem_raw[];
  FinSmemSmall& sm = *reinterpret_cast<FinSmemSmall*>(fin_smem_raw);
  __shared__ uint32_t sbo[FIN_BPC + 2];
  __shared__ uint32_t sso[FIN_BPC + 2];  // where each bucket's partitioned records start (slabs: fn->soff)
  const int p = blockIdx.y, tid = threadIdx.x;
  const int cb0 = blockIdx.x * FIN_BPC;
  if (cb0 >= nbt) return;
  const BucketFn* fn = b.fn + p;
  const uint8_t fb = b.fallback[p];
  if (tid <= FIN_BPC) {
    const uint32_t bo = b.boff[(size_t)p * b.bstride + min(cb0 + tid, nbt)];
    sbo[tid] = bo;
    sso[tid] = b.slab ? fn->soff[min(cb0 + tid, nbt)] : bo;
  }
  __syncthreads();
  if (fb) return;
  const uint32_t* pi0 = b.pidx + (size_t)p * b.pstride;
  const uint64_t* pk0 = b.pkey + (size_t)p * b.pstride;
  const int sh = 31 - __clz(nsub);
#pragma unroll 1
  for (int q = 0; q < FIN_BPC; ++q) {
    const int cb = cb0 + q;
    if (cb >= nbt) break;
    const uint32_t B0 = sbo[q], n = sbo[q + 1] - B0, S0 = sso[q];
    const FineFn ff = make_fine<NFINE>(fn, cb, nsub);
#if B200REF_FIN_PF != 0
    if (tid >= FIN_NT - 32 && q + 1 < FIN_BPC && cb + 1 < nbt) {
      // the last warp (the one most often idle in a bucket's tail) issues one prefetch per 128-byte line of
      // the next bucket's keys and positions, and of its bucket function
      const uint32_t lane = tid & 31;
      const uint32_t nS0 = sso[q + 1], nn = min(sbo[q + 2] - sbo[q + 1], (uint32_t)CAP1);
      const uintptr_t ka = (uintptr_t)(pk0 + nS0), ke = ka + (size_t)nn * 8;
      const uintptr_t ia = (uintptr_t)(pi0 + nS0), ie = ia + (size_t)nn * 4;
      const uint32_t nlk = nn ? (uint32_t)(((ke - 1) >> 7) - (ka >> 7)) + 1u : 0u;
      const uint32_t nli = nn ? (uint32_t)(((ie - 1) >> 7) - (ia >> 7)) + 1u : 0u;
      for (uint32_t l = lane; l < nlk + nli; l += 32)
        prefetch_line((const void*)(l < nlk ? ((ka >> 7) + l) << 7 : ((ia >> 7) + (l - nlk)) << 7));
      const int c1 = (cb + 1) >> sh;
      if (lane == 0) prefetch_line(&fn->fsub[c1]);
      if (lane == 1 && c1 > 0) prefetch_line(&fn->sp[c1 - 1]);
    }
#endif
    bool wanted = true;
    if (fo.npos) {  // uniform: statistics-only pass
      wanted = false;
#pragma unroll
      for (int i = 0; i < 2 * 8 + 2; ++i)  // static indices: the parameter struct stays in the constant bank
        wanted |= (i < fo.npos) && (fo.pos[i] - B0 < n);  // unsigned: B0 <= pos < B0 + n
    }
    if (wanted && n > (uint32_t)CAP1 && coarse_is_pure(fn->sp, fn->nb1, cb >> sh)) {
      // copies of one value (see coarse_is_pure): one tie run [B0, B0 + n), nothing to rank.  (Pure buckets that
      // fit on-chip take the general path below, which handles them just as well: the purity test costs a
      // dependent global read per bucket, so it is only made where it matters.)
      const uint32_t kc = 2u * B0 + n;
      uint64_t* ok = b.skey + (size_t)p * M + B0;
      uint32_t* oi = b.sidx + (size_t)p * M + B0;
      uint32_t* oc = b.kcode + (size_t)p * M + B0;
      double* orank = fo.rank ? fo.rank + (size_t)p * M : nullptr;
      for (uint32_t e = tid; e < n; e += FIN_NT) {
        const uint32_t pos = pi0[S0 + e];
        ok[e] = pk0[S0 + e];
        oi[e] = pos | TIED_BIT;
        oc[e] = kc;
        if (orank) orank[pos] = (double)(kc + 1u) * 0.5;
      }
    } else if (wanted && n != 0 && n <= (uint32_t)CAP1) {  // oversize buckets: k_finish_big
        finish_bucket<FIN_NT, NFINE, CAP1, FIN_IPT, false, false>(sm, ff, B0, S0, n, M, p, b, fo);
      __syncthreads();
    }
  }
}

// oversize buckets (CAP1 < n <= CAPBIG: a few heavily repeated values): a small grid
// strides over the list k_bscan built
__global__ void __launch_bounds__(FINB_NT) k_finish_big(uint32_t M, BucketBufs b, FinishOut fo) {
  extern __shared__ __align__(16) unsigned char fin_smem_raw[];
  FinSmemBig& sm = *reinterpret_cast<FinSmemBig*>(fin_smem_raw);
  // k_bscan / k_bscan_w keep counting past the end of the list (and send those parameters to the LSD
  // fallback): only the entries that were actually stored may be read
  const int nbig = min(*b.nbigb, b.biglist_cap);
  const int nsub = choose_sub(M);
  for (int q = blockIdx.x; q < nbig; q += gridDim.x) {
    const int p = b.biglist[2 * q], cb = b.biglist[2 * q + 1];
    if (b.fallback[p]) continue;
    const uint32_t B0 = b.boff[(size_t)p * b.bstride + cb];
    const uint32_t n = b.boff[(size_t)p * b.bstride + cb + 1] - B0;
    const uint32_t S0 = b.slab ? b.fn[p].soff[cb] : B0;
    B200REF_CHECK(q < b.biglist_cap && n <= (uint32_t)CAPBIG && B0 + n <= M, "k_finish_big: list entry out of range");
    __syncthreads();
    finish_bucket<FINB_NT, NFINE_BIG, CAPBIG, 1, true, true>(sm, make_fine<NFINE_BIG>(b.fn + p, cb, nsub), B0, S0, n, M, p, b, fo);
  }
}

}  // namespace b200ref
