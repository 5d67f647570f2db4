// Per-parameter chain statistics on the normal scores (original positions):
//   split-chain moments -> R-hat            (_split_chains :76-85, _rhat :136-151,
//                                             _variance :196-201 of diagnostics.py)
//   unsplit-chain moments + autocovariance
//   with first-negative truncation -> ESS   (_ess :154-177, _autocorr :180-193)
//
// Streaming design.  zz[p][c][i] = {z_bulk, z_folded}.  Every chain is cut into units
// (runs of <= 8 sub-tiles of one chain); a CTA streams its units through
// shared memory in 2048-draw sub-tiles and accumulates, for both kinds at once,
//   * per half of the chain: sum, sum of squares and "some draw differs from the first" (raw
//     moments: z-scores have mean ~ 0 and variance ~ 1, so sum(y^2) - n*mean^2 loses no accuracy;
//     a constant half -> variance 0 exactly)
//   * 16 raw lag products R_t = sum_i y_i * y_{i+t}, t = t0 .. t0+15, by a register
//     sliding window (16 FMAs per 2 shared loads).
// A small per-(parameter, kind) kernel turns the partial sums into chain means / variances,
// centres the lag products algebraically
//     S_t = R_t - mean * (2*total - head_t - tail_t) + (n - t) * mean^2 ,
// and walks rho_1, rho_2, ... until the first negative one (diagnostics.py:172-176).
// Pairs whose rho is still positive after 16 lags go onto a work list; further rounds
// stream only those chains, 32 lags per round.  Well-mixed chains finish in round 0,
// and no chain length limit remains (nothing has to fit on-chip).
// Because all chains of a parameter have the same length n, sum_c cov_c / (n - t) is
// accumulated as (sum_c S_c) / (n - t).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace b200ref {

constexpr int MO_NT = 128;               // threads working on one kind / one lag block
constexpr int MO_CTA = 2 * MO_NT;        // threads per moments CTA
constexpr int LAGB = 16;                 // lags per thread block half; round 0 covers lags 1..16
constexpr int LAGR = 2 * LAGB;           // lags per later round (the two CTA halves take 16 each)
constexpr int CHUNK = 16;                // consecutive draws per thread step
constexpr int MO_SUB = MO_NT * CHUNK;    // 2048 draws per sub-tile
#ifndef B200REF_MO_UNIT_TILES
#define B200REF_MO_UNIT_TILES 8
#endif
constexpr int MO_UNIT_TILES = B200REF_MO_UNIT_TILES;  // a unit (one CTA) covers at most this many sub-tiles
#ifndef B200REF_MO_CTA_TILES
#define B200REF_MO_CTA_TILES 12
#endif
constexpr int MO_CTA_TILES = B200REF_MO_CTA_TILES;  // sub-tiles one round-0 CTA streams (several units)
constexpr int MO_VALS0 = LAGB + 10;      // per (unit, kind) in round 0: 16 lag sums + 2 x {sum, sq, first, differs, present}
constexpr int LAG0X = 2 * LAGB;          // + lags 17 .. 48 for the parameters k_predict_sticky marks (tile variant only)
constexpr int MO_VALS = MO_VALS0 + LAG0X;  // stride of a (unit, kind) record: [0,16) lags 1..16, [16,26) moments, [26,58) lags 17..48
constexpr int DEC_NT = 64;

// Chains are cut into units of whole 2048-draw sub-tiles, independent of the split point: the lag
// products run over the unsplit chain (diagnostics.py:55-56), only sum / sum of squares are kept
// per half (split R-hat, diagnostics.py:76-85).
struct ChainGeom {
  uint32_t C, N, half, ntiles, tpu, nunits;  // tpu = sub-tiles per unit
};

__host__ inline ChainGeom chain_geom(int64_t C, int64_t N) {
  ChainGeom g;
  g.C = (uint32_t)C;
  g.N = (uint32_t)N;
  g.half = g.N / 2;
  g.ntiles = (g.N + MO_SUB - 1) / MO_SUB;
  const uint32_t nu = (g.ntiles + MO_UNIT_TILES - 1) / MO_UNIT_TILES;
  g.tpu = (g.ntiles + nu - 1) / nu;
  g.nunits = (g.ntiles + g.tpu - 1) / g.tpu;
  return g;
}

__device__ __forceinline__ void unit_range(const ChainGeom& g, uint32_t u, uint32_t& a, uint32_t& b) {
  a = min(g.N, u * g.tpu * (uint32_t)MO_SUB);
  b = min(g.N, (u + 1) * g.tpu * (uint32_t)MO_SUB);
}

// Shared-memory tiles hold 16-draw groups padded to 18 doubles: a thread's 16 consecutive draws
// start 144 B apart, so its 16-byte LDS.128 reads are bank-conflict free and stay 16-byte aligned.
__host__ __device__ __forceinline__ uint32_t pad_idx(uint32_t e) { return e + 2u * (e >> 4); }
constexpr int MO_HALO0 = LAGB;                      // round 0 reads 16 draws past the sub-tile
constexpr int MO_HALOR = LAGR;                      // later rounds: 32 (second half's window)
constexpr int MO_RAW0 = MO_SUB + MO_HALO0;          // double2 pairs staged per sub-tile by TMA
constexpr int MO_RAWR = MO_SUB + MO_HALOR;
constexpr int MO_PAD0 = (MO_RAW0 / CHUNK + 1) * (CHUNK + 2);
constexpr int MO_PADR = (MO_RAWR / CHUNK + 1) * (CHUNK + 2);

// v[16] per lane -> lane l returns the warp total of v[l & 15]  (16 shuffles)
__device__ __forceinline__ double warp_reduce_16(double (&v)[16]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; ++k) {
      const double a = v[k], b = v[k + s];
      const double keep = up ? b : a;
      const double send = up ? a : b;
      v[k] = keep + __shfl_xor_sync(FULL, send, s);
    }
  }
  return v[0] + __shfl_xor_sync(FULL, v[0], 16);
}

struct ChainState {
  double* part;      // [PB][C*nunits][2][MO_VALS] round-0 partial sums
  uint8_t* sticky;   // [PB] 1: round 0 also covers lags 17 .. 48 of this parameter (k_predict_sticky)
  double* partr;     // [nwork][C*nunits][LAGR]   later rounds
  double* cmean;     // [2*PB][C]  unsplit chain means
  double* csum;      // [2*PB][C]  chain totals
  double* head;      // [2*PB][C]  sum of the first (t0 - 1) draws of the chain
  double* tail;      // [2*PB][C]  sum of the last  (t0 - 1) draws
  double* varhat;    // [2*PB]
  double* rhosum;    // [2*PB]
  double* gpend;     // [2*PB] Geyer mode: first correlation of the open pair (NaN: none pending)
  double* gprev;     // [2*PB] Geyer mode: previous (capped) pair sum
  int geyer;         // 0 = the reference's first-negative truncation, 1 = Geyer initial monotone sequence (extension)
  int* wl[2];        // work lists of unfinished items (2*p + kind)
  int* nwl;          // [2] their lengths
  double* rhat;      // [2][PB] outputs, kind-major
  double* ess;       // [2][PB]
  int64_t* lag;      // [2][PB] first lag with rho < 0 (N if none)
};

// 16 consecutive doubles of a padded tile starting at group base gb (16-byte aligned) -> registers
__device__ __forceinline__ void load16(const double* __restrict__ gb, double* __restrict__ u) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double2 t = *reinterpret_cast<const double2*>(gb + 2 * k);
    u[2 * k] = t.x;
    u[2 * k + 1] = t.y;
  }
}

// One elected thread: TMA bulk copy of `npairs` {z_bulk, z_folded} pairs into the staging buffer.
__device__ __forceinline__ void stage_pairs(double2* raw, const double2* src, uint32_t npairs, uint64_t* bar) {
  fence_proxy_async();  // earlier generic reads of `raw` are ordered before the async-proxy writes
  mbar_expect_tx(bar, npairs * 16u);
  bulk_g2s(raw, src, npairs * 16u, bar);
}

// ---------------------------------------------------------------- round 0: lags 1..16 + moments
// grid (C * nunits, PB), 256 threads: one CTA streams one unit of one chain for BOTH kinds.  Each
// sub-tile of 2048 + 16 pairs arrives by one TMA bulk copy (UBLKCP) into a staging buffer and is
// split into two padded fp64 tiles (double-buffered: one barrier per sub-tile); the next copy is in
// flight while warps 0-3 accumulate z_bulk and warps 4-7 z_folded: 16 raw lag products
// R_t = sum_i y_i y_{i+t}, and per half of the chain the sum, the sum of squares and "some draw
// differs from the first" (a constant half must give variance 0 exactly).
struct Mo0Smem {
  double2 raw[MO_RAW0];
  double y[2][2][MO_PAD0];  // [buffer][kind]
  double red[2][MO_NT / 32][MO_VALS0];
  uint64_t bar;
};

// upc consecutive units of one parameter per CTA (units of a parameter are adjacent in memory): the copy of
// a unit's first sub-tile is issued while the previous unit's last sub-tile is still being accumulated, so
// only the CTA's very first copy is exposed (a chain of 4096 draws is just two sub-tiles).
__global__ void __launch_bounds__(MO_CTA, 2) k_moments0(const double2* __restrict__ zz, ChainGeom g, uint32_t upc, double* part) {
  extern __shared__ __align__(128) unsigned char mo_smem_raw[];
  Mo0Smem& S = *reinterpret_cast<Mo0Smem*>(mo_smem_raw);
  const int p = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  const int kind = tid >> 7, ct = tid & (MO_NT - 1), wp = ct >> 5;
  const uint32_t unit0 = blockIdx.x * upc, unit1 = min(unit0 + upc, g.C * g.nunits);
  const double2* zp = zz + (size_t)p * g.C * g.N;
  if (tid == 0) {
    mbar_init(&S.bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // every unit is non-empty: its first sub-tile starts below N (chain_geom)
  auto unit_start = [&](uint32_t unit, uint32_t& c, uint32_t& a, uint32_t& b) {
    c = unit / g.nunits;
    unit_range(g, unit % g.nunits, a, b);
  };
  if (tid == 0 && unit0 < unit1) {
    uint32_t c, a, b;
    unit_start(unit0, c, a, b);
    stage_pairs(S.raw, zp + (size_t)c * g.N + a, min((uint32_t)MO_RAW0, g.N - a), &S.bar);
  }
  const uint32_t half = g.half, two = 2 * g.half;
  uint32_t phase = 0, buf = 0;
  for (uint32_t unit = unit0; unit < unit1; ++unit) {
    uint32_t c, a, b;
    unit_start(unit, c, a, b);
    const double2* zc = zp + (size_t)c * g.N;
    // first draw of this unit in each half (if any): read from the staged tiles when they arrive
    const bool has1 = a < half, has2 = b > half && max(a, half) < two;
    const uint32_t pos2 = max(a, half);
    double f1 = 0.0, f2 = 0.0;
    double acc[LAGB];
#pragma unroll
    for (int l = 0; l < LAGB; ++l) acc[l] = 0.0;
    double s1 = 0.0, q1 = 0.0, s2 = 0.0, q2 = 0.0;
    bool d1 = false, d2 = false;
    for (uint32_t sub = a; sub < b; sub += MO_SUB, buf ^= 1u) {
      const uint32_t len = min((uint32_t)MO_SUB, b - sub);
      const uint32_t nel = min((uint32_t)MO_RAW0, g.N - sub);  // pairs the copy delivered; the rest reads as 0
      mbar_wait(&S.bar, phase);
      phase ^= 1u;
      double* y0 = S.y[buf][0];
      double* y1 = S.y[buf][1];
#pragma unroll
      for (int it = 0; it < MO_SUB / MO_CTA; ++it) {
        const uint32_t i = it * MO_CTA + tid;
        const double2 v = (i < nel) ? S.raw[i] : make_double2(0.0, 0.0);
        y0[pad_idx(i)] = v.x;
        y1[pad_idx(i)] = v.y;
      }
      if (tid < MO_HALO0) {
        const uint32_t i = MO_SUB + tid;
        const double2 v = (i < nel) ? S.raw[i] : make_double2(0.0, 0.0);
        y0[pad_idx(i)] = v.x;
        y1[pad_idx(i)] = v.y;
      }
      __syncthreads();  // tiles complete, staging buffer free; buffer buf^1 was last read before this barrier
      if (tid == 0) {
        if (sub + MO_SUB < b) {
          stage_pairs(S.raw, zc + sub + MO_SUB, min((uint32_t)MO_RAW0, g.N - (sub + MO_SUB)), &S.bar);
        } else if (unit + 1 < unit1) {
          uint32_t cn, an, bn;
          unit_start(unit + 1, cn, an, bn);
          stage_pairs(S.raw, zp + (size_t)cn * g.N + an, min((uint32_t)MO_RAW0, g.N - an), &S.bar);
        }
      }
      if (sub == a && has1) f1 = S.y[buf][kind][0];
      if (has2 && pos2 >= sub && pos2 < sub + len) f2 = S.y[buf][kind][pad_idx(pos2 - sub)];
      const uint32_t i0 = ct * CHUNK;
      if (i0 < len) {
        const uint32_t nvalid = min((uint32_t)CHUNK, len - i0);
        const uint32_t g0 = sub + i0;
        double w[2 * CHUNK];
        const double* gb = &S.y[buf][kind][ct * (CHUNK + 2)];
        load16(gb, w);
        load16(gb + CHUNK + 2, w + CHUNK);
        const bool in1 = g0 + CHUNK <= half, in2 = g0 >= half && g0 + CHUNK <= two;
        if (nvalid == CHUNK && (in1 || in2)) {
          // "differs from the half's first draw" on the bit patterns (integer pipe): scores are table entries,
          // equal values are equal bits (the table's only zero is +0.0)
          const long long frb = __double_as_longlong(in1 ? f1 : f2);
          double ls = 0.0, lq = 0.0;
          bool ld = false;
#pragma unroll
          for (int s = 0; s < CHUNK; ++s) {
            const double v = w[s];
            ls += v;
            lq = fma(v, v, lq);
            ld |= (__double_as_longlong(v) != frb);
#pragma unroll
            for (int l = 0; l < LAGB; ++l) acc[l] = fma(v, w[s + 1 + l], acc[l]);
          }
          if (in1) { s1 += ls; q1 += lq; d1 |= ld; } else { s2 += ls; q2 += lq; d2 |= ld; }
        } else {
#pragma unroll
          for (int s = 0; s < CHUNK; ++s) {
            const bool ok = (uint32_t)s < nvalid;
            const double v = ok ? w[s] : 0.0;
            const uint32_t gi = g0 + s;
            if (ok && gi < half) { s1 += v; q1 = fma(v, v, q1); d1 |= (v != f1); }
            else if (ok && gi < two) { s2 += v; q2 = fma(v, v, q2); d2 |= (v != f2); }
#pragma unroll
            for (int l = 0; l < LAGB; ++l) acc[l] = fma(v, w[s + 1 + l], acc[l]);
          }
        }
      }
    }
    // unit totals: 16 lag sums by the halving reduce, the half statistics by butterflies
    const double tot = warp_reduce_16(*reinterpret_cast<double (*)[16]>(acc));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(FULL, s1, o);
      q1 += __shfl_xor_sync(FULL, q1, o);
      s2 += __shfl_xor_sync(FULL, s2, o);
      q2 += __shfl_xor_sync(FULL, q2, o);
    }
    const bool w1 = __any_sync(FULL, d1), w2 = __any_sync(FULL, d2);
    double* r = S.red[kind][wp];
    if (lane < LAGB) r[lane] = tot;
    if (lane == 0) {
      r[16] = s1; r[17] = q1; r[18] = f1; r[19] = w1 ? 1.0 : 0.0; r[20] = has1 ? 1.0 : 0.0;
      r[21] = s2; r[22] = q2; r[23] = f2; r[24] = w2 ? 1.0 : 0.0; r[25] = has2 ? 1.0 : 0.0;
    }
    __syncthreads();
    if (ct < MO_VALS0) {
      double v = S.red[kind][0][ct];
      const bool copy = (ct == 18 || ct == 20 || ct == 23 || ct == 25);  // first values / presence: not summed
      if (!copy) {
#pragma unroll
        for (int k = 1; k < MO_NT / 32; ++k) v += S.red[kind][k][ct];
      }
      part[(((size_t)p * g.C * g.nunits + unit) * 2 + kind) * MO_VALS + ct] = v;
    }
    // (S.red is next written after the next unit's tile barriers)
  }
}

// ---------------------------------------------------------------- round 0, tile variant (N % 8 == 0)
// The same partial sums as k_moments0, without the de-interleave pass, without block barriers and with a quarter of
// the reductions.  zz is viewed as a 2-D fp64 tensor of rows of 8 {z_bulk, z_folded} pairs (128 B).  Every WARP owns
// whole units (runs of one chain): a tile of 32 rows + 2 halo rows arrives by one 2-D TMA copy (cp.async.bulk.tensor,
// SASS UTMALDG) in the 128-byte swizzle, so that lane t's 16-byte reads of row t (and of rows t + 1, t + 2 for the lag
// partners) are bank-conflict free straight from the landing buffer: no staging copy.  A lane keeps its 8 draws of
// BOTH kinds in registers and accumulates 2 x 16 lag products against the 24 draws that follow (256 DFMA per
// 24 LDS.128).  Three landing buffers per warp, two copies always in flight; the warps of a CTA never synchronise
// with each other, and a unit's totals are reduced inside its warp only.  Units, partial-sum layout and the decision
// kernel are those of k_moments0.
constexpr int M2_NT = 128;                       // threads per CTA (4 independent warps)
constexpr int M2_NW = M2_NT / 32;
constexpr int M2_ROWS = 32;                      // rows per tile = lanes
constexpr int M2_BOX = M2_ROWS + 6;              // + halo rows: the 16 draws after the tile's last one (lags 1 .. 16), 48 for lags up to 48
constexpr int M2_TILE_BYTES = M2_BOX * 128;
constexpr int M2_STAGE = ((M2_TILE_BYTES + 1023) / 1024) * 1024;  // landing buffers are 1024-byte aligned (swizzle atom)
#ifndef B200REF_M2_NST
#define B200REF_M2_NST 3
#endif
constexpr int M2_NST = B200REF_M2_NST;
#ifndef B200REF_M2_OCC
#define B200REF_M2_OCC 3
#endif
constexpr size_t M2_SMEM = 1024 + (size_t)M2_NW * M2_NST * M2_STAGE + (size_t)M2_NW * M2_NST * 8;

// pair `k` (0..7) of row `r` of a landing buffer (128-byte swizzle: 16-byte chunk index ^ (row & 7)).
// (`tile` is the buffer's 32-bit shared-window address: explicit ld.shared, the rounded-up base pointer is generic)
__device__ __forceinline__ double2 m2_pair(uint32_t tile, uint32_t r, uint32_t k) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(tile + r * 128u + ((k ^ (r & 7u)) << 4)));
  return v;
}

// v[8] per lane -> lane l returns the warp total of v[l & 7]
__device__ __forceinline__ double warp_reduce_8(double (&v)[8]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 4; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; ++k) {
      const double a = v[k], b = v[k + s];
      const double keep = up ? b : a;
      const double send = up ? a : b;
      v[k] = keep + __shfl_xor_sync(FULL, send, s);
    }
  }
  double t = v[0] + __shfl_xor_sync(FULL, v[0], 8);
  return t + __shfl_xor_sync(FULL, t, 16);
}

// One warp per parameter looks at the first draws of its first chain: is the lag-8 autocorrelation of either score
// far above its sampling error?  Then most of this parameter's chains will still be correlated after the 16 lags of
// round 0, and round 0 covers 48 lags for it in the same (tile) kernel instead of leaving lags 17 .. 48 to the
// slower later rounds.  Only the COST depends on the guess: lag sums are formed in the same way wherever they are
// computed, and a parameter that is guessed wrong just takes the other route.
constexpr uint32_t STICKY_DRAWS = 1024;
__global__ void __launch_bounds__(128) k_predict_sticky(const double2* __restrict__ zz, ChainGeom g, int PB, uint8_t* __restrict__ sticky) {
  const int p = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= PB) return;
  const double2* z = zz + (size_t)p * g.C * g.N;  // chain 0
  const uint32_t L = min(g.N, STICKY_DRAWS);
  uint8_t out = 0;
  if (g.N >= 128u) {
    double s[2] = {0.0, 0.0}, q[2] = {0.0, 0.0}, c8[2] = {0.0, 0.0};
    for (uint32_t i = lane; i < L; i += 32) {
      const double2 a = z[i];
      s[0] += a.x; q[0] = fma(a.x, a.x, q[0]);
      s[1] += a.y; q[1] = fma(a.y, a.y, q[1]);
      if (i + 8 < L) {
        const double2 b = z[i + 8];
        c8[0] = fma(a.x, b.x, c8[0]);
        c8[1] = fma(a.y, b.y, c8[1]);
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double ss = warp_sum(s[k]), qq = warp_sum(q[k]), cc = warp_sum(c8[k]);
      const double m = ss / (double)L, var = qq / (double)L - m * m;
      const double rho8 = (var > 0.0) ? (cc / (double)(L - 8) - m * m) / var : 0.0;
      if (rho8 > 0.15) out = 1;
    }
  }
  if (lane == 0) sticky[p] = out;
}

__global__ void __launch_bounds__(M2_NT, B200REF_M2_OCC) k_moments0t(const __grid_constant__ CUtensorMap map, ChainGeom g, uint32_t upc,
                                                                     double* part, const uint8_t* __restrict__ sticky) {
  extern __shared__ __align__(128) unsigned char mo_smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(mo_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int p = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  unsigned char* wbase = base + (size_t)wp * M2_NST * M2_STAGE;              // this warp's landing buffers
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)M2_NW * M2_NST * M2_STAGE) + wp * M2_NST;
  const uint32_t wbase_s = smem_u32(wbase);
  const uint32_t nu = g.C * g.nunits;
  const uint32_t unit0 = blockIdx.x * upc + wp, unit1 = min(blockIdx.x * upc + upc, nu);  // this warp: unit0, unit0 + 4, ...
  const uint32_t Rc = g.N >> 3;                          // rows per chain
  const uint32_t prow0 = (uint32_t)p * g.C * Rc;         // first row of this parameter
  // passes over a unit: lags 1 .. 16 (+ the moments); for a parameter marked sticky also lags 17 .. 32 and 33 .. 48
  // (the unit's tiles come again, from L2: the same loop with the partner rows 2 and 4 rows further on)
  const uint32_t nb = sticky[p] ? 3u : 1u;
  if (lane == 0) {
    for (int s = 0; s < M2_NST; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  // rows [ra, rb) of chain c for a unit
  auto unit_rows = [&](uint32_t unit, uint32_t& c, uint32_t& ra, uint32_t& rb) {
    c = unit / g.nunits;
    uint32_t a, b;
    unit_range(g, unit % g.nunits, a, b);
    ra = a >> 3;
    rb = b >> 3;
  };
  // producer cursor (lane 0): the next tile to request
  uint32_t iu = unit0, ib = 0, it_ = 0, issued = 0;
  auto issue_next = [&]() {  // lane 0 only
    while (iu < unit1) {
      uint32_t c, ra, rb;
      unit_rows(iu, c, ra, rb);
      const uint32_t nt = (rb - ra + M2_ROWS - 1) / M2_ROWS;
      if (it_ < nt) {
        const uint32_t st = issued % M2_NST;
        fence_proxy_async();  // earlier generic reads of this landing buffer are ordered before the async-proxy writes
        mbar_expect_tx(&full[st], (uint32_t)M2_TILE_BYTES);
        tma_load_2d(wbase + (size_t)st * M2_STAGE, &map, &full[st], 0, (int)(prow0 + c * Rc + ra + it_ * M2_ROWS));
        ++it_;
        ++issued;
        return;
      }
      it_ = 0;
      if (++ib == nb) { ib = 0; iu += M2_NW; }
    }
  };
  if (lane == 0)
    for (int s = 0; s < M2_NST; ++s) issue_next();
  const uint32_t half = g.half;
  uint32_t seq = 0;  // tiles consumed so far by this warp
  // 16 lag products of this lane's row (8 draws, both kinds) against the rows o, o + 1, o + 2 after it: lags
  // 8 o + 1 .. 8 o + 16.  FIRST (o = 0): the first partner row is the lane's own row, already in registers.
  auto lag_block = [&](auto first_tag, uint32_t tile, uint32_t rc, uint32_t o, const double2 (&v)[8], double (&accb)[LAGB],
                       double (&accf)[LAGB]) {
    constexpr bool FIRST = decltype(first_tag)::value;
    if (FIRST || rc + o < Rc) {
#pragma unroll
      for (int j = 1; j < 8; ++j) {
        const double2 w = FIRST ? v[j] : m2_pair(tile, lane + o, j);
#pragma unroll
        for (int s = 0; s < j; ++s) {
          accb[j - s - 1] = fma(v[s].x, w.x, accb[j - s - 1]);
          accf[j - s - 1] = fma(v[s].y, w.y, accf[j - s - 1]);
        }
      }
    }
    if (rc + o + 1 < Rc) {  // next row (same chain)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double2 w = m2_pair(tile, lane + o + 1, j);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          accb[8 + j - s - 1] = fma(v[s].x, w.x, accb[8 + j - s - 1]);
          accf[8 + j - s - 1] = fma(v[s].y, w.y, accf[8 + j - s - 1]);
        }
      }
    }
    if (rc + o + 2 < Rc) {  // the row after
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double2 w = m2_pair(tile, lane + o + 2, j);
#pragma unroll
        for (int s = j; s < 8; ++s) {
          accb[16 + j - s - 1] = fma(v[s].x, w.x, accb[16 + j - s - 1]);
          accf[16 + j - s - 1] = fma(v[s].y, w.y, accf[16 + j - s - 1]);
        }
      }
    }
  };
  for (uint32_t unit = unit0; unit < unit1; unit += M2_NW) {
    uint32_t c, ra, rb;
    unit_rows(unit, c, ra, rb);
    const uint32_t a = ra << 3, b = rb << 3;
    const uint32_t ntile = (rb - ra + M2_ROWS - 1) / M2_ROWS;
    double* pb = part + (((size_t)p * nu + unit) * 2 + 0) * MO_VALS;
    double* pf = pb + MO_VALS;
    // ---- pass 0: lags 1 .. 16 and the moments
    {
      const bool has1 = a < half, has2 = b > half && max(a, half) < 2 * half;
      const uint32_t pos2 = max(a, half);
      double accb[LAGB], accf[LAGB];
#pragma unroll
      for (int l = 0; l < LAGB; ++l) { accb[l] = 0.0; accf[l] = 0.0; }
      double hs[8];  // s1b q1b s2b q2b s1f q1f s2f q2f
#pragma unroll
      for (int l = 0; l < 8; ++l) hs[l] = 0.0;
      double2 f1 = make_double2(0.0, 0.0), f2 = make_double2(0.0, 0.0);
      bool d1b = false, d2b = false, d1f = false, d2f = false;
      for (uint32_t ti = 0; ti < ntile; ++ti, ++seq) {
        const uint32_t st = seq % M2_NST;
        const uint32_t tile = wbase_s + st * (uint32_t)M2_STAGE;
        mbar_wait(&full[st], (seq / M2_NST) & 1u);
        const uint32_t row0 = ra + ti * M2_ROWS;  // chain row of the tile's first row
        if (ti == 0 && has1) f1 = m2_pair(tile, 0, 0);
        if (has2 && (pos2 >> 3) >= row0 && (pos2 >> 3) < row0 + M2_ROWS) f2 = m2_pair(tile, (pos2 >> 3) - row0, pos2 & 7u);
        const uint32_t rc = row0 + lane;          // this lane's row of the chain
        if (rc < rb) {
          double2 v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = m2_pair(tile, lane, k);
          lag_block(std::true_type{}, tile, rc, 0u, v, accb, accf);
          // per-half sums, sums of squares and "some draw differs from the half's first one" (bit patterns: scores are
          // table entries, equal values are equal bits)
          const uint32_t g0 = rc << 3;
          const bool in1 = g0 + 8 <= half, in2 = g0 >= half && g0 + 8 <= 2 * half;
          if (in1 || in2) {
            const long long rb_ = __double_as_longlong(in1 ? f1.x : f2.x), rf_ = __double_as_longlong(in1 ? f1.y : f2.y);
            double lsb = 0.0, lqb = 0.0, lsf = 0.0, lqf = 0.0;
            bool ldb = false, ldf = false;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
              lsb += v[s].x; lqb = fma(v[s].x, v[s].x, lqb); ldb |= (__double_as_longlong(v[s].x) != rb_);
              lsf += v[s].y; lqf = fma(v[s].y, v[s].y, lqf); ldf |= (__double_as_longlong(v[s].y) != rf_);
            }
            if (in1) { hs[0] += lsb; hs[1] += lqb; d1b |= ldb; hs[4] += lsf; hs[5] += lqf; d1f |= ldf; }
            else     { hs[2] += lsb; hs[3] += lqb; d2b |= ldb; hs[6] += lsf; hs[7] += lqf; d2f |= ldf; }
          } else {
#pragma unroll
            for (int s = 0; s < 8; ++s) {
              const uint32_t gi = g0 + s;
              if (gi < half) {
                hs[0] += v[s].x; hs[1] = fma(v[s].x, v[s].x, hs[1]); d1b |= (v[s].x != f1.x);
                hs[4] += v[s].y; hs[5] = fma(v[s].y, v[s].y, hs[5]); d1f |= (v[s].y != f1.y);
              } else if (gi < 2 * half) {
                hs[2] += v[s].x; hs[3] = fma(v[s].x, v[s].x, hs[3]); d2b |= (v[s].x != f2.x);
                hs[6] += v[s].y; hs[7] = fma(v[s].y, v[s].y, hs[7]); d2f |= (v[s].y != f2.y);
              }
            }
          }
        }
        __syncwarp();  // every lane has read the landing buffer: it can take the next copy
        if (lane == 0) issue_next();
      }
      // unit totals (fixed order): halving reduces inside the warp, written straight to the unit's partial sums
      const double totb = warp_reduce_16(accb);
      const double totf = warp_reduce_16(accf);
      const double toth = warp_reduce_8(hs);
      const bool w1b = __any_sync(FULL, d1b), w2b = __any_sync(FULL, d2b), w1f = __any_sync(FULL, d1f), w2f = __any_sync(FULL, d2f);
      if (lane < LAGB) { pb[lane] = totb; pf[lane] = totf; }
      // lane l (< 8) holds hs[l]: s1b q1b s2b q2b | s1f q1f s2f q2f
      if (lane < 4) pb[16 + (lane >> 1) * 5 + (lane & 1)] = toth;
      else if (lane < 8) pf[16 + ((lane - 4) >> 1) * 5 + (lane & 1)] = toth;
      if (lane == 8) {
        pb[18] = f1.x; pb[19] = w1b ? 1.0 : 0.0; pb[20] = has1 ? 1.0 : 0.0;
        pb[23] = f2.x; pb[24] = w2b ? 1.0 : 0.0; pb[25] = has2 ? 1.0 : 0.0;
      }
      if (lane == 9) {
        pf[18] = f1.y; pf[19] = w1f ? 1.0 : 0.0; pf[20] = has1 ? 1.0 : 0.0;
        pf[23] = f2.y; pf[24] = w2f ? 1.0 : 0.0; pf[25] = has2 ? 1.0 : 0.0;
      }
    }
    // ---- passes 1, 2 (sticky parameters): lags 17 .. 32, 33 .. 48
    for (uint32_t blk = 1; blk < nb; ++blk) {
      double accb[LAGB], accf[LAGB];
#pragma unroll
      for (int l = 0; l < LAGB; ++l) { accb[l] = 0.0; accf[l] = 0.0; }
      for (uint32_t ti = 0; ti < ntile; ++ti, ++seq) {
        const uint32_t st = seq % M2_NST;
        const uint32_t tile = wbase_s + st * (uint32_t)M2_STAGE;
        mbar_wait(&full[st], (seq / M2_NST) & 1u);
        const uint32_t rc = ra + ti * M2_ROWS + lane;
        if (rc < rb) {
          double2 v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = m2_pair(tile, lane, k);
          lag_block(std::false_type{}, tile, rc, 2u * blk, v, accb, accf);
        }
        __syncwarp();
        if (lane == 0) issue_next();
      }
      const double totb = warp_reduce_16(accb);
      const double totf = warp_reduce_16(accf);
      if (lane < LAGB) {
        pb[MO_VALS0 + (blk - 1) * LAGB + lane] = totb;
        pf[MO_VALS0 + (blk - 1) * LAGB + lane] = totf;
      }
    }
  }
}

// ---------------------------------------------------------------- later rounds: one kind, lags t0 .. t0+31
// One unit = (listed item, chain, unit of the chain), 256 threads.  Two TMA-staged tiles per sub-tile: first factors y[i] and
// second factors y[i + t0 + ...]; threads 0-127 accumulate lags t0..t0+15, threads 128-255 lags
// t0+16..t0+31 over the same draws, so one pass over a sticky chain advances 32 lags.
struct MoRSmem {
  double2 rawa[MO_RAW0];
  double2 raww[MO_RAWR];
  double ya[MO_PAD0];
  double yw[MO_PADR];
  double red[2][MO_NT / 32][LAGB];
};

// One work unit (one unit of one chain of one listed item) of a later round, two-tile variant.
// `zc` = the chain's first pair, `out` = this unit's LAGR partial sums.  The CTA's mbarrier phase is carried
// across units by the caller (persistent CTAs).
__device__ __forceinline__ void moments_r_unit(MoRSmem& S, const double2* __restrict__ zc, const ChainGeom& g, uint32_t t0,
                                               int kind, uint32_t u, double* __restrict__ out, uint64_t* bar, uint32_t& phase) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int hb = tid >> 7, ct = tid & (MO_NT - 1), wp = ct >> 5;  // hb: which block of 16 lags
  uint32_t a, b;
  unit_range(g, u, a, b);
  const uint32_t lim = (g.N > t0) ? g.N - t0 : 0;  // first factors need i + t0 < N
  b = min(b, lim);
  auto stage = [&](uint32_t sub) {  // sub < lim, so both copies are non-empty
    fence_proxy_async();
    const uint32_t na = min((uint32_t)MO_RAW0, g.N - sub), nw = min((uint32_t)MO_RAWR, g.N - (sub + t0));
    mbar_expect_tx(bar, (na + nw) * 16u);
    bulk_g2s(S.rawa, zc + sub, na * 16u, bar);
    bulk_g2s(S.raww, zc + sub + t0, nw * 16u, bar);
  };
  double acc[LAGB];
#pragma unroll
  for (int l = 0; l < LAGB; ++l) acc[l] = 0.0;
  if (a < b && tid == 0) stage(a);
  for (uint32_t sub = a; sub < b; sub += MO_SUB) {
    const uint32_t len = min((uint32_t)MO_SUB, b - sub);
    const uint32_t na = min((uint32_t)MO_RAW0, g.N - sub), nw = min((uint32_t)MO_RAWR, g.N - (sub + t0));
    mbar_wait(bar, phase);
    phase ^= 1u;
    {
      const double* ra = reinterpret_cast<const double*>(S.rawa) + kind;
      const double* rw = reinterpret_cast<const double*>(S.raww) + kind;
#pragma unroll
      for (int it = 0; it < MO_SUB / MO_CTA; ++it) {
        const uint32_t i = it * MO_CTA + tid;
        S.ya[pad_idx(i)] = (i < na) ? ra[2 * i] : 0.0;
        S.yw[pad_idx(i)] = (i < nw) ? rw[2 * i] : 0.0;
      }
      if (tid < MO_HALOR) {
        const uint32_t i = MO_SUB + tid;
        S.yw[pad_idx(i)] = (i < nw) ? rw[2 * i] : 0.0;
      }
    }
    __syncthreads();
    if (tid == 0 && sub + MO_SUB < b) stage(sub + MO_SUB);
    const uint32_t i0 = ct * CHUNK;
    if (i0 < len) {
      const uint32_t nvalid = min((uint32_t)CHUNK, len - i0);
      double va[CHUNK], w[2 * CHUNK];
      load16(&S.ya[ct * (CHUNK + 2)], va);
      load16(&S.yw[(ct + hb) * (CHUNK + 2)], w);
      load16(&S.yw[(ct + hb + 1) * (CHUNK + 2)], w + CHUNK);
#pragma unroll
      for (int s = 0; s < CHUNK; ++s) {
        const double v = ((uint32_t)s < nvalid) ? va[s] : 0.0;
#pragma unroll
        for (int l = 0; l < LAGB; ++l) acc[l] = fma(v, w[s + l], acc[l]);
      }
    }
    __syncthreads();
  }
  const double tot = warp_reduce_16(*reinterpret_cast<double (*)[16]>(acc));
  if (lane < LAGB) S.red[hb][wp][lane] = tot;
  __syncthreads();
  if (ct < LAGB) {
    double v = S.red[hb][0][ct];
#pragma unroll
    for (int k = 1; k < MO_NT / 32; ++k) v += S.red[hb][k][ct];
    out[hb * LAGB + ct] = v;
  }
  __syncthreads();  // S.red and the staging buffers are free for the CTA's next unit
}

// ---------------------------------------------------------------- later rounds, near lags: ONE staged tile
// While t0 is small against the sub-tile (t0 <= MO_T0MAX) the second factors y[i + t0 + ...] lie in the same
// stretch of the chain as the first ones: one TMA copy of 2048 + t0 + 32 pairs and one de-interleave per
// sub-tile instead of two of each (the two-tile kernel above measured L1-bound, 77 %).  t0 - 1 is a multiple
// of 16 (t0 = 17 + 32 r), so a thread's second-factor window starts one draw after a 16-draw group boundary:
// it loads the two groups from that boundary plus one draw and uses window[s + l + 1].
constexpr int MO_T0MAX = 1 + LAGB + 32 * LAGR;                 // rounds 1 .. 33
constexpr int MO_RAW1 = MO_SUB + MO_T0MAX + LAGR + CHUNK;      // pairs staged per sub-tile (upper bound)
constexpr int MO_PAD1 = (MO_RAW1 / CHUNK + 2) * (CHUNK + 2);
struct MoR1Smem {
  double2 raw[MO_RAW1];
  double y[MO_PAD1];
  double red[2][MO_NT / 32][LAGB];
};

// One work unit of a later round, single-tile variant: sub-tiles [a, b) of the chain whose first pair is `zc`.
struct R1Unit {
  const double2* zc;
  uint32_t a, b;   // a >= b: nothing to do (all first factors of the unit lie past N - t0)
  int kind;
  double* out;     // LAGR partial sums
};

// `staged`: this unit's first copy was already issued (by the previous unit of this CTA); `next`: the CTA's next
// unit (or null) -- its first copy is issued as soon as this unit's last sub-tile has left the staging buffer, so
// that only the first copy of a CTA's round is exposed.
__device__ __forceinline__ void moments_r1_unit(MoR1Smem& S, const R1Unit& un, const ChainGeom& g, uint32_t t0, bool staged,
                                                const R1Unit* next, uint64_t* bar, uint32_t& phase) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int hb = tid >> 7, ct = tid & (MO_NT - 1), wp = ct >> 5;  // hb: which block of 16 lags
  const uint32_t a = un.a, b = un.b;
  const int kind = un.kind;
  const double2* zc = un.zc;
  const uint32_t span = MO_SUB + t0 + LAGR + CHUNK;  // pairs a full sub-tile touches
  double acc[LAGB];
#pragma unroll
  for (int l = 0; l < LAGB; ++l) acc[l] = 0.0;
  if (a < b && !staged && tid == 0) stage_pairs(S.raw, zc + a, min(span, g.N - a), bar);
  const uint32_t goff = (t0 - 1) / CHUNK + (uint32_t)hb;  // group offset of the second-factor window
  for (uint32_t sub = a; sub < b; sub += MO_SUB) {
    const uint32_t len = min((uint32_t)MO_SUB, b - sub);
    const uint32_t nel = min(span, g.N - sub);                       // pairs the copy delivered; the rest reads as 0
    const uint32_t need = min(span, len + t0 + LAGR + CHUNK);        // pairs this sub-tile's threads touch
    mbar_wait(bar, phase);
    phase ^= 1u;
    {
      const double* rk = reinterpret_cast<const double*>(S.raw) + kind;
      // groups of 4 reads ahead of their writes (the compiler keeps a load -> store order on shared memory)
      for (uint32_t i0 = tid; i0 < need; i0 += 4 * MO_CTA) {
        double t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t i = i0 + j * MO_CTA;
          t[j] = (i < nel) ? rk[2 * i] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t i = i0 + j * MO_CTA;
          if (i < need) S.y[pad_idx(i)] = t[j];
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (sub + MO_SUB < b) stage_pairs(S.raw, zc + sub + MO_SUB, min(span, g.N - (sub + MO_SUB)), bar);
      else if (next && next->a < next->b) stage_pairs(S.raw, next->zc + next->a, min(span, g.N - next->a), bar);
    }
    const uint32_t i0 = ct * CHUNK;
    if (i0 < len) {
      const uint32_t nvalid = min((uint32_t)CHUNK, len - i0);
      double va[CHUNK], w[2 * CHUNK + 1];
      load16(&S.y[ct * (CHUNK + 2)], va);
      load16(&S.y[(ct + goff) * (CHUNK + 2)], w);
      load16(&S.y[(ct + goff + 1) * (CHUNK + 2)], w + CHUNK);
      w[2 * CHUNK] = S.y[(ct + goff + 2) * (CHUNK + 2)];
#pragma unroll
      for (int s = 0; s < CHUNK; ++s) {
        const double v = ((uint32_t)s < nvalid) ? va[s] : 0.0;
#pragma unroll
        for (int l = 0; l < LAGB; ++l) acc[l] = fma(v, w[s + l + 1], acc[l]);
      }
    }
    __syncthreads();
  }
  const double tot = warp_reduce_16(*reinterpret_cast<double (*)[16]>(acc));
  if (lane < LAGB) S.red[hb][wp][lane] = tot;
  __syncthreads();
  if (ct < LAGB) {
    double v = S.red[hb][0][ct];
#pragma unroll
    for (int k = 1; k < MO_NT / 32; ++k) v += S.red[hb][k][ct];
    un.out[hb * LAGB + ct] = v;
  }
  __syncthreads();  // S.red is free for the CTA's next unit
}

constexpr int DEC_RED = LAGR + 2;
// block-wide sums of n values (n <= DEC_RED) over NT threads, fixed order; result in out[] (shared)
template <int NT>
__device__ __forceinline__ void dec_reduce(double* vals, int n, double (*red)[DEC_RED], double* out) {
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int k = 0; k < n; ++k) {
    double v = vals[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0) red[wp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < n) {
    double v = 0.0;
    for (int k = 0; k < NT / 32; ++k) v += red[k][threadIdx.x];
    out[threadIdx.x] = v;
  }
  __syncthreads();
}

// rho_t for the nl lags of this round; returns true when the walk is finished.
__device__ __forceinline__ bool rho_walk(const double* lagsum, int nl, uint32_t t0, uint32_t N, uint32_t C, double varhat,
                                         double& rho_sum, long long& lag_out) {
  const double mv = (double)C * varhat;
  for (int l = 0; l < nl; ++l) {
    const uint32_t t = t0 + l;
    if (t >= N) return true;
    const double rho = (lagsum[l] / (double)(N - t)) / mv;
    if (rho < 0.0) { lag_out = (long long)t; return true; }
    rho_sum += rho;
  }
  return t0 + (uint32_t)nl >= N;  // no lag left for another round
}

// ESS = m n / max(-1 + 2 sum of pairs, 1 / log10(m n))
__device__ __forceinline__ double geyer_ess(double mn, double pair_sum) {
  const double tau = -1.0 + 2.0 * pair_sum, floor_tau = 1.0 / log10(mn);
  return mn / ((tau > floor_tau) ? tau : floor_tau);
}

// EXTENSION (not in the reference): Geyer's initial monotone sequence on the same rho_t, rho_0 := 1.
// Pairs P_k = rho_2k + rho_2k+1; stop before the first negative pair, cap every pair by its
// predecessor; only complete pairs count.  `pend` carries the open pair's first term across rounds.
__device__ __forceinline__ bool geyer_walk(const double* lagsum, int nl, uint32_t t0, uint32_t N, uint32_t C, double varhat,
                                           double& pair_sum, double& pend, double& prev, long long& lag_out) {
  const double mv = (double)C * varhat;
  for (int l = 0; l < nl; ++l) {
    const uint32_t t = t0 + l;
    if (t >= N) return true;
    const double rho = (lagsum[l] / (double)(N - t)) / mv;
    if (t & 1u) {  // odd lag closes the pair opened by rho_{t-1}
      double pair = pend + rho;
      if (pair < 0.0) { lag_out = (long long)(t - 1); return true; }
      pair = (pair < prev) ? pair : prev;
      pair_sum += pair;
      prev = pair;
    } else {
      pend = rho;
    }
  }
  return t0 + (uint32_t)nl >= N;  // no lag left for another round (an open pair is dropped)
}

// ---------------------------------------------------------------- round 0 decision: R-hat, var_hat, lags 1..16
__global__ void __launch_bounds__(DEC_NT) k_decide0(const double2* __restrict__ zz, ChainGeom g, int PB,
                                                    ChainState st) {
  extern __shared__ double dsm[];
  double* cm = dsm;          // [C] unsplit means
  double* hm = dsm + g.C;    // [2C] split means
  __shared__ double red[DEC_NT / 32][DEC_RED];
  __shared__ double tot[DEC_RED];
  const int item = blockIdx.x, p = item >> 1, kind = item & 1;
  const uint32_t C = g.C, N = g.N, half = g.half;
  const double dn = (double)N, dh = (double)half;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  const bool ext = st.sticky[p] != 0;  // round 0 also holds lags 17 .. 48 of this parameter
  double vals[2 + LAGB];
  for (int k = 0; k < 2 + LAGB; ++k) vals[k] = 0.0;
  double valx[LAG0X];
  for (int k = 0; k < LAG0X; ++k) valx[k] = 0.0;
  for (uint32_t c = threadIdx.x; c < C; c += DEC_NT) {
    const double* pc = st.part + (((size_t)p * C + c) * g.nunits * 2 + kind) * MO_VALS;
    const double* y = reinterpret_cast<const double*>(zz + ((size_t)p * C + c) * N) + kind;
    double s1 = 0, q1 = 0, s2 = 0, q2 = 0;
    // a half (or the whole chain) is constant iff no unit saw a draw different from its first
    // one and all those first draws agree: its variance is then exactly 0 (diagnostics.py:196-201)
    bool k1 = true, k2 = true, seen1 = false, seen2 = false;
    double r1 = 0.0, r2 = 0.0;
    double R[LAGB], RX[LAG0X];
    for (int l = 0; l < LAGB; ++l) R[l] = 0.0;
    for (int l = 0; l < LAG0X; ++l) RX[l] = 0.0;
    for (uint32_t u = 0; u < g.nunits; ++u) {
      const double* pu = pc + (size_t)u * 2 * MO_VALS;
      for (int l = 0; l < LAGB; ++l) R[l] += pu[l];
      if (ext)
        for (int l = 0; l < LAG0X; ++l) RX[l] += pu[MO_VALS0 + l];
      const double* sv = pu + LAGB;
      if (sv[4] != 0.0) {
        if (!seen1) { r1 = sv[2]; seen1 = true; }
        s1 += sv[0]; q1 += sv[1]; k1 = k1 && (sv[3] == 0.0) && (sv[2] == r1);
      }
      if (sv[9] != 0.0) {
        if (!seen2) { r2 = sv[7]; seen2 = true; }
        s2 += sv[5]; q2 += sv[6]; k2 = k2 && (sv[8] == 0.0) && (sv[7] == r2);
      }
    }
    double sum = s1 + s2, sq = q1 + q2;
    bool kf = k1 && k2 && (r1 == r2);
    if (N & 1u) {  // the draw dropped by the split
      const double yl = y[2 * (size_t)(N - 1)];
      sum += yl; sq = fma(yl, yl, sq);
      kf = kf && (half == 0 || yl == r1);
    }
    const double m1 = s1 / dh, m2 = s2 / dh, mf = sum / dn;
    double v1 = 0.0, v2 = 0.0, vf = 0.0;
    if (half >= 2) {
      if (!k1) v1 = fmax(0.0, (q1 - s1 * m1) / (dh - 1.0));
      if (!k2) v2 = fmax(0.0, (q2 - s2 * m2) / (dh - 1.0));
    }
    if (N >= 2 && !kf) vf = fmax(0.0, (sq - sum * mf) / (dn - 1.0));
    cm[c] = mf; hm[2 * c] = m1; hm[2 * c + 1] = m2;
    vals[0] += v1 + v2;
    vals[1] += vf;
    double hs = 0.0, ts = 0.0;
    for (int l = 0; l < LAGB; ++l) {
      const uint32_t t = 1 + l;
      if (t < N) {
        hs += y[2 * (size_t)(t - 1)];
        ts += y[2 * (size_t)(N - t)];
        vals[2 + l] += R[l] - mf * (2.0 * sum - hs - ts) + (double)(N - t) * mf * mf;
      }
    }
    st.cmean[(size_t)item * C + c] = mf;
    st.csum[(size_t)item * C + c] = sum;
    // the work list a still unfinished item will be on starts at lag 17, or at lag 49 for a sticky parameter: the
    // head / tail sums kept for it are those of ITS first uncovered lag
    double hx = hs, tx = ts;
    if (ext) {
      for (int l = 0; l < LAG0X; ++l) {
        const uint32_t t = 1 + LAGB + l;
        if (t < N) {
          hx += y[2 * (size_t)(t - 1)];
          tx += y[2 * (size_t)(N - t)];
          valx[l] += RX[l] - mf * (2.0 * sum - hx - tx) + (double)(N - t) * mf * mf;
        }
      }
    }
    st.head[(size_t)item * C + c] = ext ? hx : hs;
    st.tail[(size_t)item * C + c] = ext ? tx : ts;
  }
  dec_reduce<DEC_NT>(vals, 2 + LAGB, red, tot);
  __shared__ double totx[DEC_RED];
  if (ext) dec_reduce<DEC_NT>(valx, LAG0X, red, totx);  // (uniform over the CTA)
  if (threadIdx.x == 0) {
    const size_t o = (size_t)kind * PB + p;
    // ---- R-hat on the 2C split chains of length half ----
    double rh;
    const uint32_t m = 2 * C;
    if (half < 2) {
      rh = qnan;
    } else {
      double mt = 0.0;
      for (uint32_t k = 0; k < m; ++k) mt += hm[k];
      mt /= (double)m;
      double bsum = 0.0;
      for (uint32_t k = 0; k < m; ++k) { const double d = hm[k] - mt; bsum += d * d; }
      const double var_between = dh * bsum / (double)(m - 1);
      const double var_within = tot[0] / (double)m;
      const double var_hat = (dh - 1.0) / dh * var_within + var_between / dh;
      if (var_within == 0.0) rh = (var_between == 0.0) ? 1.0 : inf;
      else rh = sqrt(var_hat / var_within);
    }
    st.rhat[o] = rh;
    // ---- ESS on the C unsplit chains ----
    st.lag[o] = (long long)N;
    if (N < 2) {
      st.ess[o] = qnan;
    } else {
      double mt = 0.0;
      for (uint32_t k = 0; k < C; ++k) mt += cm[k];
      mt /= (double)C;
      double bsum = 0.0;
      for (uint32_t k = 0; k < C; ++k) { const double d = cm[k] - mt; bsum += d * d; }
      const double var_between = (C > 1) ? dn * bsum / (double)(C - 1) : 0.0;
      const double var_within = tot[1] / (double)C;
      const double var_hat = (dn - 1.0) / dn * var_within + var_between / dn;
      if (var_hat == 0.0) {
        st.ess[o] = (double)C * dn;
      } else {
        double rho_sum = 0.0;
        long long lag = (long long)N;
        // an unfinished item goes on in the later rounds: from lag 17 (list 0, the first round's) or, with a sticky
        // parameter's 48 lags behind it, from lag 49 (list 1, the second round's)
        const int lst = ext ? 1 : 0;
        if (!st.geyer) {
          bool done = rho_walk(tot + 2, LAGB, 1, N, C, var_hat, rho_sum, lag);
          if (!done && ext) done = rho_walk(totx, LAG0X, 1 + LAGB, N, C, var_hat, rho_sum, lag);
          if (done) {
            st.ess[o] = (double)C * dn / (1.0 + 2.0 * rho_sum);
            st.lag[o] = lag;
          } else {
            st.varhat[item] = var_hat;
            st.rhosum[item] = rho_sum;
            st.wl[lst][atomicAdd(&st.nwl[lst], 1)] = item;
          }
        } else {
          double pend = 1.0, prev = inf;  // rho_0 := 1 opens the first pair
          bool done = geyer_walk(tot + 2, LAGB, 1, N, C, var_hat, rho_sum, pend, prev, lag);
          if (!done && ext) done = geyer_walk(totx, LAG0X, 1 + LAGB, N, C, var_hat, rho_sum, pend, prev, lag);
          if (done) {
            st.ess[o] = geyer_ess((double)C * dn, rho_sum);
            st.lag[o] = lag;
          } else {
            st.varhat[item] = var_hat;
            st.rhosum[item] = rho_sum;
            st.gpend[item] = pend;
            st.gprev[item] = prev;
            st.wl[lst][atomicAdd(&st.nwl[lst], 1)] = item;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- later rounds: lags t0 .. t0+31 of the listed items
// Decision for list entry `li` of the current list (NT threads of the CTA take part): sums the round's
// partial lag products over chains and units, centres them, continues the rho walk.  Partial sums and
// list entries were written by other CTAs of the same launch: read past L1 (ld.global.cg).
template <int NT>
__device__ __forceinline__ void decide_r_item(const double2* __restrict__ zz, const ChainGeom& g, int PB, uint32_t t0, int cur,
                                              int li, const ChainState& st, double (*red)[DEC_RED], double* tot) {
  const int item = __ldcg(&st.wl[cur][li]), p = item >> 1, kind = item & 1;
  const uint32_t C = g.C, N = g.N;
  double vals[LAGR];
  for (int k = 0; k < LAGR; ++k) vals[k] = 0.0;
  for (uint32_t c = threadIdx.x; c < C; c += NT) {
    const double* pc = st.partr + ((size_t)li * C + c) * g.nunits * LAGR;
    const double* y = reinterpret_cast<const double*>(zz + ((size_t)p * C + c) * N) + kind;
    double R[LAGR];
    for (int l = 0; l < LAGR; ++l) R[l] = 0.0;
    for (uint32_t u = 0; u < g.nunits; ++u)
      for (int l = 0; l < LAGR; ++l) R[l] += __ldcg(&pc[(size_t)u * LAGR + l]);
    const double mf = st.cmean[(size_t)item * C + c], sum = st.csum[(size_t)item * C + c];
    double hs = __ldcg(&st.head[(size_t)item * C + c]), ts = __ldcg(&st.tail[(size_t)item * C + c]);
    for (int l = 0; l < LAGR; ++l) {
      const uint32_t t = t0 + l;
      if (t < N) {
        hs += y[2 * (size_t)(t - 1)];
        ts += y[2 * (size_t)(N - t)];
        vals[l] += R[l] - mf * (2.0 * sum - hs - ts) + (double)(N - t) * mf * mf;
      }
    }
    st.head[(size_t)item * C + c] = hs;
    st.tail[(size_t)item * C + c] = ts;
  }
  dec_reduce<NT>(vals, LAGR, red, tot);
  if (threadIdx.x == 0) {
    const size_t o = (size_t)kind * PB + p;
    double rho_sum = __ldcg(&st.rhosum[item]);
    long long lag = (long long)N;
    if (!st.geyer) {
      if (rho_walk(tot, LAGR, t0, N, C, st.varhat[item], rho_sum, lag)) {
        st.ess[o] = (double)C * (double)N / (1.0 + 2.0 * rho_sum);
        st.lag[o] = lag;
      } else {
        st.rhosum[item] = rho_sum;
        st.wl[cur ^ 1][atomicAdd(&st.nwl[cur ^ 1], 1)] = item;
      }
    } else {
      double pend = __ldcg(&st.gpend[item]), prev = __ldcg(&st.gprev[item]);
      if (geyer_walk(tot, LAGR, t0, N, C, st.varhat[item], rho_sum, pend, prev, lag)) {
        st.ess[o] = geyer_ess((double)C * (double)N, rho_sum);
        st.lag[o] = lag;
      } else {
        st.rhosum[item] = rho_sum;
        st.gpend[item] = pend;
        st.gprev[item] = prev;
        st.wl[cur ^ 1][atomicAdd(&st.nwl[cur ^ 1], 1)] = item;
      }
    }
  }
  __syncthreads();  // red / tot are free for the CTA's next item
}

// ---------------------------------------------------------------- all later rounds in ONE launch
// Persistent cooperative grid (every CTA resident): per round of 32 lags, the CTAs stride over the work
// units of the listed (parameter, kind) items, a grid-wide barrier, the CTAs stride over the items'
// decisions (which build the next list), a grid-wide barrier.  The loop ends on the device when the list is
// empty or no lag is left -- the host neither reads the list length nor synchronises between rounds
// (round 1's driver copied `nwork` back and blocked before every round).
struct RoundsSmem {
  union {
    MoR1Smem r1;
    MoRSmem r;
  } m;
  double red[MO_CTA / 32][DEC_RED];
  double tot[DEC_RED];
  uint64_t bar;  // the one mbarrier both tile variants signal (its phase is carried across units and rounds)
};

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& target) {
  // all CTAs are co-resident (cooperative launch).  Monotone counter: barrier k completes at k * gridDim.x.
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    const long long t_start = clock64();
    while (*(volatile unsigned int*)bar < target) {
      if (clock64() - t_start > 20000000000ll) {  // ~10 s: a protocol bug traps instead of hanging the GPU
        printf("b200ref: grid barrier timed out (block %d)\n", blockIdx.x);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(MO_CTA, 2) k_ess_rounds(const double2* __restrict__ zz, ChainGeom g, int PB, ChainState st,
                                                          unsigned int* gbar) {
  extern __shared__ __align__(128) unsigned char mo_smem_raw[];
  RoundsSmem& S = *reinterpret_cast<RoundsSmem*>(mo_smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&S.bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  unsigned int target = 0;
  const uint32_t upi = g.C * g.nunits;  // work units per listed item
  int cur = 0;
  for (uint32_t t0 = 1 + LAGB; t0 < g.N; t0 += LAGR, cur ^= 1) {
    const int nwork = __ldcg(&st.nwl[cur]);  // final since the barrier that ended the previous round
    const bool first = t0 == 1u + LAGB;
    if (nwork == 0) {                        // uniform over the grid
      // the second round's list may already hold the sticky parameters' items (they start at lag 49, k_decide0)
      if (first && __ldcg(&st.nwl[1]) > 0) continue;
      break;
    }
    const uint64_t total = (uint64_t)nwork * upi;
    const uint32_t lim = (g.N > t0) ? g.N - t0 : 0;  // first factors need i + t0 < N
    auto describe = [&](uint64_t w, R1Unit& un, uint32_t& u_out) {
      const int li = (int)(w / upi);
      const uint32_t cu = (uint32_t)(w - (uint64_t)li * upi);
      const uint32_t c = cu / g.nunits, u = cu - c * g.nunits;
      const int item = __ldcg(&st.wl[cur][li]);
      un.zc = zz + ((size_t)(item >> 1) * g.C + c) * g.N;
      un.kind = item & 1;
      un.out = st.partr + (((size_t)li * g.C + c) * g.nunits + u) * LAGR;
      unit_range(g, u, un.a, un.b);
      un.b = min(un.b, lim);
      u_out = u;
    };
    if (t0 <= (uint32_t)MO_T0MAX) {
      // single-tile variant: the next unit's first copy is issued under the current unit's last sub-tile
      R1Unit cu_, nx_;
      uint32_t u_;
      bool staged = false;
      uint64_t w = blockIdx.x;
      if (w < total) describe(w, cu_, u_);
      for (; w < total; w += gridDim.x) {
        const bool has_next = w + gridDim.x < total;
        if (has_next) describe(w + gridDim.x, nx_, u_);
        moments_r1_unit(S.m.r1, cu_, g, t0, staged, has_next ? &nx_ : nullptr, &S.bar, phase);
        // the prefetch is issued from inside the sub-tile loop: only a unit that has sub-tiles can have issued it
        staged = has_next && (cu_.a < cu_.b) && (nx_.a < nx_.b);
        cu_ = nx_;
      }
    } else {
      for (uint64_t w = blockIdx.x; w < total; w += gridDim.x) {
        R1Unit un;
        uint32_t u;
        describe(w, un, u);
        moments_r_unit(S.m.r, un.zc, g, t0, un.kind, u, un.out, &S.bar, phase);
      }
    }
    // the list the decisions are about to build (the second round's was started by k_decide0: it is kept)
    if (blockIdx.x == 0 && tid == 0 && !first) st.nwl[cur ^ 1] = 0;
    grid_barrier(gbar, target);
    for (int li = blockIdx.x; li < nwork; li += gridDim.x)
      decide_r_item<MO_CTA>(zz, g, PB, t0, cur, li, st, S.red, S.tot);
    grid_barrier(gbar, target);
  }
}

}  // namespace b200ref
