// Per-parameter chain statistics on the normal scores (original positions):
//   split-chain moments -> R-hat            (_split_chains :76-85, _rhat :136-151,
//                                             _variance :196-201 of diagnostics.py)
//   unsplit-chain moments + autocovariance
//   with first-negative truncation -> ESS   (_ess :154-177, _autocorr :180-193)
//
// Streaming design.  zz[p][c][i] = {z_bulk, z_folded}.  Every chain is cut into units
// (pieces of one half-chain, <= 8192 draws); a 128-thread CTA streams its unit through
// shared memory in 2048-draw sub-tiles and accumulates, for both kinds at once,
//   * sum, sum of squares, min, max of the unit (raw moments: z-scores have mean ~ 0 and
//     variance ~ 1, so sum(y^2) - n*mean^2 loses no accuracy; min == max -> variance 0 exactly)
//   * 16 raw lag products R_t = sum_i y_i * y_{i+t}, t = t0 .. t0+15, by a register
//     sliding window (16 FMAs per 2 shared loads).
// A small per-(parameter, kind) kernel turns the partial sums into chain means / variances,
// centres the lag products algebraically
//     S_t = R_t - mean * (2*total - head_t - tail_t) + (n - t) * mean^2 ,
// and walks rho_1, rho_2, ... until the first negative one (diagnostics.py:172-176).
// Pairs whose rho is still positive after 16 lags go onto a work list; further rounds
// stream only those chains with t0 advanced by 16.  Well-mixed chains finish in round 0,
// and no chain length limit remains (nothing has to fit on-chip).
// Because all chains of a parameter have the same length n, sum_c cov_c / (n - t) is
// accumulated as (sum_c S_c) / (n - t).
#pragma once
#include "common.cuh"

namespace b200ref {

constexpr int MO_NT = 128;
constexpr int LAGB = 16;                 // lags per round
constexpr int CHUNK = 16;                // consecutive draws per thread step
constexpr int MO_SUB = MO_NT * CHUNK;    // 2048 draws per sub-tile
constexpr int MO_SEGMAX = 4 * MO_SUB;    // a unit covers at most 8192 draws
constexpr int MO_VALS = LAGB + 4;        // per (unit, kind) in round 0: 16 lag sums + {sum, sq, min, max}
constexpr int DEC_NT = 64;

struct ChainGeom {
  uint32_t C, N, half, nseg, seglen, nunits;  // nunits = 2 * nseg units per chain
};

__host__ inline ChainGeom chain_geom(int64_t C, int64_t N) {
  ChainGeom g;
  g.C = (uint32_t)C;
  g.N = (uint32_t)N;
  g.half = g.N / 2;
  g.nseg = g.half ? (g.half + MO_SEGMAX - 1) / MO_SEGMAX : 0;
  g.seglen = g.nseg ? (((g.half + g.nseg - 1) / g.nseg + CHUNK - 1) / CHUNK) * CHUNK : 0;
  g.nunits = 2 * g.nseg;
  return g;
}

__device__ __forceinline__ void unit_range(const ChainGeom& g, uint32_t u, uint32_t& a, uint32_t& b) {
  const uint32_t h = u / g.nseg, s = u % g.nseg;
  a = h * g.half + s * g.seglen;
  b = min(h * g.half + (s + 1) * g.seglen, (h + 1) * g.half);
  if (a > b) a = b;
}

__host__ __device__ __forceinline__ uint32_t pad_idx(uint32_t e) { return e + (e >> 4); }
constexpr int MO_PAD = MO_SUB + 2 * LAGB + (MO_SUB + 2 * LAGB) / 16 + 2;

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
  double* partr;     // [nwork][C*nunits][LAGB]   later rounds
  double* cmean;     // [2*PB][C]  unsplit chain means
  double* csum;      // [2*PB][C]  chain totals
  double* head;      // [2*PB][C]  sum of the first (t0 - 1) draws of the chain
  double* tail;      // [2*PB][C]  sum of the last  (t0 - 1) draws
  double* varhat;    // [2*PB]
  double* rhosum;    // [2*PB]
  int* wl[2];        // work lists of unfinished items (2*p + kind)
  int* nwl;          // [2] their lengths
  double* rhat;      // [2][PB] outputs, kind-major
  double* ess;       // [2][PB]
  int64_t* lag;      // [2][PB] first lag with rho < 0 (N if none)
};

// acc[l] += sum_s ya[i0 + s] * yw[i0 + s + l + woff]   (s < nvalid; yw zero beyond the chain)
__device__ __forceinline__ void lag_chunk(const double* __restrict__ ya, const double* __restrict__ yw, uint32_t woff,
                                          uint32_t i0, uint32_t nvalid, double* __restrict__ acc) {
  double w[LAGB + CHUNK - 1];
#pragma unroll
  for (int l = 0; l < LAGB - 1; ++l) w[l] = yw[pad_idx(i0 + woff + l)];
#pragma unroll
  for (int s = 0; s < CHUNK; ++s) {
    w[s + LAGB - 1] = yw[pad_idx(i0 + woff + s + LAGB - 1)];
    const double a = ((uint32_t)s < nvalid) ? ya[pad_idx(i0 + s)] : 0.0;
#pragma unroll
    for (int l = 0; l < LAGB; ++l) acc[l] = fma(a, w[s + l], acc[l]);
  }
}

// ---------------------------------------------------------------- round 0: lags 1..16 + moments
// grid (2 * C * nunits, PB): kind = blockIdx.x & 1.  The two sibling CTAs (adjacent block ids,
// co-scheduled) read the two halves of the same 16-byte pairs, so DRAM sees each sector once.
__global__ void __launch_bounds__(MO_NT, 4) k_moments0(const double2* __restrict__ zz, ChainGeom g, double* part) {
  __shared__ double y[MO_PAD];
  __shared__ double red[MO_NT / 32][MO_VALS];
  const int p = blockIdx.y, kind = blockIdx.x & 1;
  const uint32_t cu = blockIdx.x >> 1;
  const uint32_t c = cu / g.nunits, u = cu % g.nunits;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  uint32_t a, b;
  unit_range(g, u, a, b);
  const double* zc = reinterpret_cast<const double*>(zz + ((size_t)p * g.C + c) * g.N) + kind;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double acc[LAGB];
#pragma unroll
  for (int l = 0; l < LAGB; ++l) acc[l] = 0.0;
  double sm = 0.0, sq = 0.0, mn = inf, mx = -inf;
  for (uint32_t sub = a; sub < b; sub += MO_SUB) {
    const uint32_t len = min((uint32_t)MO_SUB, b - sub);
    __syncthreads();
    for (uint32_t i = tid; i < MO_SUB + 2 * LAGB; i += MO_NT) {
      const uint32_t gi = sub + i;
      y[pad_idx(i)] = (gi < g.N) ? zc[2 * (size_t)gi] : 0.0;
    }
    __syncthreads();
    const uint32_t i0 = tid * CHUNK;
    if (i0 < len) {
      const uint32_t nvalid = min((uint32_t)CHUNK, len - i0);
      lag_chunk(y, y, 1, i0, nvalid, acc);
#pragma unroll
      for (int s = 0; s < CHUNK; ++s) {
        if ((uint32_t)s < nvalid) {
          const double v = y[pad_idx(i0 + s)];
          sm += v; sq = fma(v, v, sq); mn = fmin(mn, v); mx = fmax(mx, v);
        }
      }
    }
  }
  // unit totals: 16 lag sums by the halving reduce, 4 statistics by butterflies
  const double tot = warp_reduce_16(*reinterpret_cast<double (*)[16]>(acc));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sm += __shfl_xor_sync(FULL, sm, o); sq += __shfl_xor_sync(FULL, sq, o);
    mn = fmin(mn, __shfl_xor_sync(FULL, mn, o)); mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
  }
  __syncthreads();
  if (lane < LAGB) red[wp][lane] = tot;
  if (lane == 0) { red[wp][16] = sm; red[wp][17] = sq; red[wp][18] = mn; red[wp][19] = mx; }
  __syncthreads();
  if (tid < MO_VALS) {
    double v = red[0][tid];
#pragma unroll
    for (int k = 1; k < MO_NT / 32; ++k) {
      const double o = red[k][tid];
      v = (tid < 18) ? v + o : (tid == 18 ? fmin(v, o) : fmax(v, o));
    }
    part[((((size_t)p * g.C + c) * g.nunits + u) * 2 + kind) * MO_VALS + tid] = v;
  }
}

// ---------------------------------------------------------------- later rounds: one kind, lags t0 .. t0+15
__global__ void __launch_bounds__(MO_NT, 4) k_moments_r(const double2* __restrict__ zz, ChainGeom g, uint32_t t0,
                                                     const int* __restrict__ wl, double* partr) {
  __shared__ double ya[MO_PAD], yw[MO_PAD];
  __shared__ double red[MO_NT / 32][LAGB];
  const int item = wl[blockIdx.y];
  const int p = item >> 1, kind = item & 1;
  const uint32_t c = blockIdx.x / g.nunits, u = blockIdx.x % g.nunits;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  uint32_t a, b;
  unit_range(g, u, a, b);
  const uint32_t lim = (g.N > t0) ? g.N - t0 : 0;  // first factors need i + t0 < N
  b = min(b, lim);
  const double* zc = reinterpret_cast<const double*>(zz + ((size_t)p * g.C + c) * g.N) + kind;
  double acc[LAGB];
#pragma unroll
  for (int l = 0; l < LAGB; ++l) acc[l] = 0.0;
  for (uint32_t sub = a; sub < b; sub += MO_SUB) {
    const uint32_t len = min((uint32_t)MO_SUB, b - sub);
    __syncthreads();
    for (uint32_t i = tid; i < MO_SUB + 2 * LAGB; i += MO_NT) {
      const uint32_t ga = sub + i, gw = sub + t0 + i;
      ya[pad_idx(i)] = (ga < g.N) ? zc[2 * (size_t)ga] : 0.0;
      yw[pad_idx(i)] = (gw < g.N) ? zc[2 * (size_t)gw] : 0.0;
    }
    __syncthreads();
    const uint32_t i0 = tid * CHUNK;
    if (i0 < len) lag_chunk(ya, yw, 0, i0, min((uint32_t)CHUNK, len - i0), acc);
  }
  const double tot = warp_reduce_16(*reinterpret_cast<double (*)[16]>(acc));
  __syncthreads();
  if (lane < LAGB) red[wp][lane] = tot;
  __syncthreads();
  if (tid < LAGB) {
    double v = red[0][tid];
#pragma unroll
    for (int k = 1; k < MO_NT / 32; ++k) v += red[k][tid];
    partr[(((size_t)blockIdx.y * g.C + c) * g.nunits + u) * LAGB + tid] = v;
  }
}

// block-wide sums of n values (n <= 18) over DEC_NT = 64 threads, fixed order; result in out[] (shared)
__device__ __forceinline__ void dec_reduce(double* vals, int n, double (*red)[20], double* out) {
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
    for (int k = 0; k < DEC_NT / 32; ++k) v += red[k][threadIdx.x];
    out[threadIdx.x] = v;
  }
  __syncthreads();
}

// rho_t for the 16 lags of this round; returns true when the walk is finished.
__device__ __forceinline__ bool rho_walk(const double* lagsum, uint32_t t0, uint32_t N, uint32_t C, double varhat,
                                         double& rho_sum, long long& lag_out) {
  const double mv = (double)C * varhat;
  for (int l = 0; l < LAGB; ++l) {
    const uint32_t t = t0 + l;
    if (t >= N) return true;
    const double rho = (lagsum[l] / (double)(N - t)) / mv;
    if (rho < 0.0) { lag_out = (long long)t; return true; }
    rho_sum += rho;
  }
  return false;
}

// ---------------------------------------------------------------- round 0 decision: R-hat, var_hat, lags 1..16
__global__ void __launch_bounds__(DEC_NT) k_decide0(const double2* __restrict__ zz, ChainGeom g, int PB,
                                                    ChainState st) {
  extern __shared__ double dsm[];
  double* cm = dsm;          // [C] unsplit means
  double* hm = dsm + g.C;    // [2C] split means
  __shared__ double red[DEC_NT / 32][20];
  __shared__ double tot[20];
  const int item = blockIdx.x, p = item >> 1, kind = item & 1;
  const uint32_t C = g.C, N = g.N, half = g.half;
  const double dn = (double)N, dh = (double)half;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  double vals[2 + LAGB];
  for (int k = 0; k < 2 + LAGB; ++k) vals[k] = 0.0;
  for (uint32_t c = threadIdx.x; c < C; c += DEC_NT) {
    const double* pc = st.part + (((size_t)p * C + c) * g.nunits * 2 + kind) * MO_VALS;
    const double* y = reinterpret_cast<const double*>(zz + ((size_t)p * C + c) * N) + kind;
    double s1 = 0, q1 = 0, mn1 = inf, mx1 = -inf, s2 = 0, q2 = 0, mn2 = inf, mx2 = -inf;
    double R[LAGB];
    for (int l = 0; l < LAGB; ++l) R[l] = 0.0;
    for (uint32_t u = 0; u < g.nunits; ++u) {
      const double* pu = pc + (size_t)u * 2 * MO_VALS;
      for (int l = 0; l < LAGB; ++l) R[l] += pu[l];
      const double* sv = pu + LAGB;
      if (u < g.nseg) { s1 += sv[0]; q1 += sv[1]; mn1 = fmin(mn1, sv[2]); mx1 = fmax(mx1, sv[3]); }
      else            { s2 += sv[0]; q2 += sv[1]; mn2 = fmin(mn2, sv[2]); mx2 = fmax(mx2, sv[3]); }
    }
    double sum = s1 + s2, sq = q1 + q2, mn = fmin(mn1, mn2), mx = fmax(mx1, mx2);
    if (N & 1u) {  // the draw dropped by the split
      const double yl = y[2 * (size_t)(N - 1)];
      sum += yl; sq = fma(yl, yl, sq); mn = fmin(mn, yl); mx = fmax(mx, yl);
    }
    const double m1 = s1 / dh, m2 = s2 / dh, mf = sum / dn;
    double v1 = 0.0, v2 = 0.0, vf = 0.0;
    if (half >= 2) {
      if (mn1 != mx1) v1 = fmax(0.0, (q1 - s1 * m1) / (dh - 1.0));
      if (mn2 != mx2) v2 = fmax(0.0, (q2 - s2 * m2) / (dh - 1.0));
    }
    if (N >= 2 && mn != mx) vf = fmax(0.0, (sq - sum * mf) / (dn - 1.0));
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
    st.head[(size_t)item * C + c] = hs;
    st.tail[(size_t)item * C + c] = ts;
  }
  dec_reduce(vals, 2 + LAGB, red, tot);
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
        if (rho_walk(tot + 2, 1, N, C, var_hat, rho_sum, lag)) {
          st.ess[o] = (double)C * dn / (1.0 + 2.0 * rho_sum);
          st.lag[o] = lag;
        } else {
          st.varhat[item] = var_hat;
          st.rhosum[item] = rho_sum;
          st.wl[0][atomicAdd(&st.nwl[0], 1)] = item;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- later rounds: lags t0 .. t0+15 of the listed items
__global__ void __launch_bounds__(DEC_NT) k_decide_r(const double2* __restrict__ zz, ChainGeom g, int PB, uint32_t t0,
                                                     int cur, ChainState st) {
  __shared__ double red[DEC_NT / 32][20];
  __shared__ double tot[20];
  const int item = st.wl[cur][blockIdx.x], p = item >> 1, kind = item & 1;
  const uint32_t C = g.C, N = g.N;
  double vals[LAGB];
  for (int k = 0; k < LAGB; ++k) vals[k] = 0.0;
  for (uint32_t c = threadIdx.x; c < C; c += DEC_NT) {
    const double* pc = st.partr + ((size_t)blockIdx.x * C + c) * g.nunits * LAGB;
    const double* y = reinterpret_cast<const double*>(zz + ((size_t)p * C + c) * N) + kind;
    double R[LAGB];
    for (int l = 0; l < LAGB; ++l) R[l] = 0.0;
    for (uint32_t u = 0; u < g.nunits; ++u)
      for (int l = 0; l < LAGB; ++l) R[l] += pc[(size_t)u * LAGB + l];
    const double mf = st.cmean[(size_t)item * C + c], sum = st.csum[(size_t)item * C + c];
    double hs = st.head[(size_t)item * C + c], ts = st.tail[(size_t)item * C + c];
    for (int l = 0; l < LAGB; ++l) {
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
  dec_reduce(vals, LAGB, red, tot);
  if (threadIdx.x == 0) {
    const size_t o = (size_t)kind * PB + p;
    double rho_sum = st.rhosum[item];
    long long lag = (long long)N;
    if (rho_walk(tot, t0, N, C, st.varhat[item], rho_sum, lag)) {
      st.ess[o] = (double)C * (double)N / (1.0 + 2.0 * rho_sum);
      st.lag[o] = lag;
    } else {
      st.rhosum[item] = rho_sum;
      st.wl[cur ^ 1][atomicAdd(&st.nwl[cur ^ 1], 1)] = item;
    }
  }
}

}  // namespace b200ref
