// Per-parameter chain statistics on the normal scores z[C][N] (original positions):
//   split-chain moments -> R-hat            (_split_chains :76-85, _rhat :136-151,
//                                             _variance :196-201 of diagnostics.py)
//   unsplit-chain moments + autocovariance
//   with first-negative truncation -> ESS   (_ess :154-177, _autocorr :180-193)
//
// One CTA per (parameter, kind) with kind = bulk | folded.  Each chain is staged in
// shared memory once per lag block; lags are evaluated 16 at a time by direct fp64
// products (a register sliding window gives 16 FMAs per 2 shared loads).  The
// reference stops at the first lag with rho < 0, so well-mixed chains need exactly
// one lag block; sticky chains loop over further blocks until the sign flips.
// Because all chains of a parameter have the same length n, sum_c cov_c / (n - t)
// is accumulated as (sum_c S_c) / (n - t).
#pragma once
#include "common.cuh"

namespace b200ref {

constexpr int CH_NT = 256;
constexpr int LAGB = 16;   // lags per block
constexpr int CHUNK = 16;  // consecutive draws per thread step

struct ChainOut {
  double* rhat;   // [2][PB]  kind-major
  double* ess;    // [2][PB]
  int64_t* lag;   // [2][PB]  first lag with rho < 0 (N if none / not evaluated)
};

__host__ __device__ __forceinline__ uint32_t pad_idx(uint32_t e) { return e + (e >> 4); }

__host__ inline size_t chain_smem_bytes(int64_t C, int64_t N) {
  const size_t ny = (size_t)pad_idx((uint32_t)N + 2 * LAGB + CHUNK) + 2;
  return (ny + 3 * (size_t)C + 160) * sizeof(double);
}

template <int NT>
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  __syncthreads();
  if (lane == 0) { red[w * 3 + 0] = a; red[w * 3 + 1] = b; red[w * 3 + 2] = c; }
  __syncthreads();
  a = 0; b = 0; c = 0;
#pragma unroll
  for (int k = 0; k < NT / 32; ++k) { a += red[k * 3 + 0]; b += red[k * 3 + 1]; c += red[k * 3 + 2]; }
}

// Accumulate S_t += sum_i y[i] * y[i + t] for t = t0 .. t0 + 15 over this thread's chunks.
// y is zero beyond N, so no bounds test is needed on the second factor.
template <int NT>
__device__ __forceinline__ void lag_block_accumulate(const double* __restrict__ y, uint32_t N, uint32_t t0,
                                                     double (&acc)[LAGB]) {
  if (t0 >= N) return;
  const uint32_t nchunks = (N - t0 + CHUNK - 1) / CHUNK;  // chunks holding some i < N - t0
  for (uint32_t q = threadIdx.x; q < nchunks; q += NT) {
    const uint32_t i0 = q * CHUNK;
    double w[LAGB + CHUNK - 1];
#pragma unroll
    for (int l = 0; l < LAGB - 1; ++l) w[l] = y[pad_idx(i0 + t0 + l)];
#pragma unroll
    for (int s = 0; s < CHUNK; ++s) {
      w[s + LAGB - 1] = y[pad_idx(i0 + t0 + s + LAGB - 1)];
      const double a = y[pad_idx(i0 + s)];
#pragma unroll
      for (int l = 0; l < LAGB; ++l) acc[l] = fma(a, w[s + l], acc[l]);
    }
  }
}

// zz: [PB][C][N] pairs {z_bulk, z_folded}; CTA (2p + kind) reads component `kind`
// (the sibling CTA reads the other half of the same sectors -> L2 hit).
__global__ void __launch_bounds__(CH_NT) k_chain(const double* __restrict__ zz, uint32_t C, uint32_t N, int PB,
                                                 ChainOut out) {
  extern __shared__ double smem[];
  const int p = blockIdx.x >> 1, kind = blockIdx.x & 1;
  const int tid = threadIdx.x;
  const uint32_t ny = pad_idx(N + 2 * LAGB + CHUNK) + 2;
  double* y = smem;             // padded chain
  double* cmean = y + ny;       // [C] unsplit chain means
  double* hmean = cmean + C;    // [2C] split-chain means
  double* red = hmean + 2 * C;  // 160 doubles scratch (8 warps x 16 lags)
  __shared__ double s_w_split, s_w_full, s_varhat, s_rhosum;
  __shared__ int s_done;
  __shared__ long long s_lag;

  const double* z = zz + 2 * (size_t)p * C * N + kind;
  const uint32_t half = N / 2;
  const double dn = (double)N, dh = (double)half;

  if (tid == 0) { s_w_split = 0.0; s_w_full = 0.0; s_rhosum = 0.0; s_done = 0; s_lag = (long long)N; }
  double acc[LAGB];
#pragma unroll
  for (int l = 0; l < LAGB; ++l) acc[l] = 0.0;

  // ---- pass 0 over chains: moments + lag block [1, 16] ----
  for (uint32_t c = 0; c < C; ++c) {
    __syncthreads();
    const double* zc = z + 2 * (size_t)c * N;
    for (uint32_t i = tid; i < N + 2 * LAGB + CHUNK; i += CH_NT) y[pad_idx(i)] = (i < N) ? zc[2 * i] : 0.0;
    __syncthreads();
    double s1 = 0, s2 = 0, s3 = 0;
    for (uint32_t i = tid; i < N; i += CH_NT) {
      const double v = y[pad_idx(i)];
      if (i < half) s1 += v; else if (i < 2 * half) s2 += v; else s3 += v;
    }
    block_sum3<CH_NT>(s1, s2, s3, red);
    const double m1 = s1 / dh, m2 = s2 / dh, mf = (s1 + s2 + s3) / dn;
    double q1 = 0, q2 = 0, qf = 0;
    for (uint32_t i = tid; i < N; i += CH_NT) {
      const uint32_t pi = pad_idx(i);
      const double v = y[pi];
      const double df = v - mf;
      qf += df * df;
      if (i < half) { const double d = v - m1; q1 += d * d; }
      else if (i < 2 * half) { const double d = v - m2; q2 += d * d; }
      y[pi] = df;  // centre for the autocovariance
    }
    block_sum3<CH_NT>(q1, q2, qf, red);
    if (tid == 0) {
      cmean[c] = mf;
      hmean[2 * c] = m1;
      hmean[2 * c + 1] = m2;
      if (half >= 2) s_w_split += q1 / (dh - 1.0) + q2 / (dh - 1.0);
      if (N >= 2) s_w_full += qf / (dn - 1.0);
    }
    __syncthreads();
    lag_block_accumulate<CH_NT>(y, N, 1, acc);
  }

  // ---- R-hat from the split-chain moments; var_hat for the ESS ----
  if (tid == 0) {
    double rh;
    const uint32_t m = 2 * C;
    if (half == 0 || half < 2) {
      rh = __longlong_as_double(0x7ff8000000000000ll);
    } else {
      double mt = 0.0;
      for (uint32_t k = 0; k < m; ++k) mt += hmean[k];
      mt /= (double)m;
      double b = 0.0;
      for (uint32_t k = 0; k < m; ++k) { const double d = hmean[k] - mt; b += d * d; }
      const double var_between = dh * b / (double)(m - 1);
      const double var_within = s_w_split / (double)m;
      const double var_hat = (dh - 1.0) / dh * var_within + var_between / dh;
      if (var_within == 0.0) rh = (var_between == 0.0) ? 1.0 : __longlong_as_double(0x7ff0000000000000ll);
      else rh = sqrt(var_hat / var_within);
    }
    out.rhat[(size_t)kind * PB + p] = rh;

    if (N < 2) {
      out.ess[(size_t)kind * PB + p] = __longlong_as_double(0x7ff8000000000000ll);
      s_done = 1;
    } else {
      double mt = 0.0;
      for (uint32_t k = 0; k < C; ++k) mt += cmean[k];
      mt /= (double)C;
      double b = 0.0;
      for (uint32_t k = 0; k < C; ++k) { const double d = cmean[k] - mt; b += d * d; }
      const double var_between = (C > 1) ? dn * b / (double)(C - 1) : 0.0;
      const double var_within = s_w_full / (double)C;
      const double var_hat = (dn - 1.0) / dn * var_within + var_between / dn;
      s_varhat = var_hat;
      if (var_hat == 0.0) {
        out.ess[(size_t)kind * PB + p] = (double)C * dn;
        s_done = 1;
      }
    }
  }
  __syncthreads();

  // ---- lag blocks until the first negative rho ----
  for (uint32_t t0 = 1; !s_done; t0 += LAGB) {
    if (t0 > 1) {
#pragma unroll
      for (int l = 0; l < LAGB; ++l) acc[l] = 0.0;
      for (uint32_t c = 0; c < C; ++c) {
        __syncthreads();
        const double* zc = z + 2 * (size_t)c * N;
        const double mf = cmean[c];
        for (uint32_t i = tid; i < N + 2 * LAGB + CHUNK; i += CH_NT) y[pad_idx(i)] = (i < N) ? zc[2 * i] - mf : 0.0;
        __syncthreads();
        lag_block_accumulate<CH_NT>(y, N, t0, acc);
      }
    }
    // block-reduce the 16 lag sums
    __syncthreads();
#pragma unroll
    for (int l = 0; l < LAGB; ++l) {
      const double v = warp_sum(acc[l]);
      if ((tid & 31) == 0) red[(tid >> 5) * LAGB + l] = v;
    }
    __syncthreads();
    if (tid == 0) {
      const double mv = (double)C * s_varhat;
      double rho_sum = s_rhosum;
      int done = 0;
      for (int l = 0; l < LAGB; ++l) {
        const uint32_t t = t0 + l;
        if (t >= N) { done = 1; break; }
        double sacc = 0.0;
        for (int wv = 0; wv < CH_NT / 32; ++wv) sacc += red[wv * LAGB + l];
        const double rho = (sacc / (double)(N - t)) / mv;
        if (rho < 0.0) { done = 1; s_lag = (long long)t; break; }
        rho_sum += rho;
      }
      s_rhosum = rho_sum;
      if (done) {
        out.ess[(size_t)kind * PB + p] = (double)C * dn / (1.0 + 2.0 * rho_sum);
        s_done = 1;
      }
    }
    __syncthreads();
  }
  if (tid == 0) out.lag[(size_t)kind * PB + p] = s_lag;
}

}  // namespace b200ref
