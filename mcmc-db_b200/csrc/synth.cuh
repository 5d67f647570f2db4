// Synthetic workload of SURVEY.md 8(d) -- bench / test utility, not on the hot path.
// Counter-based: every value is a pure function of (seed, global parameter index,
// chain, draw), so shards generated on different GPUs or in different tiles agree.
#pragma once
#include "common.cuh"

namespace b200ref {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ double std_normal(uint64_t key, uint64_t i) {
  const uint64_t h1 = mix64(key + i * 0xd1342543de82ef95ull);
  const uint64_t h2 = mix64(h1 ^ 0x2545f4914f6cdd1dull);
  const double u1 = ((double)(h1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)(h2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// One warp per (parameter, chain).  kind 0: candidate draws, AR(1) with
// phi_p in {0, 0.5, 0.9} (p mod 3), mu_p = (p mod 17) - 8, sigma_p = 2^((p mod 5) - 2);
// every 97th parameter gets a chain offset of 0.5 sigma (c mod 2) (R-hat failures);
// every 101st is rounded to 2 decimals (ties).  kind 1: reference posterior, iid
// N(mu_p, sigma_p^2), every 89th parameter's mean shifted by 0.3 sigma (compare failures).
__global__ void k_synth(double* __restrict__ x, int64_t p0, int64_t P, int64_t C, int64_t N, uint64_t seed,
                        int kind) {
  const int lane = threadIdx.x & 31;
  const int64_t chain_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chain_id >= P * C) return;
  const int64_t pl = chain_id / C, c = chain_id % C;
  const int64_t gp = p0 + pl;
  const double mu = (double)(gp % 17) - 8.0;
  const double sigma = exp2((double)(gp % 5) - 2.0);
  const uint64_t key = mix64(seed ^ mix64((uint64_t)gp * 0x100000001b3ull + (uint64_t)c + ((uint64_t)kind << 62)));
  double* xc = x + ((size_t)pl * C + c) * N;
  if (kind == 1) {
    const double shift = (gp % 89 == 88) ? 0.3 * sigma : 0.0;
    for (int64_t i = lane; i < N; i += 32) xc[i] = mu + shift + sigma * std_normal(key, (uint64_t)i);
    return;
  }
  const int r3 = (int)(gp % 3);
  const double phi = (r3 == 0) ? 0.0 : (r3 == 1 ? 0.5 : 0.9);
  const double s = sqrt(1.0 - phi * phi);
  const double off = (gp % 97 == 96) ? 0.5 * sigma * (double)(c % 2) : 0.0;
  const bool round2 = (gp % 101 == 100);
  const double plane = pow(phi, (double)(lane + 1));
  double pw[5];
  pw[0] = phi;
#pragma unroll
  for (int k = 1; k < 5; ++k) pw[k] = pw[k - 1] * pw[k - 1];
  double carry = 0.0;
  for (int64_t i0 = 0; i0 < N; i0 += 32) {
    const int64_t i = i0 + lane;
    double v = std_normal(key, (uint64_t)i);
    if (i > 0) v *= s;  // a_0 = e_0 (stationary start)
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double t = __shfl_up_sync(FULL, v, 1 << k);
      if (lane >= (1 << k)) v += pw[k] * t;
    }
    v += plane * carry;
    carry = __shfl_sync(FULL, v, 31);
    double val = mu + off + sigma * v;
    if (round2) val = rint(val * 100.0) / 100.0;
    if (i < N) xc[i] = val;
  }
}

}  // namespace b200ref
