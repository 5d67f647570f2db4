// Segmented (one segment = one parameter's pooled M = C*N draws) LSD radix sort of
// (key64, position32) pairs.  One CTA owns one parameter for the whole sort, so
// digit offsets are carried in shared memory from tile to tile and no cross-CTA
// look-back is needed.  Since the distribution sort (bucket.cuh) became the main path this
// is the FALLBACK: a fixed small grid strides over the compact list of parameters whose
// data defeated the bucket function (empty list -> the launches return at once).
//
// This is the ordering step of _rank_normalize (diagnostics.py:110) and of
// np.quantile / statistics.median (backends_numpy.py:44, diagnostics.py:97).
//
// Passes: 8 digits of 8 bits.  k_hist0 reads the draws once (HBM) and builds the
// histograms of digit 0 and of the three top digits (sign/exponent bytes, which
// are often constant within a parameter -> that pass is skipped).  Scatter pass d
// builds the histogram of digit d+1 on the fly for d = 0..3.
#pragma once
#include "common.cuh"

namespace b200ref {

constexpr int SORT_NT = 512;             // threads per CTA
constexpr int SORT_IPT = 8;              // keys per thread per tile
constexpr int SORT_NW = SORT_NT / 32;    // warps
constexpr int SORT_TILE = SORT_NT * SORT_IPT;

struct SortBufs {
  uint64_t* key[2];   // [PB][M] ping-pong
  uint32_t* idx[2];   // [PB][M]
  uint32_t* hist;     // [PB][8][256]
  int* sel;           // [PB] which of key[]/idx[] holds the current sequence
  double* sum;        // [PB] sum of the draws (for the mean)
  int* flag;          // bit0: non-finite input seen
  const int* nlist;   // number of parameters to sort (device counter)
  const int* plist;   // [*nlist] their indices; CTAs stride over this list
};

__device__ __forceinline__ void hist_add(uint32_t* h, uint32_t d, bool valid) {
  // warp-uniform digits (typical for the exponent bytes) -> one atomic per warp
  int all_same;
  __match_all_sync(FULL, valid ? d : 0xffffffffu, &all_same);
  if (all_same) {
    if (valid && (threadIdx.x & 31) == 0) atomicAdd(&h[d], 32u);
  } else if (valid) {
    atomicAdd(&h[d], 1u);
  }
}

// Reads x[p][0..M) once: histograms of digits 0,5,6,7, sum, finiteness.
__device__ __forceinline__ void hist0_body(const double* __restrict__ x, uint32_t M, const SortBufs& b, int p) {
  __shared__ uint32_t h[4][256];
  __shared__ double red[33];
  const double* xp = x + (size_t)p * M;
  for (int i = threadIdx.x; i < 4 * 256; i += 1024) (&h[0][0])[i] = 0;
  __syncthreads();
  double s = 0.0;
  bool bad = false;
  const uint32_t mround = (M + 1023u) & ~1023u;
  for (uint32_t t = threadIdx.x; t < mround; t += 1024) {
    const bool valid = t < M;
    double v = valid ? xp[t] : 0.0;
    bad |= !is_finite_bits(v);
    s += v;
    const uint64_t k = dbl2key(v);
    if (valid) atomicAdd(&h[0][k & 255u], 1u);
    hist_add(h[1], (uint32_t)(k >> 40) & 255u, valid);
    hist_add(h[2], (uint32_t)(k >> 48) & 255u, valid);
    hist_add(h[3], (uint32_t)(k >> 56) & 255u, valid);
  }
  s = block_sum<1024>(s, red);
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(b.flag, 1);
  uint32_t* hp = b.hist + (size_t)p * 8 * 256;
  for (int i = threadIdx.x; i < 256; i += 1024) {
    hp[0 * 256 + i] = h[0][i];
    hp[5 * 256 + i] = h[1][i];
    hp[6 * 256 + i] = h[2][i];
    hp[7 * 256 + i] = h[3][i];
  }
  if (threadIdx.x == 0) {
    b.sum[p] = s;
    b.sel[p] = 0;
  }
}

__global__ void __launch_bounds__(1024) k_hist0(const double* __restrict__ x, uint32_t M, SortBufs b) {
  const int n = *b.nlist;
  for (int q = blockIdx.x; q < n; q += gridDim.x) {
    hist0_body(x, M, b, b.plist[q]);
    __syncthreads();
  }
}

// One LSD pass for digit `pass` (stable).  FIRST: source is the raw draws
// (key transform on the fly, position = index); otherwise the current buffer.
template <bool FIRST>
__device__ __forceinline__ void radix_pass_body(const double* __restrict__ x, uint32_t M, int pass,
                                                const SortBufs& b, int p) {
  __shared__ uint32_t wcnt[SORT_NW][256];
  __shared__ uint32_t base[256];
  __shared__ uint32_t hnext[256];
  __shared__ int skip;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t* hp = b.hist + ((size_t)p * 8 + pass) * 256;
  const int shift = pass * 8;
  const bool want_next = pass < 4;  // digits 5..7 were counted by k_hist0

  if (tid == 0) skip = 0;
  if (tid < 256) hnext[tid] = 0;
  __syncthreads();
  // exclusive scan of the 256-bin histogram (threads 0..255 = warps 0..7)
  uint32_t c = 0, incl = 0;
  if (tid < 256) {
    c = hp[tid];
    // digit constant over the parameter: nothing to do.  Only passes that do not
    // have to count the next digit (4..7) may be skipped.
    if (pass >= 4 && c == M) skip = 1;
    incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t n = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += n;
    }
    wcnt[0][tid] = incl;  // temp
  }
  __syncthreads();
  if (tid < 256) {
    uint32_t pre = 0;
    for (int k = 0; k < w; ++k) pre += wcnt[0][k * 32 + 31];
    base[tid] = pre + incl - c;
  }
  __syncthreads();
  if (skip) return;

  const int src = FIRST ? 0 : b.sel[p];
  const int dst = FIRST ? 0 : (src ^ 1);
  const uint64_t* kin = b.key[src] + (size_t)p * M;
  const uint32_t* iin = b.idx[src] + (size_t)p * M;
  uint64_t* kout = b.key[dst] + (size_t)p * M;
  uint32_t* iout = b.idx[dst] + (size_t)p * M;
  const double* xp = x + (size_t)p * M;

  for (uint32_t tile = 0; tile < M; tile += SORT_TILE) {
    for (int i = tid; i < SORT_NW * 256; i += SORT_NT) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    uint64_t key[SORT_IPT];
    uint32_t idx[SORT_IPT];
    uint32_t off[SORT_IPT];
    const uint32_t wbase = tile + w * (32 * SORT_IPT) + lane;
#pragma unroll
    for (int it = 0; it < SORT_IPT; ++it) {
      const uint32_t e = wbase + it * 32;
      if (e < M) {
        if (FIRST) {
          key[it] = dbl2key(xp[e]);
          idx[it] = e;
        } else {
          key[it] = kin[e];
          idx[it] = iin[e];
        }
      } else {
        key[it] = ~0ull;
        idx[it] = 0;
      }
    }
#pragma unroll
    for (int it = 0; it < SORT_IPT; ++it) {
      const uint32_t e = wbase + it * 32;
      const bool valid = e < M;
      const uint32_t d = (uint32_t)(key[it] >> shift) & 255u;
      const unsigned peers = __match_any_sync(FULL, valid ? d : (256u + lane));
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (valid && lane == leader) {
        old = wcnt[w][d];
        wcnt[w][d] = old + __popc(peers);
      }
      old = __shfl_sync(FULL, old, leader);
      off[it] = old + __popc(peers & ((1u << lane) - 1u));
      __syncwarp();
      if (want_next) hist_add(hnext, (uint32_t)(key[it] >> (shift + 8)) & 255u, valid);
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t s = base[tid];
#pragma unroll
      for (int k = 0; k < SORT_NW; ++k) {
        const uint32_t cw = wcnt[k][tid];
        wcnt[k][tid] = s;
        s += cw;
      }
      base[tid] = s;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SORT_IPT; ++it) {
      const uint32_t e = wbase + it * 32;
      if (e < M) {
        const uint32_t d = (uint32_t)(key[it] >> shift) & 255u;
        const uint32_t pos = wcnt[w][d] + off[it];
        kout[pos] = key[it];
        iout[pos] = idx[it];
      }
    }
    __syncthreads();
  }
  if (want_next && tid < 256) b.hist[((size_t)p * 8 + pass + 1) * 256 + tid] = hnext[tid];
  if (tid == 0) b.sel[p] = dst;
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_NT) k_radix_pass(const double* __restrict__ x, uint32_t M, int pass,
                                                        SortBufs b) {
  const int n = *b.nlist;
  for (int q = blockIdx.x; q < n; q += gridDim.x) {
    radix_pass_body<FIRST>(x, M, pass, b, b.plist[q]);
    __syncthreads();
  }
}

}  // namespace b200ref
