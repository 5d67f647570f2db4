// Ingestion of the reference's long draws table (store.py:79-95, convert.py:150-161):
// a Parquet / Arrow parameter column whose rows are ordered by (chain, draw) IS the device
// layout x[p][C][N].  The reference re-buckets and re-sorts the rows per parameter in Python;
// here the order is verified once for the whole table on the device, and a table whose rows
// are not in that order is permuted by one gather shared by all parameters.
#pragma once
#include "common.cuh"

namespace b200ref {

// flags: bit0 some row pair is out of (chain, draw) order, bit1 a (chain, draw) pair repeats
__global__ void k_row_order(const int64_t* __restrict__ chain, const int64_t* __restrict__ draw, int64_t rows,
                            int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int f = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < rows; i += stride) {
    const int64_t c0 = chain[i - 1], c1 = chain[i], d0 = draw[i - 1], d1 = draw[i];
    if (c1 < c0 || (c1 == c0 && d1 < d0)) f |= 1;
    if (c1 == c0 && d1 == d0) f |= 2;
  }
  f = __reduce_or_sync(FULL, f);
  if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

// out[p][r] = in[p][perm[r]]: rows are read through the permutation (gather), written coalesced
__global__ void k_gather_rows(const double* __restrict__ in, const int64_t* __restrict__ perm, int64_t P, int64_t M,
                              double* __restrict__ out) {
  const int64_t p = blockIdx.y;
  const double* ip = in + p * M;
  double* op = out + p * M;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < M; r += stride) op[r] = ip[perm[r]];
}

// Transposing load of the reference's numpy layout: `Draws.to_numpy()` gives rows = (chain, draw), columns =
// parameters, PARAMETER FASTEST (draws.py:20-29).  out[pb][r] = in[r][p0 + pb] for a block of PB parameter
// columns: 32 x 32 tiles through padded shared memory, both the global reads (32 consecutive parameters of
// a row = 256 B) and the global writes (32 consecutive rows of a parameter) are full sectors.  A TMA tensor
// map is not used here: its global stride (8 P bytes) must be a multiple of 16 B, i.e. an even P.
constexpr int TR_TILE = 32, TR_ROWS = 8;
__global__ void __launch_bounds__(TR_TILE * TR_ROWS) k_transpose_rows(const double* __restrict__ in, int64_t M, int64_t P,
                                                                      int64_t p0, int64_t PB, double* __restrict__ out) {
  __shared__ double tile[TR_TILE][TR_TILE + 1];
  const int64_t r0 = (int64_t)blockIdx.x * TR_TILE, c0 = (int64_t)blockIdx.y * TR_TILE;
  const int tx = threadIdx.x & (TR_TILE - 1), ty = threadIdx.x / TR_TILE;
#pragma unroll
  for (int j = ty; j < TR_TILE; j += TR_ROWS) {
    const int64_t r = r0 + j, c = c0 + tx;
    if (r < M && c < PB) tile[j][tx] = in[r * P + p0 + c];
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < TR_TILE; j += TR_ROWS) {
    const int64_t c = c0 + j, r = r0 + tx;
    if (c < PB && r < M) out[c * M + r] = tile[tx][j];
  }
}

}  // namespace b200ref
