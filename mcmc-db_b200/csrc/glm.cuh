// Batched GLM log-density: logistic regression  lp_s = sum_n [ y_n * eta_ns - log(1 + exp(eta_ns)) ],
// eta = X beta_s  (Stan: y ~ bernoulli_logit(X * beta), e.g. packaged wells_data-wells_dist.stan:9-11).
// The reference ships the Stan text and data of its posteriordb-style models but no
// log-density code: this is an extension; parity is pinned only against
// an fp64 numpy restatement of the Stan likelihood kept with the tests, tolerance 1e-6 relative.
//
// tcgen05 design (sm_100a):
//   * the contraction eta[S x N] = beta[S x D] * X[N x D]^T runs on the 5th-gen tensor cores;
//     fp64 inputs are split into bf16 hi + lo parts (16 significant bits) and three MMAs
//     (hi*hi + hi*lo + lo*hi) accumulate in fp32 in TMEM -- |d eta| ~ 1e-5 |eta|, which moves lp by
//     < 1e-7 relative because the errors average out over the N terms.  The NORMAL family squares the
//     residual y - eta, which amplifies d eta by |eta| / |y - eta|: there only the draws' DEVIATIONS from their
//     mean are contracted (y - X beta_s = (y - X mean(beta)) - X (beta_s - mean(beta)); the first bracket is one
//     fp64 matrix-vector product per call, folded into the response) -- exact algebra, no epilogue cost, and the
//     error of the split then scales with the spread of the draws instead of with eta;
//   * UMMA tile 128 draws x 256 data rows x K 16, cta_group::1; DRAWS are the M dimension so
//     that every TMEM lane (= one epilogue thread) owns one draw and can keep a private
//     running sum over data rows -- no cross-thread reduction in the epilogue;
//   * operands arrive by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a 2-stage
//     shared-memory ring (96 KB per stage: beta hi/lo 128x64, X hi/lo 256x64 bf16),
//     full/empty mbarriers, one elected thread issues the MMAs, tcgen05.commit frees stages;
//   * two 256-column accumulators (all 512 TMEM columns) ping-pong between the MMA warp and
//     8 epilogue warps, which read eta with tcgen05.ld.32x32b.x32 and apply the fused
//     likelihood y*eta - softplus(eta) in fp32 (ex2/lg2 MUFU), summing 32 columns in fp32 and
//     then into an fp64 accumulator: the 26 GB eta matrix never exists in memory.
//   * persistent CTAs; a work item = (draw tile, slice of the data rows); the <= NSPLIT partial
//     sums per draw are combined by a tiny kernel in a fixed order.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"

namespace b200ref {

constexpr int GLM_BM = 128;   // draws per tile   (UMMA M)
constexpr int GLM_BN = 256;   // data rows per tile (UMMA N)
// bf16 elements per k-block = one swizzle row: 64 (128-byte swizzle, 2 stages of 96 KB) or 32 (64-byte swizzle, 4 stages
// of 48 KB).  Same shared memory, but with 4 stages the producer runs three k-blocks ahead of the MMAs instead of one:
// a stage's MMAs take ~770 / ~1540 cycles, a TMA round trip for it about as long as the longer of the two.
#ifndef B200REF_GLM_BK
#define B200REF_GLM_BK 32
#endif
constexpr int GLM_BK = B200REF_GLM_BK;
static_assert(GLM_BK == 64 || GLM_BK == 32, "one 128-byte or 64-byte swizzle row per k-block");
constexpr int GLM_KPAD = 64;  // the contraction length is padded to a multiple of this (workspace layout)
constexpr int GLM_STAGES = (GLM_BK == 64) ? 2 : 4;
// Epilogue warps: 8 (two per TMEM lane quadrant, 128 accumulator columns each) or 16 (four per quadrant, 64 columns
// each).  The fused likelihood (two MUFU + ~10 ALU instructions per element) is what the MMAs wait for: with 4 warps
// per scheduler instead of 2 the MUFU and ALU pipes overlap.
#ifndef B200REF_GLM_EPI_WARPS
#define B200REF_GLM_EPI_WARPS 16
#endif
constexpr int GLM_EPI_WARPS = B200REF_GLM_EPI_WARPS;
static_assert(GLM_EPI_WARPS == 8 || GLM_EPI_WARPS == 16, "2 or 4 epilogue warps per TMEM lane quadrant");
constexpr int GLM_EPI_WARP0 = 4;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4.. epilogue
constexpr int GLM_EPI_THREADS = GLM_EPI_WARPS * 32;
constexpr int GLM_NT = GLM_EPI_WARP0 * 32 + GLM_EPI_THREADS;
constexpr int GLM_EPI_COLS = GLM_BN / (GLM_EPI_WARPS / 4);  // accumulator columns per epilogue warp
constexpr int GLM_PART_SLOTS = GLM_EPI_WARPS / 8;           // partial-sum slots per row slice: exactly two adds each
constexpr uint32_t GLM_A_BYTES = GLM_BM * GLM_BK * 2;  // 16 KB
constexpr uint32_t GLM_B_BYTES = GLM_BN * GLM_BK * 2;  // 32 KB
constexpr uint32_t GLM_STAGE_BYTES = 2 * GLM_A_BYTES + 2 * GLM_B_BYTES;  // 96 KB
constexpr size_t GLM_SMEM_BYTES = 1024 + (size_t)GLM_STAGES * GLM_STAGE_BYTES + 4 * GLM_BN * 4 + 256;

// Likelihood families (Stan `~` semantics: terms that do not depend on the parameters are dropped):
//   0 bernoulli_logit  sum_n y eta - softplus(eta)
//   1 binomial_logit   sum_n y eta - trials softplus(eta)
//   2 poisson_log      sum_n y eta - exp(eta)
//   3 normal           sum_n -(y - eta)^2        (the caller scales by 1 / (2 sigma_s^2) and adds -N log sigma_s)
constexpr int GLM_BERNOULLI_LOGIT = 0, GLM_BINOMIAL_LOGIT = 1, GLM_POISSON_LOG = 2, GLM_NORMAL = 3;

struct GlmParams {
  int64_t N, S;
  int Dpad;            // multiple of 64
  int n_row_tiles;     // ceil(N / 256)
  int n_draw_tiles;    // ceil(S / 128)
  int nsplit;          // row-tile slices per draw tile
  const float* yf;     // [n_row_tiles * 256] y as float, padded with 0
  const float* af;     // [n_row_tiles * 256] binomial trials as float (binomial_logit), else unused
  double* lp_part;     // [nsplit * GLM_PART_SLOTS][S], zero-initialised
};

// ---- PTX wrappers -------------------------------------------------------------------------
// (smem_u32 / mbar_* live in common.cuh)
// (tma_load_2d lives in common.cuh)
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, 128 x 256 x 16, bf16 in, fp32 accumulate
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
  return r;
}

// K-major, 128-byte-swizzled operand tile (rows of 64 bf16 = 128 B, 8-row groups 1024 B apart):
// start address >> 4 | LBO 1 | SBO 64 (1024 B >> 4) | version 1 (Blackwell) | layout SWIZZLE_128B (2)
// (64-byte swizzle, GLM_BK = 32: rows of 64 B, 8-row groups 512 B apart: SBO 32, layout SWIZZLE_64B (4))
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(GLM_BK == 64 ? 64 : 32) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(GLM_BK == 64 ? 2 : 4) << 61;
  return d;
}
// kind::f16: D fp32 (bits 4-5 = 1), A bf16 (bits 7-9 = 1), B bf16 (bits 10-12 = 1), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28
constexpr uint32_t GLM_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GLM_BN >> 3) << 17) |
                               ((uint32_t)(GLM_BM >> 4) << 24);

// ---- helper kernels -------------------------------------------------------------------------
// fp64 [rows][D] (minus an optional per-column centre) -> bf16 hi, lo [rows][Dpad] (zero padded):
// x ~ hi + lo with 16 significant bits
__global__ void k_glm_split(const double* __restrict__ src, const double* __restrict__ center, int64_t rows, int D, int Dpad,
                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Dpad) return;
  const int64_t r = i / Dpad;
  const int d = (int)(i % Dpad);
  double x = (d < D) ? src[r * D + d] : 0.0;
  if (center && d < D) x -= center[d];
  const __nv_bfloat16 h = __double2bfloat16(x);
  const double rem = x - (double)__bfloat162float(h);
  hi[i] = h;
  lo[i] = __double2bfloat16(rem);
}
// mean draw: center[d] = mean_s beta[s][d]  (one CTA per column, fixed reduction tree: deterministic)
__global__ void __launch_bounds__(256) k_glm_center(const double* __restrict__ beta, int64_t S, int D, double* __restrict__ center) {
  __shared__ double red[33];
  const int d = blockIdx.x;
  double a = 0.0;
  for (int64_t s = threadIdx.x; s < S; s += 256) a += beta[s * D + d];
  a = block_sum<256>(a, red);
  if (threadIdx.x == 0) center[d] = a / (double)S;
}
// response minus the linear predictor of the mean draw, in fp64, stored as float (padding rows 0): one warp per row
__global__ void __launch_bounds__(256) k_glm_y_centered(const double* __restrict__ y, const double* __restrict__ X,
                                                        const double* __restrict__ center, int64_t N, int D, int64_t Npad,
                                                        float* __restrict__ yf) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= Npad) return;
  double a = 0.0;
  if (row < N)
    for (int d = lane; d < D; d += 32) a = fma(X[row * D + d], center[d], a);
  a = warp_sum(a);
  if (lane == 0) yf[row] = (row < N) ? (float)(y[row] - a) : 0.0f;
}
__global__ void k_glm_y(const double* __restrict__ y, int64_t N, int64_t Npad, float* __restrict__ yf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Npad) yf[i] = (i < N) ? (float)y[i] : 0.0f;
}
// sigma != null (normal family): part holds -SSR_s;  lp = -N log sigma_s + part / (2 sigma_s^2)
__global__ void k_glm_reduce(const double* __restrict__ part, int nsplit, int64_t S, const double* __restrict__ sigma,
                             int64_t N, double* __restrict__ lp) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double a = 0.0;
  for (int h = 0; h < nsplit; ++h) a += part[(size_t)h * S + s];
  if (sigma) {
    const double sg = sigma[s];
    a = -(double)N * log(sg) + 0.5 * a / (sg * sg);
  }
  lp[s] = a;
}

// ---- the fused GEMM + likelihood kernel -------------------------------------------------------
template <int FAM>
__global__ void __launch_bounds__(GLM_NT, 1)
k_glm_lp(const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
               const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl, GlmParams prm) {
  extern __shared__ unsigned char glm_smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)glm_smem_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B: 1 KB
  unsigned char* stage_base = smem;
  float* ys = (float*)(smem + (size_t)GLM_STAGES * GLM_STAGE_BYTES);  // [2][256] y, then [2][256] trials
  float* as = ys + 2 * GLM_BN;
  uint64_t* bars = (uint64_t*)(ys + 4 * GLM_BN);
  uint64_t* full = bars;                         // [GLM_STAGES]  TMA -> MMA
  uint64_t* empty = bars + GLM_STAGES;           // [GLM_STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * GLM_STAGES;       // [2]  MMA -> epilogue
  uint64_t* tempty = bars + 2 * GLM_STAGES + 2;  // [2]  epilogue -> MMA
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * GLM_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = prm.Dpad / GLM_BK;
  const int tiles_per_split = (prm.n_row_tiles + prm.nsplit - 1) / prm.nsplit;
  const int nitems = prm.n_draw_tiles * prm.nsplit;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < GLM_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], GLM_EPI_THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int dt = item / prm.nsplit, h = item % prm.nsplit;
        const int t0 = h * tiles_per_split, t1 = min(t0 + tiles_per_split, prm.n_row_tiles);
        for (int rt = t0; rt < t1; ++rt) {
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char* sb = stage_base + (size_t)stage * GLM_STAGE_BYTES;
            mbar_expect_tx(&full[stage], GLM_STAGE_BYTES);
            tma_load_2d(sb, &map_bh, &full[stage], kb * GLM_BK, dt * GLM_BM);
            tma_load_2d(sb + GLM_A_BYTES, &map_bl, &full[stage], kb * GLM_BK, dt * GLM_BM);
            tma_load_2d(sb + 2 * GLM_A_BYTES, &map_xh, &full[stage], kb * GLM_BK, rt * GLM_BN);
            tma_load_2d(sb + 2 * GLM_A_BYTES + GLM_B_BYTES, &map_xl, &full[stage], kb * GLM_BK, rt * GLM_BN);
            if (++stage == GLM_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      int stage = 0, abuf = 0;
      uint32_t phase = 0, aphase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int h = item % prm.nsplit;
        const int t0 = h * tiles_per_split, t1 = min(t0 + tiles_per_split, prm.n_row_tiles);
        for (int rt = t0; rt < t1; ++rt) {
          mbar_wait(&tempty[abuf], aphase ^ 1);  // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)abuf * GLM_BN;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sb = smem_u32(stage_base + (size_t)stage * GLM_STAGE_BYTES);
            const uint64_t dah = umma_desc_k_sw128(sb), dal = umma_desc_k_sw128(sb + GLM_A_BYTES);
            const uint64_t dbh = umma_desc_k_sw128(sb + 2 * GLM_A_BYTES);
            const uint64_t dbl = umma_desc_k_sw128(sb + 2 * GLM_A_BYTES + GLM_B_BYTES);
#pragma unroll
            for (int k = 0; k < GLM_BK / 16; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);  // 16 bf16 = 32 bytes = 2 x 16-byte units
              tc_mma_bf16(tmem_d, dah + adv, dbh + adv, GLM_IDESC, (kb | k) ? 1u : 0u);
              tc_mma_bf16(tmem_d, dah + adv, dbl + adv, GLM_IDESC, 1u);
              tc_mma_bf16(tmem_d, dal + adv, dbh + adv, GLM_IDESC, 1u);
            }
            tc_commit(&empty[stage]);  // frees the smem stage once these MMAs retire
            if (++stage == GLM_STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit(&tfull[abuf]);  // accumulator complete
          if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
      }
    }
  } else if (warp >= GLM_EPI_WARP0) {
    // ================= epilogue: eta -> y*eta - softplus(eta), summed over data rows =================
    const int ew = warp - GLM_EPI_WARP0;       // 0 .. GLM_EPI_WARPS-1
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int cpart = ew >> 2;                 // which GLM_EPI_COLS columns
    const int et = threadIdx.x - GLM_EPI_WARP0 * 32;  // 0 .. GLM_EPI_THREADS-1
    const int row = quad * 32 + lane;          // draw within the tile
    int abuf = 0;
    uint32_t aphase = 0;
    const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int dt = item / prm.nsplit, h = item % prm.nsplit;
      const int t0 = h * tiles_per_split, t1 = min(t0 + tiles_per_split, prm.n_row_tiles);
      double acc = 0.0;
      for (int rt = t0; rt < t1; ++rt) {
        // stage y for this tile's 256 data rows (one element per thread of the first 8 epilogue warps)
        if (et < GLM_BN) {
          ys[abuf * GLM_BN + et] = prm.yf[(size_t)rt * GLM_BN + et];
          if (FAM == GLM_BINOMIAL_LOGIT) as[abuf * GLM_BN + et] = prm.af[(size_t)rt * GLM_BN + et];
        }
        mbar_wait(&tfull[abuf], aphase);
        tc_fence_after();
        asm volatile("bar.sync 1, %0;" ::"n"(GLM_EPI_THREADS) : "memory");
        const int64_t nvalid = prm.N - (int64_t)rt * GLM_BN;  // rows of this tile that exist
        // y and the binomial trials of the tile's data rows, 4 rows per LDS.128 broadcast.  (explicit ld.shared: the
        // 1 KB round-up of the dynamic shared-memory base goes through an integer, after which the compiler only
        // knows a generic pointer and emits LD.E instead of LDS)
        const uint32_t yt4 = smem_u32(ys + abuf * GLM_BN), at4 = smem_u32(as + abuf * GLM_BN);
        // MASK: only the last row tile of the design has rows that do not exist (their terms are dropped)
        auto tile_terms = [&](auto mask_tag) {
          constexpr bool MASK = decltype(mask_tag)::value;
#pragma unroll 1
          for (int cc = 0; cc < GLM_EPI_COLS / 32; ++cc) {
            const int col0 = cpart * GLM_EPI_COLS + cc * 32;
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(abuf * GLM_BN + col0), v);
            float part = 0.0f;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 y4 = lds128(yt4 + (uint32_t)(col0 + 4 * j4) * 4u);
              float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (FAM == GLM_BINOMIAL_LOGIT) a4 = lds128(at4 + (uint32_t)(col0 + 4 * j4) * 4u);
              const float yy[4] = {y4.x, y4.y, y4.z, y4.w}, aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int j = 4 * j4 + q;
                const float eta = __uint_as_float(v[j]);
                float term;
                if (FAM == GLM_BERNOULLI_LOGIT || FAM == GLM_BINOMIAL_LOGIT) {
                  const float a = fabsf(eta);
                  float e, l;
                  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-a * LOG2E));
                  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));
                  const float sp = fmaxf(eta, 0.0f) + l * LN2;  // softplus
                  term = yy[q] * eta - ((FAM == GLM_BINOMIAL_LOGIT) ? aa[q] * sp : sp);
                } else if (FAM == GLM_POISSON_LOG) {
                  float e;
                  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(eta * LOG2E));
                  term = yy[q] * eta - e;
                } else {
                  const float r = yy[q] - eta;
                  term = -r * r;
                }
                if (MASK && col0 + j >= nvalid) term = 0.0f;
                part += term;
              }
            }
            acc += (double)part;
          }
        };
        if (nvalid >= GLM_BN) tile_terms(std::false_type{});
        else tile_terms(std::true_type{});
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[abuf]);
        if (++abuf == 2) { abuf = 0; aphase ^= 1; }
      }
      const int64_t s = (int64_t)dt * GLM_BM + row;
      // two warps share a slot: exactly two adds, order-free
      if (s < prm.S) atomicAdd(&prm.lp_part[((size_t)h * GLM_PART_SLOTS + (cpart >> 1)) * prm.S + s], acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace b200ref
