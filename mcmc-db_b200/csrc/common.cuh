// Shared device helpers for the b200ref kernels (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b200ref {

constexpr unsigned FULL = 0xffffffffu;

// Checked build (-DB200REF_CHECKED, `python __graft_entry__.py --checked` -> libb200ref_checked.so): every store
// whose address depends on a scheduling-dependent atomic or on another kernel's output is bounds-checked and a
// violation traps with a message.  (compute-sanitizer is closed on the GPU pool this was developed on; the checked
// build run over tools/sanitize_smoke.py and the edge-case tests is the substitute -- profiles/r2/.)
#ifdef B200REF_CHECKED
#define B200REF_CHECK(cond, what)                                                                       \
  do {                                                                                                  \
    if (!(cond)) {                                                                                      \
      printf("b200ref CHECK failed: %s (%s:%d, block %d,%d thread %d)\n", what, __FILE__, __LINE__,     \
             blockIdx.x, blockIdx.y, threadIdx.x);                                                      \
      __trap();                                                                                         \
    }                                                                                                   \
  } while (0)
#else
#define B200REF_CHECK(cond, what) ((void)0)
#endif
// Bit 31 of a sorted element's original position (C*N < 2^31): the element belongs to a tie run and
// its code k = i + j is in kcode; an untied element at sorted position t has k = 2t + 1.
constexpr uint32_t TIED_BIT = 0x80000000u;

// Order-preserving map double -> uint64.  -0.0 is canonicalised to +0.0 first
// because the reference groups ties with Python `==` (diagnostics.py:116), under
// which -0.0 == 0.0.  Inputs are finite (checked by the first kernel).
__device__ __forceinline__ uint64_t dbl2key(double v) {
  uint64_t u = (uint64_t)__double_as_longlong(v);
  if (u == 0x8000000000000000ull) u = 0ull;
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key2dbl(uint64_t k) {
  uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}
__device__ __forceinline__ bool is_finite_bits(double v) {
  return (((uint64_t)__double_as_longlong(v) >> 52) & 0x7ffull) != 0x7ffull;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// Block-wide sum; result valid in every thread.  `red` needs >= 33 doubles.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < NT / 32) ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// Inclusive forward max-scan / reverse min-scan of one uint32 per (thread, item),
// blocked arrangement (thread t owns positions t*IPT .. t*IPT+IPT-1).
// `sm` needs NT/32 uint32.
template <int NT, int IPT>
__device__ __forceinline__ void block_scan_max_fwd(uint32_t (&v)[IPT], uint32_t* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 1; i < IPT; ++i) v[i] = max(v[i], v[i - 1]);
  uint32_t tot = v[IPT - 1];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(FULL, tot, o);
    if (lane >= o) tot = max(tot, n);
  }
  __syncthreads();
  if (lane == 31) sm[w] = tot;
  __syncthreads();
  uint32_t pre = 0;  // exclusive prefix over previous warps
  for (int k = 0; k < w; ++k) pre = max(pre, sm[k]);
  uint32_t prev = __shfl_up_sync(FULL, tot, 1);
  if (lane == 0) prev = 0;
  pre = max(pre, prev);
#pragma unroll
  for (int i = 0; i < IPT; ++i) v[i] = max(v[i], pre);
}

template <int NT, int IPT>
__device__ __forceinline__ void block_scan_min_rev(uint32_t (&v)[IPT], uint32_t* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = IPT - 2; i >= 0; --i) v[i] = min(v[i], v[i + 1]);
  uint32_t tot = v[0];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_down_sync(FULL, tot, o);
    if (lane + o < 32) tot = min(tot, n);
  }
  __syncthreads();
  if (lane == 0) sm[w] = tot;
  __syncthreads();
  uint32_t pre = 0xffffffffu;
  for (int k = w + 1; k < NT / 32; ++k) pre = min(pre, sm[k]);
  uint32_t nxt = __shfl_down_sync(FULL, tot, 1);
  if (lane == 31) nxt = 0xffffffffu;
  pre = min(pre, nxt);
#pragma unroll
  for (int i = 0; i < IPT; ++i) v[i] = min(v[i], pre);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t_start = clock64();
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if ((spin & 1023u) == 1023u && clock64() - t_start > 4000000000ll) {  // ~2 s
      printf("b200ref: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// 1-D TMA bulk copy global -> shared (16-byte aligned addresses, size a multiple of 16);
// completion is signalled on `bar` as `bytes` transferred.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 2-D TMA tile load (tensor map in kernel-parameter space); completion is signalled on `bar`.  SASS: UTMALDG.2D.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// orders earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

}  // namespace b200ref
