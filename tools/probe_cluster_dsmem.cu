// Hardware probe: cluster residency + DSMEM / smem-atomic / L2-gather throughput on B200.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdint.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, uint64_t v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;" :: "r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_cluster_u64(uint32_t addr) {
  uint64_t v; asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory"); return v;
}
__device__ __forceinline__ uint32_t atom_cluster_add(uint32_t addr, uint32_t v) {
  uint32_t o; asm volatile("atom.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(v) : "memory"); return o;
}
__device__ __forceinline__ void red_cluster_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared::cluster.add.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}

// mode 0: scattered remote st.u64 ; 1: scattered remote ld.u64 ; 2: remote atom.add w/ return ; 3: remote red.add
// 4: local smem atomicAdd w/ return (spread) ; 5: local scattered STS.64 ; 6: local scattered LDS.64
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_dsmem(int slots_per_cta, int iters, unsigned long long* out, long long* cyc) {
  extern __shared__ __align__(16) unsigned char sm[];
  cg::cluster_group cl = cg::this_cluster();
  const uint32_t G = cl.num_blocks(), me = cl.block_rank();
  uint64_t* buf = (uint64_t*)sm;
  for (int i = threadIdx.x; i < slots_per_cta; i += blockDim.x) buf[i] = i;
  cl.sync();
  const uint32_t base = smem_u32(buf);
  uint32_t h = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
  uint64_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h = h * 1664525u + 1013904223u;
      const uint32_t g = (h >> 8) % ((uint32_t)slots_per_cta * G);
      const uint32_t dst = g / slots_per_cta, slot = g % slots_per_cta;
      if (MODE == 0) st_cluster_u64(mapa(base + slot * 8, dst), (uint64_t)h);
      else if (MODE == 1) acc += ld_cluster_u64(mapa(base + slot * 8, dst));
      else if (MODE == 2) acc += atom_cluster_add(mapa(base + slot * 8, dst), 1u);
      else if (MODE == 3) red_cluster_add(mapa(base + slot * 8, dst), 1u);
      else if (MODE == 4) acc += atomicAdd((uint32_t*)&buf[slot], 1u);
      else if (MODE == 5) buf[slot] = h;
      else if (MODE == 6) acc += buf[slot];
    }
  }
  cl.sync();
  const long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; }
  if (acc == 0x1234567) out[0] = acc;
}

// random 8-byte gathers from a table in L2 (mode 0) and window-local gathers (mode 1: 32 lanes inside a 256-entry window)
__global__ void __launch_bounds__(1024, 1) k_gather(const double* tab, uint32_t n, int iters, int mode, double* out, long long* cyc) {
  uint32_t h = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 777u;
  double acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h = h * 1664525u + 1013904223u;
      uint32_t idx;
      if (mode == 0) idx = (h >> 4) % n;
      else {
        const uint32_t wbase = ((blockIdx.x * 32 + (threadIdx.x >> 5)) * 7919u + it * 8 + j) * 256u % (n - 256);
        idx = wbase + ((h >> 9) & 255u);
      }
      acc += __ldg(tab + idx);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 1.2345) out[0] = acc;
}

template <int MODE>
void run_dsmem(int G, int smem_bytes, const char* name) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  CK(cudaFuncSetAttribute(k_dsmem<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  if (G > 8) CK(cudaFuncSetAttribute(k_dsmem<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem_bytes; cfg.attrs = at; cfg.numAttrs = 1;
  int ncl = 0;
  cfg.gridDim = dim3(G * 64);
  cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, k_dsmem<MODE>, &cfg);
  if (e != cudaSuccess) { printf("%s G=%d occupancy query: %s\n", name, G, cudaGetErrorString(e)); cudaGetLastError(); return; }
  if (ncl < 1) { printf("%s G=%d: 0 clusters\n", name, G); return; }
  cfg.gridDim = dim3(G * ncl);
  unsigned long long* out; long long* cyc;
  CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&cyc, 8 * G * ncl));
  const int iters = 256, slots = smem_bytes / 8;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(a);
    e = cudaLaunchKernelEx(&cfg, k_dsmem<MODE>, slots, iters, out, cyc);
    cudaEventRecord(b);
    if (e != cudaSuccess) { printf("%s G=%d launch: %s\n", name, G, cudaGetErrorString(e)); cudaGetLastError(); return; }
    CK(cudaDeviceSynchronize());
  }
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long hc[1024]; CK(cudaMemcpy(hc, cyc, 8 * G * ncl, cudaMemcpyDeviceToHost));
  long long mx = 0; double av = 0; for (int i = 0; i < G * ncl; ++i) { mx = hc[i] > mx ? hc[i] : mx; av += hc[i]; } av /= G * ncl;
  const double ops = 1024.0 * iters * 8;
  printf("%-22s G=%2d smem=%6d clusters=%3d (SMs %3d): %.3f ms, cycles/CTA avg %.0f max %lld -> %.3f cyc per op per SM (%.2f ops/cyc/SM)\n",
         name, G, smem_bytes, ncl, G * ncl, ms, av, mx, av / ops, ops / av);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  printf("%s SMs=%d smem/block optin=%zu L2=%d MB\n", pr.name, pr.multiProcessorCount, pr.sharedMemPerBlockOptin, pr.l2CacheSize >> 20);
  for (int G : {1, 2, 4, 8, 16}) {
    for (int sb : {100 * 1024, 200 * 1024, 227 * 1024}) {
      cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      CK(cudaFuncSetAttribute(k_dsmem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb));
      if (G > 8) CK(cudaFuncSetAttribute(k_dsmem<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = sb; cfg.attrs = at; cfg.numAttrs = 1; cfg.gridDim = dim3(G * 64);
      int ncl = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, k_dsmem<0>, &cfg);
      printf("maxActiveClusters G=%2d smem=%6d: %d (%s)\n", G, sb, ncl, cudaGetErrorString(e)); cudaGetLastError();
    }
  }
  const int sb = 128 * 1024;
  for (int G : {8, 16}) {
    run_dsmem<0>(G, sb, "remote st.u64 scatter");
    run_dsmem<1>(G, sb, "remote ld.u64 gather");
    run_dsmem<2>(G, sb, "remote atom.add ret");
    run_dsmem<3>(G, sb, "remote red.add");
  }
  run_dsmem<4>(1, sb, "local atomicAdd ret");
  run_dsmem<5>(1, sb, "local STS.64 scatter");
  run_dsmem<6>(1, sb, "local LDS.64 gather");
  // L2 gathers
  const uint32_t n = 524288;  // 4 MB table
  double* tab; CK(cudaMalloc(&tab, n * 8)); CK(cudaMemset(tab, 0, n * 8));
  double* out; long long* cyc; CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&cyc, 8 * 148));
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) { k_gather<<<148, 1024>>>(tab, n, 256, mode, out, cyc); CK(cudaDeviceSynchronize()); }
    long long hc[148]; CK(cudaMemcpy(hc, cyc, 8 * 148, cudaMemcpyDeviceToHost));
    double av = 0; for (int i = 0; i < 148; ++i) av += hc[i]; av /= 148;
    printf("L2 gather mode %d (0 random, 1 window-256): %.3f cyc per op per SM\n", mode, av / (1024.0 * 256 * 8));
  }
  return 0;
}
