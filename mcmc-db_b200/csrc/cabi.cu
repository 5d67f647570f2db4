// C-ABI of libb200ref.so (declared in include/b200ref.h): host orchestration of the
// sm_100a kernels for mcmc-ref's diagnostics / stats / compare path.
#include "../../include/b200ref.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <mutex>
#include <vector>

#include "as241.cuh"
#include "bucket.cuh"
#include "chain.cuh"
#include "common.cuh"
#include "glm.cuh"
#include "ingest.cuh"
#include "rank.cuh"
#include "sort.cuh"
#include "synth.cuh"

using namespace b200ref;

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return fail(B200REF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),  \
                  __FILE__, __LINE__);                                                       \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// zz [rows][8 pairs] as a 2-D fp64 tensor (row = 16 doubles = 128 B = one swizzle row); box = M2_BOX rows: k_moments0t
int make_map_pairs(CUtensorMap* map, const void* zz, uint64_t rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(B200REF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {16, rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {16, (cuuint32_t)M2_BOX};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(zz), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200REF_ERR_CUDA, "cuTensorMapEncodeTiled (score pairs) failed with code %d", (int)r);
  return B200REF_OK;
}

// ---- optional per-kernel timing (bench.py's roofline block): CUDA events recorded on the
// launching stream around every kernel while enabled.
enum KernelId { K_ZTABLE, K_SAMPLE, K_BHIST, K_BSCAN, K_PARTITION, K_FINISH, K_FINISHB, K_HIST0, K_RADIX, K_RANKC, K_PREP,
                K_MPART, K_FOLD, K_CHAIN, K_DECIDE, K_FINALIZE, K_SUMMARY, K_COMPARE, K_CHECKS, K_SYNTH, K_GLM_PREP, K_GLM, K_INGEST, K_COUNT };
const char* const kKernelNames[K_COUNT] = {"k_ztable", "k_sample", "k_bhist", "k_bscan", "k_partition", "k_finish", "k_finish_big",
                                           "k_hist0", "k_radix_pass", "k_rank_codes", "k_param_prep",
                                           "k_merge_partition", "k_fold", "k_moments", "k_decide", "k_finalize",
                                           "k_summary", "k_compare", "k_checks", "k_synth", "k_glm_prep", "k_glm_lp", "k_ingest"};
struct ProfRec { int id; cudaEvent_t a, b; };
bool g_prof = false;
std::vector<ProfRec> g_prof_recs;
std::mutex g_prof_mu;
double g_prof_ms[K_COUNT];
uint64_t g_prof_n[K_COUNT];

// returns the index of the record this launch owns (-1: profiling off), so that two host threads
// launching at once each close their own record
inline long prof_begin(int id, cudaStream_t st) {
  if (!g_prof) return -1;
  ProfRec r;
  r.id = id;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_recs.push_back(r);
  return (long)g_prof_recs.size() - 1;
}
inline void prof_end(long rec, cudaStream_t st) {
  if (rec < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if ((size_t)rec < g_prof_recs.size()) cudaEventRecord(g_prof_recs[(size_t)rec].b, st);
}

#define LAUNCH(id, st, ...)                                 \
  do {                                                      \
    const long prof_rec_ = prof_begin((id), (st));          \
    __VA_ARGS__;                                            \
    prof_end(prof_rec_, (st));                              \
    g_launches.fetch_add(1, std::memory_order_relaxed);     \
  } while (0)

// Side stream for the (rare) LSD fallback, one per host thread.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
constexpr int MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < MAX_DEVICES) ? dev : 0;
}
// one pair of helper streams per (host thread, device): a process may drive several GPUs
SideStream& side_stream(int slot = 0) {
  thread_local SideStream pool[MAX_DEVICES][2];
  SideStream& ss = pool[current_device()][slot & 1];
  if (!ss.s) {
    cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming);
  }
  return ss;
}

constexpr size_t ALIGN = 256;
inline size_t up(size_t v) { return (v + ALIGN - 1) / ALIGN * ALIGN; }

// Long chains: when a parameter's zz row (16 B per draw) is far beyond what L2 can merge into full sectors, the
// score scatter goes through at most 32 destination windows of >= 2^20 positions (16 MB of zz; 2^19 / 2^20 / 2^21 measured
// 14.7 / 14.5 / 15.7 ms for k_fold + k_unbin on config 4, direct scatter 17.9 ms); 0 = direct scatter.
#ifndef B200REF_BIN_MIN_SHIFT
#define B200REF_BIN_MIN_SHIFT 20
#endif
inline int fold_bin_shift(int64_t M) {
  if (M <= ((int64_t)3 << 20)) return 0;
  int lg = 0;
  while (((int64_t)1 << lg) < M) ++lg;
  return std::max(B200REF_BIN_MIN_SHIFT, lg - 5);
}

// Single-pass partition (k_part1, narrow path): every coarse bucket owns a slack-padded slab of the parameter's
// record region (capacities: k_sample).  0 at compile time = the two-pass partition (k_bhist + k_bscan + k_partition).
#ifndef B200REF_SINGLE_PASS
#define B200REF_SINGLE_PASS 1
#endif
// Round 0 of the chain statistics by the tile variant (k_moments0t) whenever the chain length allows (N % 8 == 0)
#ifndef B200REF_MOMENTS_TILE
#define B200REF_MOMENTS_TILE 1
#endif
// Round 0 covers 48 instead of 16 lags for the parameters k_predict_sticky marks (tile variant only)
#ifndef B200REF_STICKY_ROUND0
#define B200REF_STICKY_ROUND0 1
#endif
inline bool use_slabs(int64_t M) { return B200REF_SINGLE_PASS != 0 && choose_sub(M) == 1; }
// records of one parameter's slab region: 3 x the mean per bucket + 64 (k_sample), + room for pure buckets
inline size_t slab_records(int64_t M) { return use_slabs(M) ? (size_t)5 * (size_t)M + 32768 : (size_t)M; }

// Workspace carve-up for one batch of PB parameters with M pooled draws each.
struct Layout {
  size_t ztab, zodd, flag, header;  // shared part
  size_t pkey, skey, pidx, sidx, kcode, zz, rzf, bcur, fn, cnt, boff, cur, spart, fallback, nfb, plist, nbigb, biglist, hist, sel, sum, mean, sd, med, m0, apos,
      crhat, cess, clag, cpart, csticky, cpartr, cmean, csum, chead, ctail, cvarhat, crhosum, cgpend, cgprev, cwl0, cwl1, cnwl, cgbar, per_batch;
  uint32_t ntiles, nchunks;
  void compute(int64_t PB, int64_t M, bool diag, int64_t C = 1, int64_t N = 1) {
    ntiles = (uint32_t)((M + RANK_TILE - 1) / RANK_TILE);
    nchunks = (uint32_t)((M + CHUNK_P - 1) / CHUNK_P);
    size_t o = 0;
    ztab = o; o += up(2 * (size_t)M * sizeof(double));
    zodd = o; o += up((size_t)M * sizeof(double));
    flag = o; o += up(sizeof(int));
    header = o;
    size_t b = 0;
    const size_t n = (size_t)PB * (size_t)M;
    const size_t nrec = (size_t)PB * slab_records(M);  // partitioned records (slack-padded slabs on the single-pass path)
    pkey = b; b += up(nrec * 8);
    skey = b; b += up(n * 8);
    pidx = b; b += up(nrec * 4);
    sidx = b; b += up(n * 4);
    kcode = b; b += up(n * 4);
    zz = b; b += diag ? up(n * 16) : 0;
    const bool binned = diag && fold_bin_shift(M) != 0;  // long chains: record segments for the two-step scatter
    rzf = b; b += binned ? up(n * 8) : 0;
    bcur = b; b += binned ? up((size_t)PB * 32 * 4) : 0;
    fn = b; b += up((size_t)PB * sizeof(BucketFn));
    const bool wide = choose_sub(M) > 1;  // long chains: sub-bucketed coarse buckets, no per-chunk counts
    const bool slabs = use_slabs(M);
    cnt = b; b += (wide || slabs) ? 0 : up((size_t)PB * nchunks * NB1_MAX * 4);
    boff = b; b += up((size_t)PB * (wide ? NBW_MAX + 1 : NB1_MAX + 1) * 4);
    cur = b; b += wide ? up((size_t)PB * NBW_MAX * 4) : (slabs ? up((size_t)PB * NB1_MAX * 4) : 0);
    spart = b; b += up((size_t)PB * nchunks * 2 * 8);
    fallback = b; b += up((size_t)PB);
    nfb = b; b += up(sizeof(int));
    nbigb = b; b += up(sizeof(int));
    biglist = b; b += up((size_t)PB * BIGLIST_PER_PARAM * 2 * 4);
    plist = b; b += up((size_t)PB * 4);
    hist = b; b += up((size_t)PB * 8 * 256 * 4);
    sel = b; b += up((size_t)PB * 4);
    sum = b; b += up((size_t)PB * 8);
    mean = b; b += up((size_t)PB * 8);
    sd = b; b += up((size_t)PB * 8);
    med = b; b += up((size_t)PB * 8);
    m0 = b; b += up((size_t)PB * 4);
    apos = b; b += up((size_t)PB * (ntiles + 1) * 4);
    crhat = b; b += up((size_t)PB * 2 * 8);
    cess = b; b += up((size_t)PB * 2 * 8);
    clag = b; b += up((size_t)PB * 2 * 8);
    if (diag) {
      const ChainGeom g = chain_geom(C, N);
      const size_t units = (size_t)C * g.nunits;
      cpart = b; b += up((size_t)PB * units * 2 * MO_VALS * 8);
      csticky = b; b += up((size_t)PB);
      cpartr = b; b += up((size_t)2 * PB * units * LAGR * 8);
      cmean = b; b += up((size_t)2 * PB * C * 8);
      csum = b; b += up((size_t)2 * PB * C * 8);
      chead = b; b += up((size_t)2 * PB * C * 8);
      ctail = b; b += up((size_t)2 * PB * C * 8);
      cvarhat = b; b += up((size_t)2 * PB * 8);
      crhosum = b; b += up((size_t)2 * PB * 8);
      cgpend = b; b += up((size_t)2 * PB * 8);
      cgprev = b; b += up((size_t)2 * PB * 8);
      cwl0 = b; b += up((size_t)2 * PB * 4);
      cwl1 = b; b += up((size_t)2 * PB * 4);
      cnwl = b; b += up(2 * sizeof(int));
      cgbar = b; b += up(sizeof(unsigned int));
    }
    per_batch = b;
  }
};

// Largest batch (<= P) whose layout fits in ws_bytes; 0 if not even one parameter fits.
int64_t fit_batch(int64_t P, int64_t M, bool diag, size_t ws_bytes, int64_t cap, int64_t C, int64_t N) {
  int64_t hi = std::min<int64_t>(P, cap), lo = 0;
  Layout l;
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo + 1) / 2;
    l.compute(mid, M, diag, C, N);
    if (l.header + l.per_batch <= ws_bytes) lo = mid; else hi = mid - 1;
  }
  return lo;
}

constexpr int64_t BATCH_CAP = 4096;
constexpr int MAX_CHAINS = 4096;  // k_decide0 keeps 3 doubles per chain in shared memory
constexpr int FALLBACK_GRID = 296;  // CTAs striding over the fallback list (2 per SM)

// CTAs of the persistent rounds kernel: as many as are resident at once (cooperative launch), per device.
int rounds_grid() {
  static int grid_dev[MAX_DEVICES] = {};
  const int dev = current_device();
  if (grid_dev[dev] == 0) {
    int per_sm = 0, nsm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ess_rounds, MO_CTA, sizeof(RoundsSmem)) != cudaSuccess ||
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return -1;
    grid_dev[dev] = per_sm * nsm > 0 ? per_sm * nsm : -1;
  }
  return grid_dev[dev];
}

// ---- small kernels used only by the orchestration ----

struct FinalizeArgs {
  const double* mean;     // [PB]
  const double* sd;       // [PB]
  const double* crhat;    // [2][PB] or null
  const double* cess;     // [2][PB]
  const long long* clag;  // [2][PB]
  double* mean_out;       // [PB] or null (already offset to this batch)
  double* std_out;
  double* rhat;           // or null
  double* rhat_bulk;
  double* rhat_tail;
  double* ess_bulk;
  double* ess_tail;
  int64_t* lag_bulk;
  int64_t* lag_tail;
};

__global__ void k_finalize(FinalizeArgs a, uint32_t M, int PB) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= PB) return;
  if (a.mean_out) a.mean_out[p] = a.mean[p];
  if (a.std_out) a.std_out[p] = a.sd[p];
  if (a.crhat) {
    const double rb = a.crhat[p], rt = a.crhat[(size_t)PB + p];
    // Python max(rb, rt): keeps rb unless rt > rb (diagnostics.py:40)
    if (a.rhat) a.rhat[p] = (rt > rb) ? rt : rb;
    if (a.rhat_bulk) a.rhat_bulk[p] = rb;
    if (a.rhat_tail) a.rhat_tail[p] = rt;
    if (a.ess_bulk) a.ess_bulk[p] = a.cess[p];
    if (a.ess_tail) a.ess_tail[p] = a.cess[(size_t)PB + p];
    if (a.lag_bulk) a.lag_bulk[p] = a.clag[p];
    if (a.lag_tail) a.lag_tail[p] = a.clag[(size_t)PB + p];
  }
}

// compare_stats inner loop (compare.py:39-43)
__global__ void k_compare(const double* __restrict__ ref, const double* __restrict__ act, int64_t n, double tol,
                          double* __restrict__ rel, uint8_t* __restrict__ passed, int* __restrict__ nfail) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = ref[i], a = act[i];
  // Python max(abs(r), 1e-12): keeps abs(r) unless 1e-12 > abs(r)  (NaN stays NaN)
  const double ar = fabs(r);
  const double denom = (1e-12 > ar) ? 1e-12 : ar;
  const double e = __ddiv_rn(fabs(__dsub_rn(a, r)), denom);
  const bool ok = e <= tol;
  if (rel) rel[i] = e;
  if (passed) passed[i] = ok ? 1 : 0;
  if (!ok && nfail) atomicAdd(nfail, 1);
}

// _checks (convert.py:165-166): counts of violations
__global__ void k_checks(const double* __restrict__ rhat, const double* __restrict__ essb, int64_t P, int* viol) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  if (!(essb[i] > 400.0)) atomicAdd(&viol[0], 1);
  if (!(rhat[i] < 1.01)) atomicAdd(&viol[1], 1);
}

// Summary rows [P][16] of the fused pass (include/b200ref.h B200REF_S_*).
__global__ void k_summary(const double* rhat, const double* essb, const double* esst, const double* mean,
                          const double* sd, const double* q, const double* rmean, const double* rsd,
                          const double* rq, double tol, double z_crit, int64_t Mref, int64_t P, double* out,
                          int* nfail) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double* o = out + p * B200REF_SUMMARY_K;
  o[B200REF_S_RHAT] = rhat[p];
  o[B200REF_S_ESS_BULK] = essb[p];
  o[B200REF_S_ESS_TAIL] = esst[p];
  o[B200REF_S_MEAN] = mean[p];
  o[B200REF_S_STD] = sd[p];
  o[B200REF_S_Q5] = q[p * 3 + 0];
  o[B200REF_S_Q50] = q[p * 3 + 1];
  o[B200REF_S_Q95] = q[p * 3 + 2];
  o[B200REF_S_REF_MEAN] = rmean[p];
  o[B200REF_S_REF_STD] = rsd[p];
  o[B200REF_S_REF_Q5] = rq[p * 3 + 0];
  o[B200REF_S_REF_Q50] = rq[p * 3 + 1];
  o[B200REF_S_REF_Q95] = rq[p * 3 + 2];
  int flags = 0;
  {
    const double ar = fabs(rmean[p]);
    const double e = __ddiv_rn(fabs(__dsub_rn(mean[p], rmean[p])), (1e-12 > ar) ? 1e-12 : ar);
    o[B200REF_S_REL_MEAN] = e;
    if (e <= tol) flags |= 1;
  }
  {
    const double ar = fabs(rsd[p]);
    const double e = __ddiv_rn(fabs(__dsub_rn(sd[p], rsd[p])), (1e-12 > ar) ? 1e-12 : ar);
    o[B200REF_S_REL_STD] = e;
    if (e <= tol) flags |= 2;
  }
  o[B200REF_S_FLAGS] = (double)flags;
  if (flags != 3) atomicAdd(nfail, 1);
  {  // extension: Monte-Carlo standard errors and the z-score of the mean difference
    const double se_c = __ddiv_rn(sd[p], __dsqrt_rn(essb[p]));
    const double se_r = __ddiv_rn(rsd[p], __dsqrt_rn((double)Mref));
    const double z = __ddiv_rn(__dsub_rn(mean[p], rmean[p]),
                               __dsqrt_rn(__dadd_rn(__dmul_rn(se_c, se_c), __dmul_rn(se_r, se_r))));
    o[B200REF_S_MCSE_MEAN] = se_c;
    o[B200REF_S_REF_SE_MEAN] = se_r;
    o[B200REF_S_Z_MEAN] = z;
    o[B200REF_S_Z_PASS] = (fabs(z) <= z_crit) ? 1.0 : 0.0;
  }
}

// Opt-in dynamic shared-memory sizes: function attributes are per device; several host threads may arrive at once.
int ensure_kernel_attributes() {
  static std::once_flag attr_once[MAX_DEVICES];
  static cudaError_t attr_err[MAX_DEVICES];
  const int dev = current_device();
  std::call_once(attr_once[dev], [dev] {
    cudaError_t e = cudaSuccess;
    auto set = [&e](const void* f, size_t bytes) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    };
    set((const void*)k_partition, sizeof(PartSmem));
    set((const void*)k_part1, sizeof(Part1Smem));
    set((const void*)k_finish, sizeof(FinSmemSmall));
    set((const void*)k_finish_big, sizeof(FinSmemBig));
    set((const void*)k_moments0, sizeof(Mo0Smem));
    set((const void*)k_moments0t, M2_SMEM);
    set((const void*)k_ess_rounds, sizeof(RoundsSmem));
    set((const void*)k_decide0, 3 * (size_t)MAX_CHAINS * sizeof(double));  // 96 KB at the advertised 4096 chains
    attr_err[dev] = e;
  });
  CUDA_TRY(attr_err[dev]);
  return B200REF_OK;
}

// Chain statistics of the scores zz[PB][C][N] = {z_bulk, z_folded}: split R-hat, ESS (round 0 for everybody, then the
// persistent rounds kernel for sticky chains).  `cs.rhat / ess / lag` must be set by the caller; the rest is carved here.
int launch_chain_stats(const double2* zz, int PB, int64_t C, int64_t N, char* wb, const Layout& L, bool geyer,
                   cudaStream_t st, ChainState& cs) {
  // ---- streaming chain statistics: round 0 for everybody, further rounds for sticky chains ----
  const ChainGeom g = chain_geom(C, N);
  cs.part = (double*)(wb + L.cpart);
  cs.partr = (double*)(wb + L.cpartr);
  cs.sticky = (uint8_t*)(wb + L.csticky);
  cs.cmean = (double*)(wb + L.cmean);
  cs.csum = (double*)(wb + L.csum);
  cs.head = (double*)(wb + L.chead);
  cs.tail = (double*)(wb + L.ctail);
  cs.varhat = (double*)(wb + L.cvarhat);
  cs.rhosum = (double*)(wb + L.crhosum);
  cs.gpend = (double*)(wb + L.cgpend);
  cs.gprev = (double*)(wb + L.cgprev);
  cs.geyer = geyer ? 1 : 0;
  cs.wl[0] = (int*)(wb + L.cwl0);
  cs.wl[1] = (int*)(wb + L.cwl1);
  cs.nwl = (int*)(wb + L.cnwl);
  CUDA_TRY(cudaMemsetAsync(cs.nwl, 0, 2 * sizeof(int), st));
  if (g.nunits) {
    // units per CTA: about MO_CTA_TILES sub-tiles of work behind one exposed first copy
    const uint32_t upc = std::max<uint32_t>(1u, MO_CTA_TILES / g.tpu);
    const unsigned gx = (unsigned)((C * g.nunits + upc - 1) / upc);
    if (B200REF_MOMENTS_TILE != 0 && (N % 8) == 0 && ((uintptr_t)zz % 128) == 0) {
      // tile variant: 2-D TMA into swizzled landing buffers, both kinds per thread (no de-interleave pass); for the
      // parameters whose first chain looks sticky it also covers lags 17 .. 48
      if (B200REF_STICKY_ROUND0) {
        LAUNCH(K_CHAIN, st, k_predict_sticky<<<(PB + 3) / 4, 128, 0, st>>>(zz, g, PB, cs.sticky));
      } else {
        CUDA_TRY(cudaMemsetAsync(cs.sticky, 0, (size_t)PB, st));
      }
      CUtensorMap map;
      const int rc = make_map_pairs(&map, zz, (uint64_t)PB * (uint64_t)C * (uint64_t)(N / 8));
      if (rc) return rc;
      // every warp of a CTA owns whole units: about 32 tiles (of 32 rows) of work behind one exposed first copy
      const uint32_t upw = std::max<uint32_t>(1u, 32u / (g.tpu * (MO_SUB / 8 / M2_ROWS)));
      const uint32_t upc_t = M2_NW * upw;
      const unsigned gxt = (unsigned)((C * g.nunits + upc_t - 1) / upc_t);
      LAUNCH(K_CHAIN, st, k_moments0t<<<dim3(gxt, PB), M2_NT, M2_SMEM, st>>>(map, g, upc_t, cs.part, cs.sticky));
    } else {
      CUDA_TRY(cudaMemsetAsync(cs.sticky, 0, (size_t)PB, st));
      LAUNCH(K_CHAIN, st, k_moments0<<<dim3(gx, PB), MO_CTA, sizeof(Mo0Smem), st>>>(zz, g, upc, cs.part));
    }
  }
  LAUNCH(K_DECIDE, st, k_decide0<<<2 * PB, DEC_NT, 3 * (size_t)C * sizeof(double), st>>>(zz, g, PB, cs));
  if ((uint32_t)N > 1u + LAGB) {
    // every later round (32 lags each) in ONE cooperative launch: the grid loops on the device until the
    // work list is empty -- no list length travels to the host, nothing synchronises between rounds
    unsigned int* gbar = (unsigned int*)(wb + L.cgbar);
    CUDA_TRY(cudaMemsetAsync(gbar, 0, sizeof(unsigned int), st));
    const double2* zzc = zz;
    ChainGeom gg = g;
    int pbv = PB;
    void* args[] = {(void*)&zzc, (void*)&gg, (void*)&pbv, (void*)&cs, (void*)&gbar};
    const int grid = rounds_grid();
    if (grid < 1) return fail(B200REF_ERR_CUDA, "k_ess_rounds cannot be made resident on this device");
    LAUNCH(K_CHAIN, st, CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_ess_rounds, dim3((unsigned)grid), dim3(MO_CTA), args,
                                                            sizeof(RoundsSmem), st)));
  }
  return B200REF_OK;
}

// zz[i] = {zb[i], zf[i]} (zf null: zb twice) -- staging of caller-supplied scores for the chain kernels
__global__ void k_interleave(const double* __restrict__ zb, const double* __restrict__ zf, int64_t n, double2* __restrict__ zz) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double a = zb[i];
    zz[i] = make_double2(a, zf ? zf[i] : a);
  }
}

struct Mode {
  bool diag;         // run fold + chain kernels
  bool want_ranks;   // scatter average ranks (ranks API)
  bool geyer = false;  // extension: Geyer initial-monotone ESS truncation instead of the reference's first-negative rule
};

// Process P parameters in batches.  `x` [P][M]; outputs are offset per batch.
int run(const double* x, int64_t P, int64_t C, int64_t N, const double* q_host, int nq, const b200ref_outputs* out,
        double* rank_b, double* z_b, double* rank_f, double* z_f, Mode mode, void* ws, size_t ws_bytes,
        cudaStream_t st, int slot = 0, int** deferred_flag = nullptr) {
  const int64_t M = C * N;
  if (!x || P < 0 || C < 1 || N < 1) return fail(B200REF_ERR_ARG, "bad shape P=%lld C=%lld N=%lld", (long long)P,
                                                 (long long)C, (long long)N);
  if (M >= (1ll << 31)) return fail(B200REF_ERR_ARG, "C*N = %lld exceeds 2^31-1", (long long)M);
  if (nq < 0 || nq > B200REF_MAX_QUANTILES) return fail(B200REF_ERR_ARG, "nq=%d out of range", nq);
  if (P == 0) return B200REF_OK;
  if (!ws) return fail(B200REF_ERR_WORKSPACE, "null workspace");
  if (mode.diag && C > MAX_CHAINS) return fail(B200REF_ERR_ARG, "more than %d chains are not supported (C=%lld)", MAX_CHAINS, (long long)C);
  // numpy.quantile raises for levels outside [0, 1] (backends_numpy.py:44); NaN fails both comparisons
  for (int i = 0; i < nq; ++i)
    if (!(q_host[i] >= 0.0 && q_host[i] <= 1.0))
      return fail(B200REF_ERR_ARG, "quantile level %g is outside [0, 1]", q_host[i]);
  const bool need_z = mode.diag;
  const bool ws_zz = mode.diag;  // interleaved {z_bulk, z_folded} scratch for k_chain
  const int64_t PBmax = fit_batch(P, M, ws_zz, ws_bytes, BATCH_CAP, C, N);
  if (PBmax < 1) return fail(B200REF_ERR_WORKSPACE, "workspace of %zu bytes cannot hold one parameter", ws_bytes);
  Layout L;
  L.compute(PBmax, M, ws_zz, C, N);
  char* w = (char*)ws;
  double* ztab = (double*)(w + L.ztab);
  double* zodd = (double*)(w + L.zodd);
  int* flag = (int*)(w + L.flag);
  char* wb = w + L.header;
  {
    const int rc = ensure_kernel_attributes();
    if (rc) return rc;
  }

  CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
  const bool want_scores = need_z || z_b || z_f || rank_f;
  if (want_scores) LAUNCH(K_ZTABLE, st, k_ztable<<<296, 256, 0, st>>>(ztab, zodd, (uint32_t)M));
  QuantArgs qa;
  qa.nq = nq;
  for (int i = 0; i < nq; ++i) qa.q[i] = q_host[i];
  const int nb1 = choose_nb1(M), nsub = choose_sub(M);
  const bool wide = nsub > 1;
  const bool slabs = use_slabs(M);

  for (int64_t p0 = 0; p0 < P; p0 += PBmax) {
    const int PB = (int)std::min<int64_t>(PBmax, P - p0);
    const double* xb = x + (size_t)p0 * M;
    BucketBufs bb;
    bb.fn = (BucketFn*)(wb + L.fn);
    bb.cnt = (uint32_t*)(wb + L.cnt);
    bb.boff = (uint32_t*)(wb + L.boff);
    bb.bstride = wide ? NBW_MAX + 1 : NB1_MAX + 1;
    bb.cur = (uint32_t*)(wb + L.cur);
    bb.spart = (double*)(wb + L.spart);
    bb.fallback = (uint8_t*)(wb + L.fallback);
    bb.nfb = (int*)(wb + L.nfb);
    bb.nbigb = (int*)(wb + L.nbigb);
    bb.biglist = (int*)(wb + L.biglist);
    bb.biglist_cap = PB * BIGLIST_PER_PARAM;
    bb.plist = (int*)(wb + L.plist);
    bb.flag = flag;
    bb.pkey = (uint64_t*)(wb + L.pkey);
    bb.pidx = (uint32_t*)(wb + L.pidx);
    bb.pstride = slab_records(M);
    bb.slab = slabs ? 1 : 0;
    bb.slab_cap = (uint32_t)std::min<size_t>(slab_records(M), 0x7fffffffu);
    bb.skey = (uint64_t*)(wb + L.skey);
    bb.sidx = (uint32_t*)(wb + L.sidx);
    bb.kcode = (uint32_t*)(wb + L.kcode);
    bb.bmeta = bb.kcode;  // dead before k_finish / k_rank_codes write the tie codes
    bb.sel = (int*)(wb + L.sel);
    bb.mean = (double*)(wb + L.mean);
    bb.sd = (double*)(wb + L.sd);
    double* rank_b_batch = rank_b ? rank_b + (size_t)p0 * M : nullptr;

    // ---- distribution sort: sample -> histogram -> offsets -> partition -> on-chip finish ----
    const dim3 gc(L.nchunks, PB);
    LAUNCH(K_SAMPLE, st, k_sample<<<PB, 512, 0, st>>>(xb, (uint32_t)M, bb));
    if (slabs) {
      // single pass: the partition itself counts (one global atomic per (chunk, bucket) reserves the run)
      CUDA_TRY(cudaMemsetAsync(bb.cur, 0, (size_t)PB * NB1_MAX * 4, st));
      LAUNCH(K_PARTITION, st, k_part1<<<gc, P1_NT, sizeof(Part1Smem), st>>>(xb, (uint32_t)M, L.nchunks, bb));
      LAUNCH(K_BSCAN, st, k_bscan1<<<PB, NB1_MAX, 0, st>>>((uint32_t)M, L.nchunks, bb));
    } else if (!wide) {
      LAUNCH(K_BHIST, st, k_bhist<<<gc, PART_NT, 0, st>>>(xb, (uint32_t)M, L.nchunks, bb));
      LAUNCH(K_BSCAN, st, k_bscan<<<PB, NB1_MAX, 0, st>>>((uint32_t)M, L.nchunks, bb));
    } else {
      CUDA_TRY(cudaMemsetAsync(bb.cur, 0, (size_t)PB * NBW_MAX * 4, st));
      LAUNCH(K_BHIST, st, k_bhist_w<<<gc, PART_NT, sizeof(WideSmem), st>>>(xb, (uint32_t)M, L.nchunks, bb));
      LAUNCH(K_BSCAN, st, k_bscan_w<<<PB, BSW_NT, 0, st>>>((uint32_t)M, L.nchunks, bb));
    }

    // ---- distribution sort, second half (main stream) ----
    // ---- LSD fallback for parameters the bucket function could not balance: a fixed small grid
    //      strides over the (usually empty) list, on a side stream, overlapped with k_partition /
    //      k_finish which skip those parameters. ----
    SortBufs sb;
    sb.key[0] = bb.pkey; sb.key[1] = bb.skey;
    sb.idx[0] = bb.pidx; sb.idx[1] = bb.sidx;
    sb.hist = (uint32_t*)(wb + L.hist);
    sb.sel = bb.sel;
    sb.sum = (double*)(wb + L.sum);
    sb.flag = flag;
    sb.nlist = bb.nfb;
    sb.plist = bb.plist;
    RankBufs rb;
    rb.key[0] = sb.key[0]; rb.key[1] = sb.key[1];
    rb.idx[0] = sb.idx[0]; rb.idx[1] = sb.idx[1];
    rb.sel = sb.sel;
    rb.kcode = bb.kcode;
    rb.fallback = bb.fallback;
    rb.nfb = bb.nfb;
    rb.plist = bb.plist;
    rb.med = (double*)(wb + L.med);
    rb.m0 = (uint32_t*)(wb + L.m0);
    rb.apos = (uint32_t*)(wb + L.apos);
    rb.ztab = want_scores ? ztab : nullptr;
    rb.zodd = zodd;
    rb.flag = flag;
    const dim3 gt(L.ntiles, PB);
    auto launch_lsd = [&]() -> int {
      // while per-kernel timing is on, the fallback runs on the main stream so that every
      // event pair brackets one kernel alone (overlapped launches would share their interval)
      SideStream& ss = side_stream(slot);
      cudaStream_t fs = g_prof ? st : ss.s;
      CUDA_TRY(cudaEventRecord(ss.fork, st));
      CUDA_TRY(cudaStreamWaitEvent(fs, ss.fork, 0));
      const int fbg = std::min(PB, FALLBACK_GRID);
      LAUNCH(K_HIST0, fs, k_hist0<<<fbg, 1024, 0, fs>>>(xb, (uint32_t)M, sb));
      LAUNCH(K_RADIX, fs, k_radix_pass<true><<<fbg, SORT_NT, 0, fs>>>(xb, (uint32_t)M, 0, sb));
      for (int pass = 1; pass < 8; ++pass)
        LAUNCH(K_RADIX, fs, k_radix_pass<false><<<fbg, SORT_NT, 0, fs>>>(xb, (uint32_t)M, pass, sb));
      LAUNCH(K_RANKC, fs, k_rank_codes<<<dim3(L.ntiles, std::min(PB, 8)), RANK_NT, 0, fs>>>(rb, (uint32_t)M, rank_b_batch));
      CUDA_TRY(cudaEventRecord(ss.join, fs));
      return B200REF_OK;
    };
    if (!slabs) {
      // (two-pass partition: the list is known after k_bscan, the fallback overlaps k_partition / k_finish)
      const int rc = launch_lsd();
      if (rc) return rc;
      if (!wide)
        LAUNCH(K_PARTITION, st, k_partition<<<gc, PART_NT, sizeof(PartSmem), st>>>(xb, (uint32_t)M, L.nchunks, bb));
      else
        LAUNCH(K_PARTITION, st, k_partition_w<<<gc, PART_NT, sizeof(WideSmem), st>>>(xb, (uint32_t)M, L.nchunks, bb));
    }
    FinishOut fin;
    memset(&fin, 0, sizeof(fin));
    fin.rank = rank_b_batch;
    if (!want_scores && !rank_b) {
      // statistics only: the order statistics k_param_prep reads (quantile_lerp's lo / hi, the median's one or two)
      auto want = [&](int64_t t) { fin.pos[fin.npos++] = (uint32_t)std::min<int64_t>(std::max<int64_t>(t, 0), M - 1); };
      want((M - 1) / 2);
      want(M / 2);
      for (int i = 0; i < nq; ++i) {
        const double vi = (double)(M - 1) * q_host[i];
        const int64_t lo = (int64_t)std::floor(vi);
        want(lo);
        want(lo + 1);
      }
    }
    LAUNCH(K_FINISH, st, k_finish<<<dim3((nb1 * nsub + FIN_BPC - 1) / FIN_BPC, PB), FIN_NT, sizeof(FinSmemSmall), st>>>((uint32_t)M, nsub, nb1 * nsub, bb, fin));
    fin.npos = 0;  // oversize buckets are always ranked
    LAUNCH(K_FINISHB, st, k_finish_big<<<std::min(PB * 4, FALLBACK_GRID), FINB_NT, sizeof(FinSmemBig), st>>>((uint32_t)M, bb, fin));
    if (slabs) {
      // the LSD path works in the dense [p][M] views of pkey / pidx, which overlap other parameters' slabs: it
      // starts when the finish kernels have read them
      const int rc = launch_lsd();
      if (rc) return rc;
    }
    CUDA_TRY(cudaStreamWaitEvent(st, side_stream(slot).join, 0));

    PrepOut po;
    po.quant = (out && out->quantiles) ? out->quantiles + (size_t)p0 * nq : nullptr;
    po.ostats = (out && out->order_stats) ? out->order_stats + (size_t)p0 * nq * 2 : nullptr;
    po.median = (out && out->median) ? out->median + p0 : nullptr;
    LAUNCH(K_PREP, st, k_param_prep<<<(PB + 127) / 128, 128, 0, st>>>(rb, (uint32_t)M, PB, qa, po));

    double2* zz = ws_zz ? (double2*)(wb + L.zz) : nullptr;
    if (want_scores) {
      LAUNCH(K_MPART, st, k_merge_partition<<<dim3((L.ntiles + 1 + 127) / 128, PB), 128, 0, st>>>(rb, (uint32_t)M, L.ntiles, PB));
      FoldOut fo;
      memset(&fo, 0, sizeof(fo));
      fo.zz = zz;
      fo.z = z_b ? z_b + (size_t)p0 * M : nullptr;
      fo.zf = z_f ? z_f + (size_t)p0 * M : nullptr;
      fo.rankf = rank_f ? rank_f + (size_t)p0 * M : nullptr;
      fo.bin_shift = (zz && !fo.z && !fo.zf && !fo.rankf) ? fold_bin_shift(M) : 0;
      if (fo.bin_shift) {
        // record segments live in buffers that are dead by now (the partitioned keys / positions) + one more plane
        fo.cursor = (uint32_t*)(wb + L.bcur);
        fo.rpos = bb.pidx;
        fo.rzb = (double*)bb.pkey;
        fo.rzf = (double*)(wb + L.rzf);
        CUDA_TRY(cudaMemsetAsync(fo.cursor, 0, (size_t)PB * 32 * 4, st));
      }
      LAUNCH(K_FOLD, st, k_fold<<<gt, RANK_NT, 0, st>>>(rb, (uint32_t)M, L.ntiles, fo));
      if (fo.bin_shift) {
        const unsigned gx = (unsigned)((M + UNBIN_NT * UNBIN_IPT - 1) / (UNBIN_NT * UNBIN_IPT));
        LAUNCH(K_FOLD, st, k_unbin<<<dim3(gx, PB), UNBIN_NT, 0, st>>>(fo, bb.fallback, flag, (uint32_t)M));
      }
    }
    ChainState cs;
    memset(&cs, 0, sizeof(cs));
    cs.rhat = (double*)(wb + L.crhat);
    cs.ess = (double*)(wb + L.cess);
    cs.lag = (int64_t*)(wb + L.clag);
    if (mode.diag) {
      const int rc = launch_chain_stats(zz, PB, C, N, wb, L, mode.geyer, st, cs);
      if (rc) return rc;
    }
    if (out) {
      FinalizeArgs fa;
      memset(&fa, 0, sizeof(fa));
      fa.mean = bb.mean;
      fa.sd = bb.sd;
      fa.mean_out = out->mean ? out->mean + p0 : nullptr;
      fa.std_out = out->std ? out->std + p0 : nullptr;
      if (mode.diag) {
        fa.crhat = cs.rhat;
        fa.cess = cs.ess;
        fa.clag = (const long long*)cs.lag;
        fa.rhat = out->rhat ? out->rhat + p0 : nullptr;
        fa.rhat_bulk = out->rhat_bulk ? out->rhat_bulk + p0 : nullptr;
        fa.rhat_tail = out->rhat_tail ? out->rhat_tail + p0 : nullptr;
        fa.ess_bulk = out->ess_bulk ? out->ess_bulk + p0 : nullptr;
        fa.ess_tail = out->ess_tail ? out->ess_tail + p0 : nullptr;
        fa.lag_bulk = out->lag_bulk ? out->lag_bulk + p0 : nullptr;
        fa.lag_tail = out->lag_tail ? out->lag_tail + p0 : nullptr;
      }
      LAUNCH(K_FINALIZE, st, k_finalize<<<(PB + 127) / 128, 128, 0, st>>>(fa, (uint32_t)M, PB));
    }
  }
  CUDA_TRY(cudaGetLastError());
  if (deferred_flag) {  // the caller joins the stream and reads the flag itself
    *deferred_flag = flag;
    return B200REF_OK;
  }
  int hflag = 0;
  CUDA_TRY(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (hflag & 1) return fail(B200REF_ERR_NONFINITE, "non-finite value (NaN or inf) in input draws");
  return B200REF_OK;
}

}  // namespace

extern "C" {

int b200ref_version(void) { return B200REF_VERSION; }
const char* b200ref_last_error(void) { return g_err; }
uint64_t b200ref_kernel_launches(void) { return g_launches.load(); }

size_t b200ref_workspace_bytes(int64_t P, int64_t C, int64_t N) {
  if (P < 1) P = 1;
  Layout l;
  l.compute(std::min<int64_t>(P, BATCH_CAP), C * N, true, C, N);
  return l.header + l.per_batch;
}
size_t b200ref_workspace_min_bytes(int64_t C, int64_t N) {
  Layout l;
  l.compute(1, C * N, true, C, N);
  return l.header + l.per_batch;
}

int b200ref_profile(int enable) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof = enable != 0;
  return B200REF_OK;
}

int b200ref_profile_read(double* ms_out, uint64_t* launches_out, int n, int reset) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      g_prof_ms[r.id] += ms;
      g_prof_n[r.id] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof_recs.clear();
  for (int i = 0; i < n && i < K_COUNT; ++i) {
    if (ms_out) ms_out[i] = g_prof_ms[i];
    if (launches_out) launches_out[i] = g_prof_n[i];
  }
  if (reset) {
    for (int i = 0; i < K_COUNT; ++i) { g_prof_ms[i] = 0.0; g_prof_n[i] = 0; }
  }
  return K_COUNT;
}

const char* b200ref_profile_name(int id) { return (id >= 0 && id < K_COUNT) ? kKernelNames[id] : ""; }

int b200ref_ranks_f64(const double* x, int64_t P, int64_t C, int64_t N, double* avg_rank, double* z,
                      double* avg_rank_folded, double* z_folded, void* ws, size_t ws_bytes, void* stream) {
  g_err[0] = 0;
  Mode m{false, true};
  return run(x, P, C, N, nullptr, 0, nullptr, avg_rank, z, avg_rank_folded, z_folded, m, ws, ws_bytes,
             (cudaStream_t)stream);
}

int b200ref_diagnostics_f64(const double* x, int64_t P, int64_t C, int64_t N, const double* q_host, int nq,
                            const b200ref_outputs* out, void* ws, size_t ws_bytes, void* stream) {
  g_err[0] = 0;
  if (!out) return fail(B200REF_ERR_ARG, "null outputs");
  if (nq > 0 && !q_host) return fail(B200REF_ERR_ARG, "null quantile levels");
  Mode m{true, false};
  return run(x, P, C, N, q_host, nq, out, nullptr, nullptr, nullptr, nullptr, m, ws, ws_bytes,
             (cudaStream_t)stream);
}

int b200ref_diagnostics_ex_f64(const double* x, int64_t P, int64_t C, int64_t N, const double* q_host, int nq,
                               int ess_mode, const b200ref_outputs* out, void* ws, size_t ws_bytes, void* stream) {
  g_err[0] = 0;
  if (!out) return fail(B200REF_ERR_ARG, "null outputs");
  if (nq > 0 && !q_host) return fail(B200REF_ERR_ARG, "null quantile levels");
  if (ess_mode != B200REF_ESS_FIRST_NEGATIVE && ess_mode != B200REF_ESS_GEYER)
    return fail(B200REF_ERR_ARG, "unknown ess_mode %d", ess_mode);
  Mode m{true, false, ess_mode == B200REF_ESS_GEYER};
  return run(x, P, C, N, q_host, nq, out, nullptr, nullptr, nullptr, nullptr, m, ws, ws_bytes,
             (cudaStream_t)stream);
}

int b200ref_chain_stats_f64(const double* zb, const double* zf, int64_t P, int64_t C, int64_t N, int ess_mode,
                            const b200ref_outputs* out, void* ws, size_t ws_bytes, void* stream) {
  g_err[0] = 0;
  if (!out) return fail(B200REF_ERR_ARG, "null outputs");
  if (!zb || P < 0 || C < 1 || N < 1) return fail(B200REF_ERR_ARG, "bad shape P=%lld C=%lld N=%lld", (long long)P, (long long)C, (long long)N);
  if (C > MAX_CHAINS) return fail(B200REF_ERR_ARG, "more than %d chains are not supported (C=%lld)", MAX_CHAINS, (long long)C);
  if (C * N >= (1ll << 31)) return fail(B200REF_ERR_ARG, "C*N = %lld exceeds 2^31-1", (long long)(C * N));
  if (ess_mode != B200REF_ESS_FIRST_NEGATIVE && ess_mode != B200REF_ESS_GEYER)
    return fail(B200REF_ERR_ARG, "unknown ess_mode %d", ess_mode);
  if (P == 0) return B200REF_OK;
  if (!ws) return fail(B200REF_ERR_WORKSPACE, "null workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = C * N;
  const int64_t PBmax = fit_batch(P, M, true, ws_bytes, BATCH_CAP, C, N);
  if (PBmax < 1) return fail(B200REF_ERR_WORKSPACE, "workspace of %zu bytes cannot hold one parameter", ws_bytes);
  Layout L;
  L.compute(PBmax, M, true, C, N);
  char* wb = (char*)ws + L.header;
  {
    int rc = ensure_kernel_attributes();
    if (rc) return rc;
  }
  for (int64_t p0 = 0; p0 < P; p0 += PBmax) {
    const int PB = (int)std::min<int64_t>(PBmax, P - p0);
    double2* zz = (double2*)(wb + L.zz);
    const int64_t n = (int64_t)PB * M;
    LAUNCH(K_INGEST, st, k_interleave<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(
                             zb + (size_t)p0 * M, zf ? zf + (size_t)p0 * M : nullptr, n, zz));
    ChainState cs;
    memset(&cs, 0, sizeof(cs));
    cs.rhat = (double*)(wb + L.crhat);
    cs.ess = (double*)(wb + L.cess);
    cs.lag = (int64_t*)(wb + L.clag);
    int rc = launch_chain_stats(zz, PB, C, N, wb, L, ess_mode == B200REF_ESS_GEYER, st, cs);
    if (rc) return rc;
    FinalizeArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.crhat = cs.rhat;
    fa.cess = cs.ess;
    fa.clag = (const long long*)cs.lag;
    fa.rhat = out->rhat ? out->rhat + p0 : nullptr;
    fa.rhat_bulk = out->rhat_bulk ? out->rhat_bulk + p0 : nullptr;
    fa.rhat_tail = out->rhat_tail ? out->rhat_tail + p0 : nullptr;
    fa.ess_bulk = out->ess_bulk ? out->ess_bulk + p0 : nullptr;
    fa.ess_tail = out->ess_tail ? out->ess_tail + p0 : nullptr;
    fa.lag_bulk = out->lag_bulk ? out->lag_bulk + p0 : nullptr;
    fa.lag_tail = out->lag_tail ? out->lag_tail + p0 : nullptr;
    LAUNCH(K_FINALIZE, st, k_finalize<<<(PB + 127) / 128, 128, 0, st>>>(fa, (uint32_t)M, PB));
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(st));
  return B200REF_OK;
}

int b200ref_transpose_rows_f64(const double* in, int64_t M, int64_t P, int64_t p0, int64_t PB, double* out, void* stream) {
  g_err[0] = 0;
  if (M < 0 || P < 0 || PB < 0 || p0 < 0 || p0 + PB > P || ((M > 0 && PB > 0) && (!in || !out)))
    return fail(B200REF_ERR_ARG, "bad transpose arguments");
  if (M == 0 || PB == 0) return B200REF_OK;
  if ((PB + TR_TILE - 1) / TR_TILE > 65535) return fail(B200REF_ERR_ARG, "transpose: at most %d parameters per call", 65535 * TR_TILE);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)((M + TR_TILE - 1) / TR_TILE), (unsigned)((PB + TR_TILE - 1) / TR_TILE));
  LAUNCH(K_INGEST, st, k_transpose_rows<<<grid, TR_TILE * TR_ROWS, 0, st>>>(in, M, P, p0, PB, out));
  CUDA_TRY(cudaGetLastError());
  return B200REF_OK;
}

size_t b200ref_rows_workspace_bytes(int64_t P, int64_t C, int64_t N) {
  if (P < 1) P = 1;
  const int64_t pb = std::min<int64_t>(P, BATCH_CAP);
  return b200ref_workspace_bytes(pb, C, N) + up((size_t)pb * (size_t)(C * N) * 8) + ALIGN;
}

int b200ref_diagnostics_layout_f64(const double* x, int layout, int64_t P, int64_t C, int64_t N, const double* q_host,
                                   int nq, int ess_mode, const b200ref_outputs* out, void* ws, size_t ws_bytes,
                                   void* stream) {
  if (layout == B200REF_LAYOUT_PCN)
    return b200ref_diagnostics_ex_f64(x, P, C, N, q_host, nq, ess_mode, out, ws, ws_bytes, stream);
  g_err[0] = 0;
  if (layout != B200REF_LAYOUT_ROWS) return fail(B200REF_ERR_ARG, "unknown layout %d", layout);
  if (!out) return fail(B200REF_ERR_ARG, "null outputs");
  if (nq > 0 && !q_host) return fail(B200REF_ERR_ARG, "null quantile levels");
  if (ess_mode != B200REF_ESS_FIRST_NEGATIVE && ess_mode != B200REF_ESS_GEYER)
    return fail(B200REF_ERR_ARG, "unknown ess_mode %d", ess_mode);
  if (!x || P < 0 || C < 1 || N < 1) return fail(B200REF_ERR_ARG, "bad shape P=%lld C=%lld N=%lld", (long long)P, (long long)C, (long long)N);
  if (P == 0) return B200REF_OK;
  if (!ws) return fail(B200REF_ERR_WORKSPACE, "null workspace");
  const int64_t M = C * N;
  // parameter block: its transposed draws [pb][M] at the head of the workspace, the canonical workspace behind them
  int64_t pb = std::min<int64_t>(P, BATCH_CAP);
  const size_t min_ws = b200ref_workspace_min_bytes(C, N);
  while (pb > 1 && up((size_t)pb * M * 8) + ALIGN + std::min(b200ref_workspace_bytes(pb, C, N), std::max(min_ws, ws_bytes / 2)) > ws_bytes) pb = (pb + 1) / 2;
  const size_t head = up((size_t)pb * M * 8);
  if (head + ALIGN + min_ws > ws_bytes) return fail(B200REF_ERR_WORKSPACE, "workspace of %zu bytes cannot hold one transposed parameter block", ws_bytes);
  char* base = (char*)(((uintptr_t)ws + ALIGN - 1) / ALIGN * ALIGN);
  const size_t avail = ws_bytes - (size_t)(base - (char*)ws);
  double* xt = (double*)base;
  char* cws = base + head;
  const size_t cws_bytes = avail - head;
  Mode m{true, false, ess_mode == B200REF_ESS_GEYER};
  for (int64_t p0 = 0; p0 < P; p0 += pb) {
    const int64_t n = std::min<int64_t>(pb, P - p0);
    int rc = b200ref_transpose_rows_f64(x, M, P, p0, n, xt, stream);
    if (rc) return rc;
    b200ref_outputs o = *out;  // this block's slices of the caller's outputs
    auto adv = [&](auto*& ptr, int64_t stride) { if (ptr) ptr += p0 * stride; };
    adv(o.rhat, 1); adv(o.rhat_bulk, 1); adv(o.rhat_tail, 1); adv(o.ess_bulk, 1); adv(o.ess_tail, 1);
    adv(o.lag_bulk, 1); adv(o.lag_tail, 1); adv(o.mean, 1); adv(o.std, 1); adv(o.median, 1);
    adv(o.quantiles, nq); adv(o.order_stats, 2 * (int64_t)nq);
    rc = run(xt, n, C, N, q_host, nq, &o, nullptr, nullptr, nullptr, nullptr, m, cws, cws_bytes, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return B200REF_OK;
}

int b200ref_stats_f64(const double* x, int64_t P, int64_t M, const double* q_host, int nq,
                      const b200ref_outputs* out, void* ws, size_t ws_bytes, void* stream) {
  g_err[0] = 0;
  if (!out) return fail(B200REF_ERR_ARG, "null outputs");
  if (nq > 0 && !q_host) return fail(B200REF_ERR_ARG, "null quantile levels");
  Mode m{false, false};
  return run(x, P, 1, M, q_host, nq, out, nullptr, nullptr, nullptr, nullptr, m, ws, ws_bytes,
             (cudaStream_t)stream);
}

int b200ref_compare_f64(const double* ref, const double* act, int64_t n, double tol, double* rel_err,
                        uint8_t* passed, uint8_t* all_passed_dev, void* stream) {
  g_err[0] = 0;
  if (n < 0 || (n > 0 && (!ref || !act))) return fail(B200REF_ERR_ARG, "bad compare arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int* nfail = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&nfail, sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(nfail, 0, sizeof(int), st));
  if (n > 0) {
    LAUNCH(K_COMPARE, st, k_compare<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ref, act, n, tol, rel_err, passed, nfail));
  }
  int h = 0;
  CUDA_TRY(cudaMemcpyAsync(&h, nfail, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (all_passed_dev) {
    const uint8_t v = (h == 0) ? 1 : 0;
    CUDA_TRY(cudaMemcpyAsync(all_passed_dev, &v, 1, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  CUDA_TRY(cudaFreeAsync(nfail, st));
  CUDA_TRY(cudaGetLastError());
  return B200REF_OK;
}

int b200ref_checks_f64(const double* rhat, const double* essb, int64_t P, uint8_t* verdicts_host, void* stream) {
  g_err[0] = 0;
  if (P < 0 || !verdicts_host || (P > 0 && (!rhat || !essb))) return fail(B200REF_ERR_ARG, "bad checks arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int* viol = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&viol, 2 * sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(viol, 0, 2 * sizeof(int), st));
  if (P > 0) {
    LAUNCH(K_CHECKS, st, k_checks<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(rhat, essb, P, viol));
  }
  int h[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(h, viol, sizeof(h), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaFreeAsync(viol, st));
  verdicts_host[0] = h[0] == 0;
  verdicts_host[1] = h[1] == 0;
  return B200REF_OK;
}

// Scratch for the fused pass, carved from the END of the workspace:
//   per-parameter result columns (13 doubles) + nfail.
// workspace of the (non-diagnostic) statistics pass over the reference draws of P parameters
static size_t ref_workspace_bytes(int64_t P, int64_t Mref) {
  Layout l;
  l.compute(std::min<int64_t>(std::max<int64_t>(P, 1), BATCH_CAP), Mref, false);
  return l.header + l.per_batch;
}
struct AuxStream {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static AuxStream& aux_stream() {
  thread_local AuxStream pool[MAX_DEVICES];
  AuxStream& ax = pool[current_device()];
  if (!ax.s) {
    cudaStreamCreateWithFlags(&ax.s, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ax.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ax.join, cudaEventDisableTiming);
  }
  return ax;
}
static size_t fused_tail_bytes(int64_t P) { return up((size_t)P * 13 * 8) + up(16); }

size_t b200ref_full_pass_workspace_bytes(int64_t P, int64_t C, int64_t N, int64_t Mref) {
  return b200ref_workspace_bytes(P, C, N) + ref_workspace_bytes(P, Mref) + fused_tail_bytes(std::max<int64_t>(P, 1)) + 4 * ALIGN;
}

int b200ref_full_pass_f64(const double* cand, int64_t P, int64_t C, int64_t N, const double* ref, int64_t Mref,
                          double tol, double z_crit, double* summary, uint8_t* passed_host, void* ws,
                          size_t ws_bytes, void* stream) {
  g_err[0] = 0;
  if (!cand || !ref || !summary || P < 1 || Mref < 1) return fail(B200REF_ERR_ARG, "bad full_pass arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tail = fused_tail_bytes(P);
  if (ws_bytes <= tail) return fail(B200REF_ERR_WORKSPACE, "workspace too small");
  const size_t body = (ws_bytes - tail) / ALIGN * ALIGN;
  double* cols = (double*)((char*)ws + body);
  int* nfail = (int*)((char*)cols + up((size_t)P * 13 * 8));
  double *rhat = cols, *essb = cols + P, *esst = cols + 2 * P, *mean = cols + 3 * P, *sd = cols + 4 * P,
         *q = cols + 5 * P, *rmean = cols + 8 * P, *rsd = cols + 9 * P, *rq = cols + 10 * P;
  const double qs[3] = {0.05, 0.5, 0.95};
  b200ref_outputs oc;
  memset(&oc, 0, sizeof(oc));
  oc.rhat = rhat; oc.ess_bulk = essb; oc.ess_tail = esst; oc.mean = mean; oc.std = sd; oc.quantiles = q;
  b200ref_outputs orf;
  memset(&orf, 0, sizeof(orf));
  orf.mean = rmean; orf.std = rsd; orf.quantiles = rq;
  // The reference-draw statistics are independent of the candidate pass: when the workspace has
  // room for both, they run on an auxiliary stream underneath it (small, latency-bound launches).
  const size_t ref_need = ref_workspace_bytes(P, Mref);
  const bool overlap = !g_prof && ref_need <= body / 4;
  int rc;
  if (overlap) {
    AuxStream& ax = aux_stream();
    char* ref_ws = (char*)ws + (body - ref_need) / ALIGN * ALIGN;
    const size_t cand_body = (size_t)(ref_ws - (char*)ws);
    int* ref_flag = nullptr;
    CUDA_TRY(cudaEventRecord(ax.fork, st));
    CUDA_TRY(cudaStreamWaitEvent(ax.s, ax.fork, 0));
    rc = run(ref, P, 1, Mref, qs, 3, &orf, nullptr, nullptr, nullptr, nullptr, Mode{false, false}, ref_ws, ref_need,
             ax.s, 1, &ref_flag);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ax.join, ax.s));
    rc = run(cand, P, C, N, qs, 3, &oc, nullptr, nullptr, nullptr, nullptr, Mode{true, false}, ws, cand_body, st);
    CUDA_TRY(cudaStreamWaitEvent(st, ax.join, 0));  // also on an error return path below: the aux work is always joined
    int hflag = 0;
    CUDA_TRY(cudaMemcpyAsync(&hflag, ref_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (rc) return rc;
    if (hflag & 1) return fail(B200REF_ERR_NONFINITE, "non-finite value (NaN or inf) in input draws");
  } else {
    rc = run(cand, P, C, N, qs, 3, &oc, nullptr, nullptr, nullptr, nullptr, Mode{true, false}, ws, body, st);
    if (rc) return rc;
    rc = run(ref, P, 1, Mref, qs, 3, &orf, nullptr, nullptr, nullptr, nullptr, Mode{false, false}, ws, body, st);
    if (rc) return rc;
  }
  CUDA_TRY(cudaMemsetAsync(nfail, 0, sizeof(int), st));
  LAUNCH(K_SUMMARY, st, k_summary<<<(unsigned)((P + 127) / 128), 128, 0, st>>>(rhat, essb, esst, mean, sd, q, rmean, rsd, rq, tol, z_crit,
                                                         Mref, P, summary, nfail));
  int h = 0;
  CUDA_TRY(cudaMemcpyAsync(&h, nfail, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  if (passed_host) *passed_host = (h == 0) ? 1 : 0;
  return B200REF_OK;
}

int b200ref_full_pass_host_f64(const double* cand_host, int64_t P, int64_t C, int64_t N, const double* ref_host,
                               int64_t Mref, double tol, double z_crit, double* summary_host, uint8_t* passed_host,
                               void* ws, size_t ws_bytes) {
  g_err[0] = 0;
  if (!cand_host || !ref_host || !summary_host || P < 1) return fail(B200REF_ERR_ARG, "bad full_pass_host arguments");
  // Stage inputs in parameter slabs: two slabs in flight, H2D of slab k+1 on the copy
  // stream overlaps the kernels of slab k on the compute stream.
  const int64_t M = C * N;
  const size_t per_param_in = ((size_t)M + (size_t)Mref + B200REF_SUMMARY_K) * 8;
  cudaStream_t s_copy, s_comp;
  CUDA_TRY(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
  // split: 2 input slabs + a compute workspace sized for one slab
  // slab size: ~1/16 of the input, between 32 MB and 256 MB of draws, so the PCIe copy of one
  // slab hides the kernels of the previous one and the exposed tail is one slab's compute
  const size_t total_in = per_param_in * (size_t)P;
  const size_t slab_bytes = std::min<size_t>(std::max<size_t>(total_in / 16, (size_t)32 << 20), (size_t)256 << 20);
  int64_t slab = std::max<int64_t>(1, std::min<int64_t>(P, (int64_t)(slab_bytes / per_param_in)));
  size_t need = 0;
  for (; slab >= 1; slab = (slab > 1) ? slab / 2 : 0) {
    need = 2 * up(per_param_in * slab) + b200ref_full_pass_workspace_bytes(slab, C, N, Mref) + ALIGN;
    if (need <= ws_bytes) break;
  }
  int rc = B200REF_OK;
  if (slab < 1) {
    rc = fail(B200REF_ERR_WORKSPACE, "workspace too small for host staging");
  } else {
    char* base = (char*)ws;
    char* in[2] = {base, base + up(per_param_in * slab)};
    char* cws = base + 2 * up(per_param_in * slab);
    const size_t cws_bytes = ws_bytes - 2 * up(per_param_in * slab);
    cudaEvent_t copied[2], computed[2];
    for (int i = 0; i < 2; ++i) {
      cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&computed[i], cudaEventDisableTiming);
    }
    uint8_t all_ok = 1;
    const int64_t nslabs = (P + slab - 1) / slab;
    auto issue_copy = [&](int64_t k) {
      const int b = (int)(k & 1);
      const int64_t p0 = k * slab, pb = std::min<int64_t>(slab, P - p0);
      double* dc = (double*)in[b];
      double* dr = dc + (size_t)slab * M;
      cudaStreamWaitEvent(s_copy, computed[b], 0);
      cudaMemcpyAsync(dc, cand_host + (size_t)p0 * M, (size_t)pb * M * 8, cudaMemcpyHostToDevice, s_copy);
      cudaMemcpyAsync(dr, ref_host + (size_t)p0 * Mref, (size_t)pb * Mref * 8, cudaMemcpyHostToDevice, s_copy);
      cudaEventRecord(copied[b], s_copy);
    };
    issue_copy(0);
    for (int64_t k = 0; k < nslabs && rc == B200REF_OK; ++k) {
      const int b = (int)(k & 1);
      const int64_t p0 = k * slab, pb = std::min<int64_t>(slab, P - p0);
      if (k + 1 < nslabs) issue_copy(k + 1);
      double* dc = (double*)in[b];
      double* dr = dc + (size_t)slab * M;
      double* dsum = dr + (size_t)slab * Mref;
      cudaStreamWaitEvent(s_comp, copied[b], 0);
      uint8_t ok = 0;
      rc = b200ref_full_pass_f64(dc, pb, C, N, dr, Mref, tol, z_crit, dsum, &ok, cws, cws_bytes, s_comp);
      if (rc) break;
      all_ok &= ok;
      cudaMemcpyAsync(summary_host + (size_t)p0 * B200REF_SUMMARY_K, dsum, (size_t)pb * B200REF_SUMMARY_K * 8,
                      cudaMemcpyDeviceToHost, s_comp);
      cudaEventRecord(computed[b], s_comp);
    }
    cudaStreamSynchronize(s_comp);
    cudaStreamSynchronize(s_copy);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(copied[i]);
      cudaEventDestroy(computed[i]);
    }
    if (rc == B200REF_OK) {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) rc = fail(B200REF_ERR_CUDA, "host pass failed: %s", cudaGetErrorString(e));
      if (passed_host) *passed_host = all_ok;
    }
  }
  cudaStreamDestroy(s_copy);
  cudaStreamDestroy(s_comp);
  return rc;
}

int b200ref_row_order_i64(const int64_t* chain, const int64_t* draw, int64_t rows, int* flags_host, void* stream) {
  g_err[0] = 0;
  if (rows < 0 || !flags_host || (rows > 0 && (!chain || !draw))) return fail(B200REF_ERR_ARG, "bad row_order arguments");
  cudaStream_t st = (cudaStream_t)stream;
  *flags_host = 0;
  if (rows < 2) return B200REF_OK;
  int* dflag = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&dflag, sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(dflag, 0, sizeof(int), st));
  const unsigned blocks = (unsigned)std::min<int64_t>((rows + 255) / 256, 148 * 8);
  LAUNCH(K_INGEST, st, k_row_order<<<blocks, 256, 0, st>>>(chain, draw, rows, dflag));
  CUDA_TRY(cudaMemcpyAsync(flags_host, dflag, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaFreeAsync(dflag, st));
  CUDA_TRY(cudaGetLastError());
  return B200REF_OK;
}

int b200ref_gather_rows_f64(const double* in, const int64_t* perm, int64_t P, int64_t M, double* out, void* stream) {
  g_err[0] = 0;
  if (P < 0 || M < 0 || ((P > 0 && M > 0) && (!in || !perm || !out))) return fail(B200REF_ERR_ARG, "bad gather arguments");
  if (in == out) return fail(B200REF_ERR_ARG, "gather_rows cannot run in place");
  if (P == 0 || M == 0) return B200REF_OK;
  if (P > 65535) return fail(B200REF_ERR_ARG, "gather_rows: at most 65535 parameters per call");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned bx = (unsigned)std::min<int64_t>((M + 255) / 256, 1024);
  LAUNCH(K_INGEST, st, k_gather_rows<<<dim3(bx, (unsigned)P), 256, 0, st>>>(in, perm, P, M, out));
  CUDA_TRY(cudaGetLastError());
  return B200REF_OK;
}

int b200ref_synth_fill_f64(double* x, int64_t p0, int64_t P, int64_t C, int64_t N, uint64_t seed, int kind,
                           void* stream) {
  g_err[0] = 0;
  if (!x || P < 0 || C < 1 || N < 1) return fail(B200REF_ERR_ARG, "bad synth arguments");
  if (P == 0) return B200REF_OK;
  const int64_t nchains = P * C;
  const int warps_per_block = 8;
  const unsigned blocks = (unsigned)((nchains + warps_per_block - 1) / warps_per_block);
  LAUNCH(K_SYNTH, (cudaStream_t)stream, k_synth<<<blocks, warps_per_block * 32, 0, (cudaStream_t)stream>>>(x, p0, P, C, N, seed, kind));
  CUDA_TRY(cudaGetLastError());
  return B200REF_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// GLM log-density (logistic regression) on the tcgen05 tensor cores -- see csrc/glm.cuh
// ---------------------------------------------------------------------------------------------
namespace {

// [rows][cols] bf16 row-major; box = GLM_BK columns (one swizzle row) x box_rows rows
int make_map_bf16(CUtensorMap* map, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(B200REF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)GLM_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   GLM_BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200REF_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  return B200REF_OK;
}

// What the workspace holds of the design after a call: lets B200REF_GLM_REUSE_DESIGN skip the fp64 -> bf16 split of X
// (and y / trials) when the caller evaluates new draws against the same design.
struct GlmStaged {
  uint64_t magic;
  const void* X;
  const void* y;
  const void* trials;
  int64_t N, D;
};
constexpr uint64_t GLM_MAGIC = 0xb200f00dd5a17e55ull;

struct GlmLayout {
  size_t bh, bl, xh, xl, yf, af, center, part, hdr, total;
  int Dpad, n_row_tiles, n_draw_tiles, nsplit;
  void compute(int64_t N, int64_t D, int64_t S) {
    Dpad = (int)((D + GLM_KPAD - 1) / GLM_KPAD * GLM_KPAD);
    n_row_tiles = (int)((N + GLM_BN - 1) / GLM_BN);
    n_draw_tiles = (int)((S + GLM_BM - 1) / GLM_BM);
    // row-tile slices per draw tile: the persistent grid (one CTA per SM) runs ceil(items / SMs) waves of
    // ceil(n_row_tiles / nsplit) tiles each.  Few slices keep the CTAs of a wave on the same rows of X (they share
    // them through L2: 17 slices instead of 2 cost 5 % at config 3), so the SMALLEST split within 4 % of the best
    // waves x tiles is taken.  With few draw tiles (a draw-block shard of a multi-GPU run) that fills the last wave:
    // 64 draw tiles x 9 slices = 3.9 waves of 44 tiles, where 8 slices ran 4 waves of 49.
    {
      const int nsm = 148, kmax = std::min(64, n_row_tiles);
      auto cost = [&](int k) { return (((long)n_draw_tiles * k + nsm - 1) / nsm) * ((n_row_tiles + k - 1) / k); };
      long best = cost(1);
      for (int k = 2; k <= kmax; ++k) best = std::min(best, cost(k));
      nsplit = 1;
      while (nsplit < kmax && cost(nsplit) * 100 > best * 104) ++nsplit;
    }
    size_t o = 0;
    bh = o; o += up((size_t)S * Dpad * 2);
    bl = o; o += up((size_t)S * Dpad * 2);
    xh = o; o += up((size_t)N * Dpad * 2);
    xl = o; o += up((size_t)N * Dpad * 2);
    yf = o; o += up((size_t)n_row_tiles * GLM_BN * 4);
    af = o; o += up((size_t)n_row_tiles * GLM_BN * 4);
    center = o; o += up((size_t)Dpad * 8);
    part = o; o += up((size_t)nsplit * GLM_PART_SLOTS * S * 8);
    hdr = o; o += up(sizeof(GlmStaged));
    total = o + 1024;
  }
};

}  // namespace

extern "C" {

size_t b200ref_glm_workspace_bytes(int64_t N, int64_t D, int64_t S) {
  if (N < 1 || D < 1 || S < 1) return 0;
  GlmLayout l;
  l.compute(N, D, S);
  return l.total;
}

int b200ref_glm_lp_f64(int family, const double* X, const double* y, const double* trials, const double* beta,
                       const double* sigma, int64_t N, int64_t D, int64_t S, double* lp, void* ws, size_t ws_bytes,
                       void* stream) {
  return b200ref_glm_lp_ex_f64(family, X, y, trials, beta, sigma, N, D, S, 0, lp, ws, ws_bytes, stream);
}

int b200ref_glm_lp_ex_f64(int family, const double* X, const double* y, const double* trials, const double* beta,
                          const double* sigma, int64_t N, int64_t D, int64_t S, int flags, double* lp, void* ws,
                          size_t ws_bytes, void* stream) {
  g_err[0] = 0;
  if (!X || !y || !beta || !lp || N < 1 || D < 1 || S < 1) return fail(B200REF_ERR_ARG, "bad glm arguments");
  if (family < GLM_BERNOULLI_LOGIT || family > GLM_NORMAL) return fail(B200REF_ERR_ARG, "unknown glm family %d", family);
  if (family == GLM_BINOMIAL_LOGIT && !trials) return fail(B200REF_ERR_ARG, "binomial_logit needs the trials vector");
  if (family == GLM_NORMAL && !sigma) return fail(B200REF_ERR_ARG, "the normal family needs sigma per draw");
  if (D > 65536) return fail(B200REF_ERR_ARG, "D=%lld too large", (long long)D);
  cudaStream_t st = (cudaStream_t)stream;
  GlmLayout L;
  L.compute(N, D, S);
  if (!ws || ws_bytes < L.total) return fail(B200REF_ERR_WORKSPACE, "glm workspace needs %zu bytes", L.total);
  char* w = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16 *bh = (__nv_bfloat16*)(w + L.bh), *bl = (__nv_bfloat16*)(w + L.bl);
  __nv_bfloat16 *xh = (__nv_bfloat16*)(w + L.xh), *xl = (__nv_bfloat16*)(w + L.xl);
  float* yf = (float*)(w + L.yf);
  float* af = (float*)(w + L.af);
  double* part = (double*)(w + L.part);
  double* center = (double*)(w + L.center);
  const int64_t nb = S * L.Dpad, nx = N * L.Dpad, npad = (int64_t)L.n_row_tiles * GLM_BN;
  // the design's staged copies (bf16 split of X, float y / trials) are reused when the caller says this call's
  // design is the previous one's and the workspace header agrees
  GlmStaged want{GLM_MAGIC, X, y, (family == GLM_BINOMIAL_LOGIT) ? trials : nullptr, N, D};
  bool reuse = false;
  if (flags & B200REF_GLM_REUSE_DESIGN) {
    GlmStaged have;
    CUDA_TRY(cudaMemcpyAsync(&have, w + L.hdr, sizeof(have), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    reuse = memcmp(&have, &want, sizeof(have)) == 0;
  }
  // normal family: the squared residual amplifies the error of the split-bf16 contraction, so only the draws'
  // deviations from their mean are contracted and the mean draw's predictor (fp64) is folded into the response
  const bool centered = family == GLM_NORMAL;
  if (centered) LAUNCH(K_GLM_PREP, st, k_glm_center<<<(unsigned)D, 256, 0, st>>>(beta, S, (int)D, center));
  LAUNCH(K_GLM_PREP, st, k_glm_split<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(beta, centered ? center : nullptr, S, (int)D, L.Dpad, bh, bl));
  if (!reuse)
    LAUNCH(K_GLM_PREP, st, k_glm_split<<<(unsigned)((nx + 255) / 256), 256, 0, st>>>(X, nullptr, N, (int)D, L.Dpad, xh, xl));
  if (centered) {
    LAUNCH(K_GLM_PREP, st, k_glm_y_centered<<<(unsigned)((npad + 7) / 8), 256, 0, st>>>(y, X, center, N, (int)D, npad, yf));
  } else if (!reuse) {
    LAUNCH(K_GLM_PREP, st, k_glm_y<<<(unsigned)((npad + 255) / 256), 256, 0, st>>>(y, N, npad, yf));
    if (family == GLM_BINOMIAL_LOGIT)
      LAUNCH(K_GLM_PREP, st, k_glm_y<<<(unsigned)((npad + 255) / 256), 256, 0, st>>>(trials, N, npad, af));
  }
  if (!reuse) {
    want.y = centered ? nullptr : y;  // the staged response of a centred call depends on the draws: never reused
    CUDA_TRY(cudaMemcpyAsync(w + L.hdr, &want, sizeof(want), cudaMemcpyHostToDevice, st));
  }
  CUDA_TRY(cudaMemsetAsync(part, 0, (size_t)L.nsplit * GLM_PART_SLOTS * S * 8, st));
  CUtensorMap mbh, mbl, mxh, mxl;
  int rc;
  if ((rc = make_map_bf16(&mbh, bh, (uint64_t)S, (uint64_t)L.Dpad, GLM_BM))) return rc;
  if ((rc = make_map_bf16(&mbl, bl, (uint64_t)S, (uint64_t)L.Dpad, GLM_BM))) return rc;
  if ((rc = make_map_bf16(&mxh, xh, (uint64_t)N, (uint64_t)L.Dpad, GLM_BN))) return rc;
  if ((rc = make_map_bf16(&mxl, xl, (uint64_t)N, (uint64_t)L.Dpad, GLM_BN))) return rc;
  GlmParams prm;
  prm.N = N; prm.S = S; prm.Dpad = L.Dpad;
  prm.n_row_tiles = L.n_row_tiles; prm.n_draw_tiles = L.n_draw_tiles; prm.nsplit = L.nsplit;
  prm.yf = yf; prm.af = af; prm.lp_part = part;
  static bool attr_dev[MAX_DEVICES] = {};
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(k_glm_lp<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GLM_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_glm_lp<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GLM_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_glm_lp<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GLM_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_glm_lp<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GLM_SMEM_BYTES));
    attr = true;
  }
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(nsm, L.n_draw_tiles * L.nsplit);
  switch (family) {
    case GLM_BERNOULLI_LOGIT: LAUNCH(K_GLM, st, k_glm_lp<0><<<grid, GLM_NT, GLM_SMEM_BYTES, st>>>(mbh, mbl, mxh, mxl, prm)); break;
    case GLM_BINOMIAL_LOGIT: LAUNCH(K_GLM, st, k_glm_lp<1><<<grid, GLM_NT, GLM_SMEM_BYTES, st>>>(mbh, mbl, mxh, mxl, prm)); break;
    case GLM_POISSON_LOG: LAUNCH(K_GLM, st, k_glm_lp<2><<<grid, GLM_NT, GLM_SMEM_BYTES, st>>>(mbh, mbl, mxh, mxl, prm)); break;
    default: LAUNCH(K_GLM, st, k_glm_lp<3><<<grid, GLM_NT, GLM_SMEM_BYTES, st>>>(mbh, mbl, mxh, mxl, prm)); break;
  }
  LAUNCH(K_GLM_PREP, st, k_glm_reduce<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(part, L.nsplit * GLM_PART_SLOTS, S,
                                                                                  family == GLM_NORMAL ? sigma : nullptr, N, lp));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(st));
  return B200REF_OK;
}

int b200ref_glm_logistic_lp_f64(const double* X, const double* y, const double* beta, int64_t N, int64_t D, int64_t S,
                                double* lp, void* ws, size_t ws_bytes, void* stream) {
  return b200ref_glm_lp_f64(GLM_BERNOULLI_LOGIT, X, y, nullptr, beta, nullptr, N, D, S, lp, ws, ws_bytes, stream);
}

}  // extern "C"
