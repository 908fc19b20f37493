// C ABI of libecad_b200.so (see include/ecad_b200.h).  Host-side launch logic only; kernels live in the .cuh files.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <atomic>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges are no-ops unless a tool injects itself

#include "../../include/ecad_b200.h"
#include "attn.cuh"
#include "gemm.cuh"
#include "glue.cuh"
#include "vae.cuh"

namespace {

using namespace ecadk;

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define ECADK_CHECK_CUDA(expr)                                                                   \
  do {                                                                                           \
    cudaError_t err__ = (expr);                                                                  \
    if (err__ != cudaSuccess) return fail(ECADK_ECUDA, "%s: %s", #expr, cudaGetErrorString(err__)); \
  } while (0)

#define ECADK_REQUIRE(cond, ...)                        \
  do {                                                  \
    if (!(cond)) return fail(ECADK_EINVAL, __VA_ARGS__); \
  } while (0)

int check_launch(const char* what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(ECADK_ECUDA, "%s launch failed: %s", what, cudaGetErrorString(err));
  return ECADK_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------------------
// TMA descriptors.  cuTensorMapEncodeTiled is fetched through the runtime so the library has no link-time
// dependency on libcuda (it must load - and export its symbols - on a CPU-only box).
// ---------------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeFn>(ptr);
    }
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  uint64_t rows, cols, pitch;
  uint32_t box_rows, box_cols, swizzle;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && pitch == o.pitch && box_rows == o.box_rows &&
           box_cols == o.box_cols && swizzle == o.swizzle;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.rows); mix(k.cols); mix(k.pitch); mix(k.box_rows); mix(k.box_cols); mix(k.swizzle);
    return h;
  }
};

// Process-wide descriptor cache: a descriptor depends only on (pointer, shape, box), never on contents.
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

// 2-D bf16 tensor [rows, cols] with row pitch `pitch` elements; box = [box_rows, box_cols]; swizzle in bytes.
int make_tmap(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows,
              uint32_t box_cols, uint32_t swizzle_bytes, uint32_t elem_bytes);
int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch,
                   uint32_t box_rows, uint32_t box_cols, uint32_t swizzle_bytes) {
  return make_tmap(out, ptr, rows, cols, pitch, box_rows, box_cols, swizzle_bytes, 2);
}
// elem_bytes: 2 = bf16, 4 = fp32 (the residual stream, stored by the gated-residual epilogue)
int make_tmap(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows,
              uint32_t box_cols, uint32_t swizzle_bytes, uint32_t elem_bytes) {
  TmapKey key{ptr, rows, cols, pitch, box_rows, box_cols, swizzle_bytes | (elem_bytes << 16)};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      return ECADK_OK;
    }
  }
  EncodeFn encode = get_encode_fn();
  if (encode == nullptr) return fail(ECADK_EDRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch * elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = encode(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                      const_cast<void*>(ptr), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(ECADK_EINVAL, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu pitch=%llu box=%ux%u sw=%u",
                static_cast<int>(r), (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch,
                box_rows, box_cols, swizzle_bytes);
  }
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmap_cache.size() > 65536) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *out);
  return ECADK_OK;
}

// 3-D bf16 tensor [rows][mids][inner] (inner contiguous) with box [box_rows][1][box_cols]: the attention operands.
//   head-major  q/k/v [S*H*tokens][80]:              inner = 80, mids = 1
//   row-major   projection output [S*tokens][ld]:    inner = 72 (the head), mids = heads (144-byte stride), rows = S*tokens
// A 16-column box at inner offset 64 of a 72-wide head reads 8 real columns and 8 out-of-bounds ones, which TMA fills
// with zeros - exactly the zero padding the 80-wide shared-memory tiles need.
int make_tmap3_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t mids, uint64_t rows, uint64_t mid_stride,
                    uint64_t row_stride, uint32_t box_rows, uint32_t box_cols, uint32_t swizzle_bytes) {
  TmapKey key{ptr, rows, inner | (mids << 20) | (mid_stride << 32), row_stride, box_rows, box_cols,
              swizzle_bytes | (3u << 24)};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      return ECADK_OK;
    }
  }
  EncodeFn encode = get_encode_fn();
  if (encode == nullptr) return fail(ECADK_EDRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[3] = {inner, mids, rows};
  cuuint64_t gstride[2] = {mid_stride * 2, row_stride * 2};
  cuuint32_t box[3] = {box_cols, 1, box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(ECADK_EINVAL, "cuTensorMapEncodeTiled(3d) failed (%d) inner=%llu mids=%llu rows=%llu strides=%llu,%llu box=%ux%u",
                static_cast<int>(r), (unsigned long long)inner, (unsigned long long)mids, (unsigned long long)rows,
                (unsigned long long)mid_stride, (unsigned long long)row_stride, box_rows, box_cols);
  }
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmap_cache.size() > 65536) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *out);
  return ECADK_OK;
}

// attention operand seen through a 3-D map: ld == 0 -> head-major [S*H*tokens][80]; else row-major [S*tokens][ld]
int make_attn_tmap(CUtensorMap* out, const void* ptr, int ld, int samples, int heads, int tokens, uint32_t box_rows,
                   uint32_t box_cols, uint32_t swizzle_bytes) {
  if (ld == 0)
    return make_tmap3_bf16(out, ptr, kHeadPad, 1, static_cast<uint64_t>(samples) * heads * tokens, kHeadPad, kHeadPad,
                           box_rows, box_cols, swizzle_bytes);
  return make_tmap3_bf16(out, ptr, kHeadDim, heads, static_cast<uint64_t>(samples) * tokens, kHeadDim, ld, box_rows,
                         box_cols, swizzle_bytes);
}

// ---------------------------------------------------------------------------------------------------
// in-situ profiler: CUDA event pairs around every launch while enabled
// ---------------------------------------------------------------------------------------------------
struct ProfRec {
  int cls;
  size_t ev;  // index of the start event; stop = ev + 1
  double flops, bytes;
};
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> events;
  size_t used = 0;
  std::vector<ProfRec> recs;
};
ProfState g_prof;

struct ProfScope {
  cudaStream_t stream;
  bool active;
  ProfScope(int cls, double flops, double bytes, cudaStream_t s) : stream(s), active(g_prof.on) {
    if (!active) return;
    while (g_prof.events.size() < g_prof.used + 2) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) {
        active = false;
        return;
      }
      g_prof.events.push_back(e);
    }
    g_prof.recs.push_back(ProfRec{cls, g_prof.used, flops, bytes});
    cudaEventRecord(g_prof.events[g_prof.used], stream);
    g_prof.used += 2;
  }
  ~ProfScope() {
    if (active) cudaEventRecord(g_prof.events[g_prof.recs.back().ev + 1], stream);
  }
};

// ---------------------------------------------------------------------------------------------------
// NVTX ranges (one range per executor call and per sub-block, so a timeline or an
// `ncu --nvtx --nvtx-include` capture can be cut at the reference's attn1 / attn2 / ff granularity).
// Off by default (a push/pop pair per sub-block is host work on the launch path); ECADK_NVTX=1 or
// ecadk_set_nvtx(1) turns them on.
// ---------------------------------------------------------------------------------------------------
std::atomic<int> g_nvtx{-1};
bool nvtx_on() {
  int v = g_nvtx.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("ECADK_NVTX");
    v = (e != nullptr && atoi(e) != 0) ? 1 : 0;
    g_nvtx.store(v, std::memory_order_relaxed);
  }
  return v > 0;
}
std::atomic<long long> g_nvtx_ranges{0};
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* fmt, ...) : on(nvtx_on()) {
    if (!on) return;
    char buf[96];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    nvtxRangePushA(buf);
    g_nvtx_ranges.fetch_add(1, std::memory_order_relaxed);
  }
  ~NvtxRange() {
    if (on) nvtxRangePop();
  }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// ---------------------------------------------------------------------------------------------------
// Launch with programmatic stream serialization (PDL): the kernel may be scheduled while its predecessor is still
// running; every kernel launched this way executes griddepcontrol.wait before its first global-memory access.
// ECADK_PDL=0 falls back to plain stream order.
// ---------------------------------------------------------------------------------------------------
bool use_pdl() {
  static const bool on = [] {
    const char* e = getenv("ECADK_PDL");
    return !(e != nullptr && atoi(e) == 0);
  }();
  return on;
}

template <typename... KArgs, typename... Args>
void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------------
// GEMM launcher
// ---------------------------------------------------------------------------------------------------
// tensor maps of the gated-residual epilogue's bulk stores (dummies for the other epilogues)
struct EpiMaps {
  CUtensorMap x, cache, xb;
};

template <int BN, int EPI>
int launch_gemm_inst(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const EpiMaps& em,
                     const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPI>;
  static bool configured = false;
  auto kern = gemm_bf16_kernel<BN, EPI>;
  if (!configured) {
    ECADK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = ((p.M + kGemmBM - 1) / kGemmBM) * (p.N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  launch_pdl(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, ta2, tb, em.x, em.cache, em.xb, p);
  return check_launch("gemm_bf16_kernel");
}

// Split-K workspace of the calling thread: the executor entry points install their handle's buffer for the duration
// of the call (SplitWsScope); ecadk_set_splitk_workspace installs one for stand-alone GEMM calls.  No workspace, no
// split: launch_gemm then takes the ordinary kernels.
struct SplitWs {
  void* ptr;
  size_t bytes;
};
thread_local SplitWs g_split_ws = {nullptr, 0};
std::atomic<long long> g_splitk_launches{0};  // process-wide count of gemm_splitk_kernel launches (tests, bench)
struct SplitWsScope {
  SplitWs saved;
  SplitWsScope(void* ptr, size_t bytes) : saved(g_split_ws) {
    if (ptr != nullptr) g_split_ws = SplitWs{ptr, bytes};
  }
  ~SplitWsScope() { g_split_ws = saved; }
};
constexpr size_t kSplitWsBytes = 16u << 20;  // per handle: 36 tiles x 4 CTAs x 64 KB = 9.4 MB at M = 512, N = 1152

template <typename... KArgs, typename... Args>
void launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, int cluster, cudaStream_t stream,
                        Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// clusters of `split` CTAs of gemm_splitk_kernel<BN, EPI> that can be resident at once (0 on error); cached
template <int BN, int EPI>
int splitk_max_clusters(int split) {
  using Cfg = GemmCfg<BN, EPI>;
  static int cached[kSplitMax + 1] = {-1, -1, -1, -1, -1};
  if (cached[split] >= 0) return cached[split];
  auto kern = gemm_splitk_kernel<BN, EPI>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) {
    cudaGetLastError();
    return cached[split] = 0;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(split * num_sms());
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = split;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return cached[split] = n;
}

template <int BN, int EPI>
int launch_gemm_splitk_inst(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const EpiMaps& em,
                            const GemmParams& p, int split, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPI>;
  const int tiles = ((p.M + kGemmBM - 1) / kGemmBM) * (p.N / BN);
  launch_pdl_cluster(gemm_splitk_kernel<BN, EPI>, dim3(tiles * split), dim3(kGemmThreads), Cfg::kSmemBytes, split, stream,
                     ta, ta2, tb, em.x, em.cache, em.xb, p, static_cast<float4*>(g_split_ws.ptr));
  g_splitk_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch("gemm_splitk_kernel");
}

// Split-K plan for a small problem.  A GEMM with fewer tiles than SMs is bound by ONE CTA's operand stream through its
// SM's L2 port - (128 + BN) * (K / split) * 2 bytes at ~88 B/ns - plus ~3.2 us of fixed cost, not by MMAs, HBM or the
// chip-wide L2 rate (tools/micro/small_gemm_cold_hot.py, M = 512, PDL-chained graph replays, weights cold / in L2):
//     FF2  K = 4608:  BN 64 unsplit 19.6 / 18.4 us;  BN 64 x 2  12.8 / 12.2;  BN 128 x 3  12.2 / 11.4
//     out  K = 1152:  BN 64 unsplit  8.2 /  7.7 us;  BN 64 x 2   7.1 /  6.8;  BN 128 x 3   8.1 /  7.6
// The reduction costs ~1.3 us when every warp moves one chunk (BN 64 x 2) and ~2.7 us at BN 128 x 3 (three chunks out,
// two partials in), which eats the smaller stream at K = 1152; BN 128 x 4 would stream a third of BN 64 x 2 but only 33
// clusters of 4 are resident on 148 SMs (GPC granularity) and a tile row needs 36.  The plan therefore considers the
// two-way splits only; ecadk_set_splitk_force overrides it (tests and measurements cover 3 and 4).
struct SplitForce {
  int bn, split;
};
SplitForce g_split_force = {0, 0};

inline double splitk_cost_ns(int K, int bn, int split) {
  return (128.0 + bn) * (static_cast<double>(K) / split) * 2.0 / 88.0 + (split > 1 ? 1300.0 : 0.0);
}

template <int EPI>
int splitk_resident(int bn, int split) {
  return bn == 128 ? splitk_max_clusters<128, EPI>(split) : splitk_max_clusters<64, EPI>(split);
}

template <int EPI>
void plan_splitk(const GemmParams& p, int* bn_out, int* split_out) {
  *split_out = 1;
  static const bool allow = [] {
    const char* e = getenv("ECADK_GEMM_SPLITK");
    return !(e != nullptr && atoi(e) == 0);
  }();
  if (!allow || g_split_ws.ptr == nullptr) return;
  const int m_tiles = (p.M + kGemmBM - 1) / kGemmBM;
  const int num_kb = p.K / kGemmBK;
  if (2 * m_tiles * (p.N / 128) > num_sms()) return;  // enough 128-wide tiles for the ordinary kernels
  auto feasible = [&](int bn, int split) {
    if ((bn != 64 && bn != 128) || split < 2 || split > kSplitMax) return false;
    if (p.N % bn != 0 || num_kb < 2 * split || (bn / 32) < split) return false;
    const int tiles = m_tiles * (p.N / bn);
    if (static_cast<size_t>(tiles) * split * bn * 128 * sizeof(float) > g_split_ws.bytes) return false;
    return tiles <= splitk_resident<EPI>(bn, split);
  };
  if (g_split_force.split != 0) {
    if (feasible(g_split_force.bn, g_split_force.split)) {
      *bn_out = g_split_force.bn;
      *split_out = g_split_force.split;
    }
    return;
  }
  double best = splitk_cost_ns(p.K, 64, 1);  // what the ordinary path would do with this problem (see launch_gemm)
  const int cand[2][2] = {{64, 2}, {128, 2}};
  for (const auto& c : cand) {
    const double cost = splitk_cost_ns(p.K, c[0], c[1]);
    if (cost >= best || !feasible(c[0], c[1])) continue;
    best = cost;
    *bn_out = c[0];
    *split_out = c[1];
  }
  static const bool debug = getenv("ECADK_DEBUG_PLAN") != nullptr;
  if (debug) {
    fprintf(stderr, "[ecadk] gemm EPI %d M=%d N=%d K=%d -> BN=%d split=%d (resident clusters of 2/3/4: %d/%d/%d)\n", EPI,
            p.M, p.N, p.K, *split_out > 1 ? *bn_out : 0, *split_out, splitk_resident<EPI>(128, 2),
            splitk_resident<EPI>(128, 3), splitk_resident<EPI>(128, 4));
  }
}

template <int BN, int EPI>
int launch_gemm2_inst(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const CUtensorMap& tb_tail,
                      const EpiMaps& em, const GemmParams& p, int tail, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN, EPI>;
  static bool configured = false;
  auto kern = gemm2_bf16_kernel<BN, EPI>;
  if (!configured) {
    ECADK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = ((p.M + 2 * kGemmBM - 1) / (2 * kGemmBM)) * ((p.N - tail) / BN + (tail ? 1 : 0));
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  launch_pdl(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, ta2, tb, tb_tail, em.x, em.cache, em.xb, p,
             tail);
  return check_launch("gemm2_bf16_kernel");
}

// 3x3 convolution with C_out = 128: the taps of one kernel row share their A tile (gemm2_bf16_kernel, TAP3)
template <int EPI>
int launch_gemm2_tap3_inst(const CUtensorMap& ta, const CUtensorMap& tb, const EpiMaps& em, const GemmParams& p,
                           cudaStream_t stream) {
  using Cfg = Gemm2Cfg<128, EPI, true>;
  static bool configured = false;
  auto kern = gemm2_bf16_kernel<128, EPI, true>;
  if (!configured) {
    ECADK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int rows_pair = 2 * Cfg::kSub * kGemmBM;
  const int tiles = ((p.M + rows_pair - 1) / rows_pair) * (p.N / 128);
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  launch_pdl(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, ta, tb, tb, em.x, em.cache, em.xb, p, 0);
  return check_launch("gemm2_bf16_kernel<TAP3>");
}

// Tile-N choice: among the widths that divide N, minimise waves x per-tile time (~ BN + fixed overhead).
int pick_bn(int m_tiles, int N, int workers) {
  const int cands[3] = {256, 192, 128};
  int best = 0;
  double best_cost = 1e30;
  for (int bn : cands) {
    if (N % bn) continue;
    const int tiles = m_tiles * (N / bn);
    const int waves = (tiles + workers - 1) / workers;
    const double cost = static_cast<double>(waves) * (bn + 24.0);
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

// ECADK_GEMM_CTA_GROUP=1|2 forces the 1-CTA / 2-CTA kernel (A/B measurements); default: 2-CTA when the problem has
// enough 256-row tiles to fill every CTA pair, else the 1-CTA kernel (small-batch latency configuration).
int gemm_cta_group(int M, int N) {
  static int forced = [] {
    const char* e = getenv("ECADK_GEMM_CTA_GROUP");
    return e ? atoi(e) : 0;
  }();
  if (forced == 1 || forced == 2) return forced;
  const int m2 = (M + 2 * kGemmBM - 1) / (2 * kGemmBM);
  return (m2 * (N / 128) >= num_sms() / 2) ? 2 : 1;
}

template <int EPI>
int launch_gemm(const void* a, const void* w, GemmParams& p, cudaStream_t stream) {
  ECADK_REQUIRE(a && w, "gemm: null operand");
  ECADK_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: bad shape M=%d N=%d K=%d", p.M, p.N, p.K);
  ECADK_REQUIRE(p.K % kGemmBK == 0, "gemm: K=%d must be a multiple of %d", p.K, kGemmBK);
  ECADK_REQUIRE(p.N % 128 == 0, "gemm: N=%d must be a multiple of 128", p.N);
  ECADK_REQUIRE(aligned16(a) && aligned16(w), "gemm: operands must be 16-byte aligned");
  const int group = gemm_cta_group(p.M, p.N);
  ProfScope prof(ECADK_PROF_GEMM, 2.0 * p.M * p.N * p.K, 0.0, stream);
  CUtensorMap ta, ta2, tb;
  int rc;
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  if constexpr (EPI == EPI_GATED_RESIDUAL && ECADK_EPI_TMA_STORE) {
    // bulk-store maps of one 32 x 32 chunk: fp32 residual stream (128-byte rows, 128B swizzle), bf16 cache / shadow
    // (64-byte rows, 64B swizzle)
    ECADK_REQUIRE(p.x != nullptr && aligned16(p.x), "gemm: residual stream must be 16-byte aligned");
    if ((rc = make_tmap(&em.x, p.x, p.M, p.N, p.N, 32, 32, 128, 4))) return rc;
    em.cache = em.x;
    em.xb = em.x;
    if (p.cache != nullptr && (rc = make_tmap(&em.cache, p.cache, p.M, p.N, p.N, 32, 32, 64, 2))) return rc;
    if (p.xb != nullptr && (rc = make_tmap(&em.xb, p.xb, p.M, p.N, p.N, 32, 32, 64, 2))) return rc;
  }
  if (p.a2 != nullptr) {  // A = [a (k1 columns) | a2 (K - k1 columns)], each with its own row pitch
    ECADK_REQUIRE(p.k1 > 0 && p.k1 < p.K && p.k1 % kGemmBK == 0 && aligned16(p.a2),
                  "gemm: split A needs 0 < k1=%d < K=%d, a multiple of %d", p.k1, p.K, kGemmBK);
    if ((rc = make_tmap_bf16(&ta, a, p.M, p.k1, p.k1, kGemmBM, kGemmBK, 128))) return rc;
    if ((rc = make_tmap_bf16(&ta2, p.a2, p.M, p.K - p.k1, p.K - p.k1, kGemmBM, kGemmBK, 128))) return rc;
    p.kb_split = p.k1 / kGemmBK;
  } else {
    // implicit-GEMM convolution: the A matrix is the bordered NHWC image [M, C], K = taps * C (see GemmParams)
    const int a_cols = p.conv_taps > 1 ? p.K / p.conv_taps : p.K;
    if ((rc = make_tmap_bf16(&ta, a, p.M, a_cols, a_cols, kGemmBM, kGemmBK, 128))) return rc;
    ta2 = ta;
    p.kb_split = p.K / kGemmBK;
  }
  if constexpr (EPI == EPI_CONV) {
    // ECADK_CONV_TAP3=0 keeps the nine-loads-per-block form (A/B measurements)
    static const bool tap3 = [] {
      const char* e = getenv("ECADK_CONV_TAP3");
      return !(e != nullptr && atoi(e) == 0);
    }();
    if (tap3 && group == 2 && p.conv_taps == 9 && p.N == 128) {
      const int c_in = p.K / 9;
      if ((rc = make_tmap_bf16(&ta, a, p.M, c_in, c_in, kTap3Rows, kGemmBK, 128))) return rc;
      if ((rc = make_tmap_bf16(&tb, w, p.N, p.K, p.K, 64, kGemmBK, 128))) return rc;
      return launch_gemm2_tap3_inst<EPI>(ta, tb, em, p, stream);
    }
  }
  if (group == 2) {
    // ECADK_GEMM_TAIL=0 disables the 256-wide + 128-wide-tail tiling (A/B measurements)
    static const bool use_tail = [] {
      const char* e = getenv("ECADK_GEMM_TAIL");
      return !(e != nullptr && atoi(e) == 0);
    }();
    if (use_tail && p.N > 256 && p.N % 256 == 128) {
      // N = k*256 + 128 (1152, 3456): k full-width tiles + one 128-wide tail tile per row block
      CUtensorMap tb_tail;
      if ((rc = make_tmap_bf16(&tb, w, p.N, p.K, p.K, 128, kGemmBK, 128))) return rc;
      if ((rc = make_tmap_bf16(&tb_tail, w, p.N, p.K, p.K, 64, kGemmBK, 128))) return rc;
      return launch_gemm2_inst<256, EPI>(ta, ta2, tb, tb_tail, em, p, 128, stream);
    }
    const int bn = pick_bn((p.M + 2 * kGemmBM - 1) / (2 * kGemmBM), p.N, num_sms() / 2);
    ECADK_REQUIRE(bn != 0, "gemm: no tile width divides N=%d", p.N);
    if ((rc = make_tmap_bf16(&tb, w, p.N, p.K, p.K, bn / 2, kGemmBK, 128))) return rc;
    switch (bn) {
      case 256: return launch_gemm2_inst<256, EPI>(ta, ta2, tb, tb, em, p, 0, stream);
      case 192: return launch_gemm2_inst<192, EPI>(ta, ta2, tb, tb, em, p, 0, stream);
      default: return launch_gemm2_inst<128, EPI>(ta, ta2, tb, tb, em, p, 0, stream);
    }
  }
  int bn = pick_bn((p.M + kGemmBM - 1) / kGemmBM, p.N, num_sms());
  ECADK_REQUIRE(bn != 0, "gemm: no tile width divides N=%d", p.N);
  {
    int split = 1, sbn = 0;
    plan_splitk<EPI>(p, &sbn, &split);
    if (split > 1) {
      if ((rc = make_tmap_bf16(&tb, w, p.N, p.K, p.K, sbn, kGemmBK, 128))) return rc;
      return sbn == 128 ? launch_gemm_splitk_inst<128, EPI>(ta, ta2, tb, em, p, split, stream)
                        : launch_gemm_splitk_inst<64, EPI>(ta, ta2, tb, em, p, split, stream);
    }
  }
  // small problems (batch-1 latency configuration): when even 128-wide tiles leave more than half of the SMs idle,
  // 64-wide tiles double the CTAs - each streams its own W slice, so the per-SM operand stream (what bounds a lone
  // 128 x K strip, e.g. FF2 at M = 512: 2.3 MB through each of 36 SMs) halves.  ECADK_GEMM_BN64=0 disables it.
  static const bool allow64 = [] {
    const char* e = getenv("ECADK_GEMM_BN64");
    return !(e != nullptr && atoi(e) == 0);
  }();
  if (allow64 && bn == 128 && p.N % 64 == 0 && 2 * ((p.M + kGemmBM - 1) / kGemmBM) * (p.N / 128) <= num_sms()) bn = 64;
  if ((rc = make_tmap_bf16(&tb, w, p.N, p.K, p.K, bn, kGemmBK, 128))) return rc;
  switch (bn) {
    case 64: return launch_gemm_inst<64, EPI>(ta, ta2, tb, em, p, stream);
    case 256: return launch_gemm_inst<256, EPI>(ta, ta2, tb, em, p, stream);
    case 192: return launch_gemm_inst<192, EPI>(ta, ta2, tb, em, p, stream);
    default: return launch_gemm_inst<128, EPI>(ta, ta2, tb, em, p, stream);
  }
}

// ---------------------------------------------------------------------------------------------------
// attention launcher
// ---------------------------------------------------------------------------------------------------
template <int NK, bool HAS_BIAS>
int launch_attn_inst(const CUtensorMap* tm, const AttnParams& p, int samples, cudaStream_t stream) {
  using Cfg = AttnCfg<NK>;
  static bool configured = false;
  auto kern = attn_tile_kernel<NK, HAS_BIAS>;
  if (!configured) {
    ECADK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  dim3 grid(p.q_tokens / kAttnBM, samples * p.heads);
  kern<<<grid, kAttnThreads, Cfg::kSmemBytes, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
  return check_launch("attn_tile_kernel");
}

template <int NK, bool HAS_BIAS>
int launch_attn_pair_inst(const CUtensorMap* tq, const CUtensorMap* tm, const AttnParams& p, int items,
                          cudaStream_t stream) {
  using Cfg = AttnPairCfg<NK>;
  static bool configured = false;
  auto kern = attn_pair_kernel<NK, HAS_BIAS>;
  if (!configured) {
    ECADK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int grid = items < num_sms() ? items : num_sms();
  launch_pdl(kern, dim3(grid), dim3(kAttnPairThreads), Cfg::kSmemBytes, stream, tq[0], tq[1], tm[2], tm[3], tm[4], tm[5], p,
             items);
  return check_launch("attn_pair_kernel");
}

template <int NK, bool HAS_BIAS>
int launch_attn_pair2_inst(const CUtensorMap* tq, const CUtensorMap* tm, const CUtensorMap& to, const AttnParams& p,
                           int items, cudaStream_t stream) {
  using Cfg = AttnPair2Cfg<NK, HAS_BIAS>;
  static bool configured = false;
  auto kern = attn_pair2_kernel<NK, HAS_BIAS>;
  if (!configured) {
    ECADK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int grid = items < num_sms() ? items : num_sms();
  launch_pdl(kern, dim3(grid), dim3(kPair2Threads), Cfg::kSmemBytes, stream, tq[0], tq[1], tm[2], tm[3], tm[5], to, p,
             items);
  return check_launch("attn_pair2_kernel");
}

template <int HD, bool HAS_BIAS>
int launch_attn_flash_inst(const CUtensorMap* tq, const CUtensorMap* tkv, const AttnParams& p, int n_keys, int items,
                           cudaStream_t stream) {
  using Cfg = AttnFlashCfg<HD>;
  static bool configured = false;
  auto kern = attn_flash_kernel<HD, HAS_BIAS>;
  if (!configured) {
    ECADK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int grid = items < num_sms() ? items : num_sms();
  launch_pdl(kern, dim3(grid), dim3(kFlashThreads), Cfg::kSmemBytes, stream, tq[0], tq[1], tkv[0], tkv[1], tkv[2], tkv[3], p,
             n_keys, items);
  return check_launch("attn_flash_kernel");
}

// second-generation streaming kernel (head dim 72): next Q K^T issued while the block's softmax runs
template <bool HAS_BIAS>
int launch_attn_flash2_inst(const CUtensorMap* tq, const CUtensorMap* tkv, const AttnParams& p, int n_keys, int items,
                            cudaStream_t stream) {
  using Cfg = AttnFlash2Cfg;
  static bool configured = false;
  auto kern = attn_flash2_kernel<HAS_BIAS>;
  if (!configured) {
    ECADK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int grid = items < num_sms() ? items : num_sms();
  launch_pdl(kern, dim3(grid), dim3(kFlash2Threads), Cfg::kSmemBytes, stream, tq[0], tq[1], tkv[0], tkv[1], tkv[3], p,
             n_keys, items);
  return check_launch("attn_flash2_kernel");
}

// FLUX joint attention: head_dim 128 (no padding), no key bias; out row pitch given (the single-stream blocks write
// straight into the [attn | mlp] concat buffer).
int launch_attention_d128(const void* q, const void* k, const void* v, void* out, int out_ld, void* out_lo,
                          int split_tokens, int samples, int heads, int q_tokens, int n_keys, cudaStream_t stream) {
  ECADK_REQUIRE(split_tokens == 0 || (out_lo != nullptr && split_tokens > 0 && split_tokens < q_tokens),
                "attention_d128: split_tokens=%d needs out_lo and 0 < split < q_tokens", split_tokens);
  ECADK_REQUIRE(q && k && v && out, "attention_d128: null pointer");
  ECADK_REQUIRE(n_keys > 0 && n_keys % 128 == 0, "attention_d128: n_keys=%d must be a multiple of 128", n_keys);
  ECADK_REQUIRE(q_tokens > 0 && q_tokens % 256 == 0, "attention_d128: q_tokens=%d must be a multiple of 256", q_tokens);
  ECADK_REQUIRE(out_ld >= heads * 128 && out_ld % 8 == 0, "attention_d128: out_ld=%d", out_ld);
  ECADK_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out), "attention_d128: 16-byte alignment");
  const uint64_t q_rows = static_cast<uint64_t>(samples) * heads * q_tokens;
  const uint64_t k_rows = static_cast<uint64_t>(samples) * heads * n_keys;
  ProfScope prof(ECADK_PROF_ATTENTION, 4.0 * samples * heads * q_tokens * n_keys * 128.0, 0.0, stream);
  AttnParams p;
  p.heads = heads;
  p.q_tokens = q_tokens;
  p.out_ld = out_ld;
  p.scale_log2e = static_cast<float>(1.4426950408889634 / std::sqrt(128.0));
  p.bias = nullptr;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out_lo = static_cast<__nv_bfloat16*>(out_lo);
  p.split_tokens = split_tokens;
  CUtensorMap tq[2], tkv[4];
  int rc;
  // the streaming kernel addresses its operands through 3-D maps (see make_tmap3_bf16): head-major, middle extent 1
  p.q_rowmajor = 0;
  p.kv_rowmajor = 0;
  if ((rc = make_tmap3_bf16(&tq[0], q, 128, 1, q_rows, 128, 128, 256, 64, 128))) return rc;
  tq[1] = tq[0];
  if ((rc = make_tmap3_bf16(&tkv[0], k, 128, 1, k_rows, 128, 128, kFlashKB, 64, 128))) return rc;
  tkv[1] = tkv[0];
  if ((rc = make_tmap3_bf16(&tkv[2], v, 128, 1, k_rows, 128, 128, kFlashKB, 64, 128))) return rc;
  tkv[3] = tkv[2];
  const int items = samples * heads * (q_tokens / 256);
  return launch_attn_flash_inst<128, false>(tq, tkv, p, n_keys, items, stream);
}

int launch_attention_ex(const void* q, int q_ld, const void* k, const void* v, int kv_ld, const float* bias, void* out,
                        int samples, int heads, int q_tokens, int n_keys, cudaStream_t stream);
int launch_attention(const void* q, const void* k, const void* v, const float* bias, void* out, int samples,
                     int heads, int q_tokens, int n_keys, cudaStream_t stream) {
  return launch_attention_ex(q, 0, k, v, 0, bias, out, samples, heads, q_tokens, n_keys, stream);
}

// q_ld / kv_ld: 0 = head-major operand [S, H, tokens, 80]; > 0 = row-major [S*tokens, ld] (the plain output of the
// projection GEMM; the pointer addresses the first column of head 0 of THIS operand, e.g. qkv + 1152 for K)
int launch_attention_ex(const void* q, int q_ld, const void* k, const void* v, int kv_ld, const float* bias, void* out,
                        int samples, int heads, int q_tokens, int n_keys, cudaStream_t stream) {
  ECADK_REQUIRE(q && k && v && out, "attention: null pointer");
  ECADK_REQUIRE(q_ld == 0 || (q_ld >= heads * kHeadDim && q_ld % 8 == 0), "attention: q_ld=%d", q_ld);
  ECADK_REQUIRE(kv_ld == 0 || (kv_ld >= heads * kHeadDim && kv_ld % 8 == 0), "attention: kv_ld=%d", kv_ld);
  ECADK_REQUIRE(n_keys > 0 && n_keys % 128 == 0, "attention: n_keys=%d must be a multiple of 128", n_keys);
  ECADK_REQUIRE(q_tokens > 0 && q_tokens % kAttnBM == 0, "attention: q_tokens=%d must be a multiple of 128", q_tokens);
  ECADK_REQUIRE(n_keys <= 256 || q_tokens % 256 == 0,
                "attention: q_tokens=%d must be a multiple of 256 when n_keys > 256", q_tokens);
  ECADK_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out), "attention: 16-byte alignment");
  const uint64_t q_rows = static_cast<uint64_t>(samples) * heads * q_tokens;
  const uint64_t k_rows = static_cast<uint64_t>(samples) * heads * n_keys;
  ProfScope prof(ECADK_PROF_ATTENTION, 4.0 * samples * heads * q_tokens * n_keys * kHeadDim, 0.0, stream);
  // ECADK_ATTN_MODE=tile / flash force the one-tile-per-CTA / the streaming kernel (A/B measurements)
  static const int forced = [] {
    const char* e = getenv("ECADK_ATTN_MODE");
    if (e == nullptr) return 0;
    return strcmp(e, "tile") == 0 ? 1 : (strcmp(e, "flash") == 0 ? 2 : (strcmp(e, "pair1") == 0 ? 3 : 0));
  }();
  AttnParams p;
  p.heads = heads;
  p.q_tokens = q_tokens;
  p.out_ld = heads * kHeadDim;
  p.scale_log2e = static_cast<float>(1.4426950408889634 / std::sqrt(static_cast<double>(kHeadDim)));
  p.bias = bias;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out_lo = nullptr;
  p.split_tokens = 0;
  p.q_rowmajor = q_ld != 0;
  p.kv_rowmajor = kv_ld != 0;
  int rc;
  // (256 keys WITH a key bias on row-major operands - a padded self-attention of fewer than 256 tokens - also streams:
  // the 256-query kernel has no shared memory left for bias slices and the first-generation one reads head-major only)
  const bool biased256 = bias != nullptr && n_keys == 256 && (q_ld != 0 || kv_ld != 0);
  const bool use_flash = (n_keys > 256 || q_tokens > 256 || forced == 2 || biased256) && q_tokens % 256 == 0;
  if (use_flash) {
    // long key sequences: stream 128-key blocks with an online softmax
    CUtensorMap tq[2], tkv[4];
    if ((rc = make_attn_tmap(&tq[0], q, q_ld, samples, heads, q_tokens, 256, 64, 128))) return rc;
    if ((rc = make_attn_tmap(&tq[1], q, q_ld, samples, heads, q_tokens, 256, 16, 32))) return rc;
    if ((rc = make_attn_tmap(&tkv[0], k, kv_ld, samples, heads, n_keys, kFlashKB, 64, 128))) return rc;
    if ((rc = make_attn_tmap(&tkv[1], k, kv_ld, samples, heads, n_keys, kFlashKB, 16, 32))) return rc;
    if ((rc = make_attn_tmap(&tkv[2], v, kv_ld, samples, heads, n_keys, kFlashKB, 64, 128))) return rc;
    if ((rc = make_attn_tmap(&tkv[3], v, kv_ld, samples, heads, n_keys, kFlashKB, 16, 32))) return rc;
    const int items = samples * heads * (q_tokens / 256);
    // Long key sequences run the second-generation kernel (next Q K^T issued during the block's softmax): measured
    // 1471 against 1638 us at 4096 keys, but 225 against 216 us at 1024 keys and 311 against 245 us at 384 keys, where
    // its longer item prologue / epilogue shows (profiles/r2_attention_times_final.txt).  ECADK_ATTN_FLASH=1 / 2 force
    // the first / second generation (A/B measurements).
    static const int gen = [] {
      const char* e = getenv("ECADK_ATTN_FLASH");
      return e == nullptr ? 0 : atoi(e);
    }();
    if (gen == 2 || (gen != 1 && n_keys >= 2048)) {
      return bias ? launch_attn_flash2_inst<true>(tq, tkv, p, n_keys, items, stream)
                  : launch_attn_flash2_inst<false>(tq, tkv, p, n_keys, items, stream);
    }
    return bias ? launch_attn_flash_inst<72, true>(tq, tkv, p, n_keys, items, stream)
                : launch_attn_flash_inst<72, false>(tq, tkv, p, n_keys, items, stream);
  }
  const bool pair2 = q_tokens == 256 && forced != 1 && forced != 3 && !(n_keys == 256 && bias != nullptr);
  ECADK_REQUIRE(pair2 || (q_ld == 0 && kv_ld == 0),
                "attention: row-major operands need the 256-query or the streaming kernel (q_tokens=%d n_keys=%d)",
                q_tokens, n_keys);
  if (pair2) {
    // second-generation 256-query kernel: per-tile Q slots, per-item K/V double buffering, output through smem + TMA
    // store (256 keys WITH a key bias - text lengths in (128, 256], no shipped configuration - stays on the first kernel:
    // the second one has no shared memory left for the bias slices there)
    CUtensorMap tq1[2], tk[6], to;
    if ((rc = make_attn_tmap(&tq1[0], q, q_ld, samples, heads, q_tokens, kAttnBM, 64, 128))) return rc;
    if ((rc = make_attn_tmap(&tq1[1], q, q_ld, samples, heads, q_tokens, kAttnBM, 16, 32))) return rc;
    if ((rc = make_attn_tmap(&tk[2], k, kv_ld, samples, heads, n_keys, n_keys, 64, 128))) return rc;
    if ((rc = make_attn_tmap(&tk[3], k, kv_ld, samples, heads, n_keys, n_keys, 16, 32))) return rc;
    if ((rc = make_attn_tmap(&tk[5], v, kv_ld, samples, heads, n_keys, n_keys, 16, 32))) return rc;
    if ((rc = make_tmap_bf16(&to, out, static_cast<uint64_t>(samples) * q_tokens, heads * kHeadDim, p.out_ld, kAttnBM,
                             kHeadDim, 0)))
      return rc;
    const int items = samples * heads;
    if (n_keys == 256) return launch_attn_pair2_inst<256, false>(tq1, tk, to, p, items, stream);
    return bias ? launch_attn_pair2_inst<128, true>(tq1, tk, to, p, items, stream)
                : launch_attn_pair2_inst<128, false>(tq1, tk, to, p, items, stream);
  }
  CUtensorMap tm[6];
  if ((rc = make_tmap_bf16(&tm[0], q, q_rows, kHeadPad, kHeadPad, kAttnBM, 64, 128))) return rc;
  if ((rc = make_tmap_bf16(&tm[1], q, q_rows, kHeadPad, kHeadPad, kAttnBM, 16, 32))) return rc;
  if ((rc = make_tmap_bf16(&tm[2], k, k_rows, kHeadPad, kHeadPad, n_keys, 64, 128))) return rc;
  if ((rc = make_tmap_bf16(&tm[3], k, k_rows, kHeadPad, kHeadPad, n_keys, 16, 32))) return rc;
  if ((rc = make_tmap_bf16(&tm[4], v, k_rows, kHeadPad, kHeadPad, n_keys, 64, 128))) return rc;
  if ((rc = make_tmap_bf16(&tm[5], v, k_rows, kHeadPad, kHeadPad, n_keys, 16, 32))) return rc;
  if (q_tokens == 256 && forced != 1) {
    CUtensorMap tq[2];
    if ((rc = make_tmap_bf16(&tq[0], q, q_rows, kHeadPad, kHeadPad, 256, 64, 128))) return rc;
    if ((rc = make_tmap_bf16(&tq[1], q, q_rows, kHeadPad, kHeadPad, 256, 16, 32))) return rc;
    const int items = samples * heads;
    if (n_keys == 256) {
      return bias ? launch_attn_pair_inst<256, true>(tq, tm, p, items, stream)
                  : launch_attn_pair_inst<256, false>(tq, tm, p, items, stream);
    }
    return bias ? launch_attn_pair_inst<128, true>(tq, tm, p, items, stream)
                : launch_attn_pair_inst<128, false>(tq, tm, p, items, stream);
  }
  if (n_keys == 256) {
    return bias ? launch_attn_inst<256, true>(tm, p, samples, stream) : launch_attn_inst<256, false>(tm, p, samples, stream);
  }
  return bias ? launch_attn_inst<128, true>(tm, p, samples, stream) : launch_attn_inst<128, false>(tm, p, samples, stream);
}

int launch_residual_ln(const EcadkResidualLnArgs& a, cudaStream_t stream) {
  ECADK_REQUIRE(a.x != nullptr && a.rows > 0 && a.tokens > 0, "residual_ln: bad x/rows/tokens");
  ECADK_REQUIRE(a.dim == 512 || a.dim == 1152 || a.dim == 3072, "residual_ln: dim=%d (supported: 512, 1152, 3072)",
                a.dim);
  ECADK_REQUIRE(a.n_reuse >= 0 && a.n_reuse <= ECADK_MAX_REUSE, "residual_ln: n_reuse=%d", a.n_reuse);
  ECADK_REQUIRE(a.h == nullptr || (a.shift_temb && a.scale_temb),
                "residual_ln: LayerNorm path needs the per-sample shift/scale vectors (tables are optional)");
  ECADK_REQUIRE(a.n_reuse > 0 || a.h != nullptr || a.xb != nullptr, "residual_ln: nothing to do");
  ResidualLnParams p;
  p.x = a.x;
  p.xb = static_cast<__nv_bfloat16*>(a.xb);
  p.h = static_cast<__nv_bfloat16*>(a.h);
  p.M = a.rows;
  p.tokens = a.tokens;
  p.n_reuse = a.n_reuse;
  for (int i = 0; i < ECADK_MAX_REUSE; ++i) {
    p.reuse[i].cache = static_cast<const __nv_bfloat16*>(a.reuse[i].cache);
    p.reuse[i].gate_table = a.reuse[i].gate_table;
    p.reuse[i].gate_temb = a.reuse[i].gate_temb;
    if (i < a.n_reuse) {
      ECADK_REQUIRE(p.reuse[i].cache != nullptr, "residual_ln: reuse[%d].cache is null", i);
    }
  }
  p.shift_table = a.shift_table;
  p.scale_table = a.scale_table;
  p.shift_temb = a.shift_temb;
  p.scale_temb = a.scale_temb;
  p.temb_stride = a.temb_stride;
  p.eps = a.eps;
  ECADK_REQUIRE(a.tokens % kRlnRowsPerBlock == 0, "residual_ln: tokens=%d must be a multiple of %d", a.tokens,
                kRlnRowsPerBlock);
  // small problems (batch-1 latency configuration): 8 rows per block so that the grid still covers the SMs
  const bool small = (a.rows + kRlnRowsPerBlock - 1) / kRlnRowsPerBlock < 2 * num_sms();
  const int rpb = small ? 8 : kRlnRowsPerBlock;
  const int grid = (a.rows + rpb - 1) / rpb;
  const int smem = (2 + a.n_reuse) * a.dim * 4;
  const double elems = static_cast<double>(a.rows) * a.dim;
  ProfScope prof(ECADK_PROF_GLUE, 0.0,
                 elems * (4.0 + (a.n_reuse > 0 ? 4.0 : 0.0) + 2.0 * a.n_reuse + (a.xb ? 2.0 : 0.0) + (a.h ? 2.0 : 0.0)),
                 stream);
  auto launch = [&](auto kern, int max_smem, bool& configured) -> int {
    if (!configured && max_smem > 48 * 1024) {
      ECADK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      configured = true;
    }
    launch_pdl(kern, dim3(grid), dim3(256), smem, stream, p);
    return ECADK_OK;
  };
  int rc;
  if (a.dim == 512) {
    static bool c0 = false, c1 = false;
    rc = small ? launch(residual_ln_kernel<4, 8>, 0, c0) : launch(residual_ln_kernel<4>, 0, c1);
  } else if (a.dim == 1152) {  // up to (2 + 12) staged vectors = 63 KB > the 48 KB default
    static bool c0 = false, c1 = false;
    const int mx = (2 + ECADK_MAX_REUSE) * 1152 * 4;
    rc = small ? launch(residual_ln_kernel<9, 8>, mx, c0) : launch(residual_ln_kernel<9>, mx, c1);
  } else {
    static bool c0 = false, c1 = false;
    const int mx = (2 + ECADK_MAX_REUSE) * 3072 * 4;
    rc = small ? launch(residual_ln_kernel<24, 8>, mx, c0) : launch(residual_ln_kernel<24>, mx, c1);
  }
  if (rc) return rc;
  return check_launch("residual_ln_kernel");
}

}  // namespace

// =====================================================================================================
// exported C ABI
// =====================================================================================================
struct EcadkHandle_ {
  int device;
  EcadkModelDesc desc;
  std::vector<EcadkBlockWeights> blocks;
  void* split_ws = nullptr;  // split-K partial tiles of the small-batch GEMMs (gemm_splitk_kernel), kSplitWsBytes
};

struct EcadkFluxHandle_ {
  int device;
  EcadkFluxDesc desc;
  std::vector<EcadkFluxDoubleWeights> dbl;
  std::vector<EcadkFluxSingleWeights> sgl;
};

namespace {

// One residual stream of the FLUX executor with its lazily applied cached-residual reuses (same idea as the PixArt
// executor: a reused component only queues (cache, gate); the next kernel that reads the stream folds them in).
struct FluxStream {
  float* x;
  int rows, tokens, dim, mod_stride;
  float eps;
  cudaStream_t stream;
  int* launches;
  EcadkResidualLnArgs pend;

  void reset() {
    memset(&pend, 0, sizeof(pend));
    pend.x = x;
    pend.rows = rows;
    pend.tokens = tokens;
    pend.dim = dim;
    pend.temb_stride = mod_stride;
    pend.eps = eps;
  }
  int flush() {
    if (pend.n_reuse == 0 && pend.h == nullptr && pend.xb == nullptr) return ECADK_OK;
    const int rc = launch_residual_ln(pend, stream);
    ++*launches;
    reset();
    return rc;
  }
  int push(const void* cache, const float* gate_vec) {
    if (pend.n_reuse == ECADK_MAX_REUSE) {
      const int rc = flush();
      if (rc) return rc;
    }
    pend.reuse[pend.n_reuse++] = EcadkReuse{cache, nullptr, gate_vec};
    return ECADK_OK;
  }
  // h = LN(x + pending) * (1 + scale) + shift, modulation vectors per sample (no table)
  int layer_norm(void* h, const float* shift_vec, const float* scale_vec) {
    pend.h = h;
    pend.shift_temb = shift_vec;
    pend.scale_temb = scale_vec;
    return flush();
  }
};

}  // namespace

extern "C" {

int ecadk_abi_version(void) { return ECADK_ABI_VERSION; }

const char* ecadk_last_error(void) { return g_last_error.c_str(); }

int ecadk_device_check(int device) {
  cudaDeviceProp prop;
  ECADK_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    return fail(ECADK_EARCH, "device %d is sm_%d%d; libecad_b200 is built for sm_100a only", device, prop.major,
                prop.minor);
  }
  return ECADK_OK;
}

int ecadk_residual_ln(const EcadkResidualLnArgs* args, ecadk_stream_t stream) {
  ECADK_REQUIRE(args != nullptr, "residual_ln: null args");
  return launch_residual_ln(*args, static_cast<cudaStream_t>(stream));
}

int ecadk_patch_embed_padded(const float* latents, const float* wt, const float* bias, const float* pos, float* x,
                             int samples, int channels, int hl, int wl, int dim, int tokens_pad,
                             ecadk_stream_t stream_) {
  ECADK_REQUIRE(latents && wt && bias && pos && x, "patch_embed: null pointer");
  ECADK_REQUIRE(channels * 4 <= 64 && hl % 2 == 0 && wl % 2 == 0 && dim % 4 == 0, "patch_embed: bad shape");
  const int n_real = (hl / 2) * (wl / 2);
  ECADK_REQUIRE(tokens_pad >= n_real, "patch_embed: tokens_pad=%d < %d tokens", tokens_pad, n_real);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, stream);
  PatchEmbedParams p{latents, wt, bias, pos, x, samples, channels, hl, wl, dim, tokens_pad};
  const int tokens = samples * n_real;
  patch_embed_kernel<<<(tokens + kPatchTokensPerBlock - 1) / kPatchTokensPerBlock, 288, 0, stream>>>(p);
  int rc = check_launch("patch_embed_kernel");
  if (rc == ECADK_OK && tokens_pad > n_real) {
    // padding rows of every sample start at zero: LayerNorm of a zero row is zero, every later value stays finite,
    // and the self-attention masks them as keys (EcadkBlocksArgs.self_bias)
    const size_t row = static_cast<size_t>(dim) * sizeof(float);
    ECADK_CHECK_CUDA(cudaMemset2DAsync(x + static_cast<size_t>(n_real) * dim, tokens_pad * row, 0,
                                       (tokens_pad - n_real) * row, samples, stream));
  }
  return rc;
}

int ecadk_patch_embed(const float* latents, const float* wt, const float* bias, const float* pos, float* x,
                      int samples, int channels, int hl, int wl, int dim, ecadk_stream_t stream) {
  return ecadk_patch_embed_padded(latents, wt, bias, pos, x, samples, channels, hl, wl, dim, (hl / 2) * (wl / 2), stream);
}

int ecadk_timestep_sinusoid(const float* t, float* out, int samples, int dim, ecadk_stream_t stream) {
  ECADK_REQUIRE(t && out && samples > 0 && dim > 0 && dim % 2 == 0, "timestep_sinusoid: bad args");
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, static_cast<cudaStream_t>(stream));
  const int n = samples * (dim / 2);
  timestep_sinusoid_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(t, out, samples, dim);
  return check_launch("timestep_sinusoid_kernel");
}

int ecadk_small_linear(const float* x, int ldx, const float* w, const float* b, float* y, int samples, int k, int o,
                       int ldy, int y_off, int act_in, int accumulate, ecadk_stream_t stream) {
  ECADK_REQUIRE(x && w && b && y && samples > 0 && k > 0 && o > 0, "small_linear: bad args");
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, static_cast<cudaStream_t>(stream));
  ECADK_REQUIRE(ldx >= k, "small_linear: ldx=%d < k=%d", ldx, k);
  SmallLinearParams p{x, w, b, y, samples, k, o, ldy, y_off, act_in, accumulate, ldx};
  small_linear_kernel<<<dim3((o + 7) / 8, (samples + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("small_linear_kernel");
}

int ecadk_cast_f32_bf16(const float* in, void* out, size_t n, ecadk_stream_t stream) {
  ECADK_REQUIRE(in && out && n % 4 == 0 && aligned16(in), "cast_f32_bf16: n must be a multiple of 4, 16B aligned");
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, static_cast<cudaStream_t>(stream));
  const size_t n4 = n / 4;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > static_cast<size_t>(num_sms()) * 16) blocks = static_cast<size_t>(num_sms()) * 16;
  if (blocks == 0) return ECADK_OK;
  cast_to_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, static_cast<__nv_bfloat16*>(out), n4);
  return check_launch("cast_to_bf16_kernel");
}

int ecadk_silu_f32_bf16(const float* in, void* out, size_t n, ecadk_stream_t stream) {
  ECADK_REQUIRE(in && out && n % 4 == 0 && aligned16(in), "silu_f32_bf16: n must be a multiple of 4, 16B aligned");
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, static_cast<cudaStream_t>(stream));
  const size_t n4 = n / 4;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > static_cast<size_t>(num_sms()) * 16) blocks = static_cast<size_t>(num_sms()) * 16;
  if (blocks == 0) return ECADK_OK;
  silu_to_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, static_cast<__nv_bfloat16*>(out), n4);
  return check_launch("silu_to_bf16_kernel");
}

int ecadk_average_halves(void* buf, size_t half_elems, ecadk_stream_t stream) {
  ECADK_REQUIRE(buf && half_elems % 8 == 0 && aligned16(buf), "average_halves: need 16B alignment, multiple of 8");
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, static_cast<cudaStream_t>(stream));
  const size_t n8 = half_elems / 8;
  size_t blocks = (n8 + 255) / 256;
  if (blocks > static_cast<size_t>(num_sms()) * 16) blocks = static_cast<size_t>(num_sms()) * 16;
  if (blocks == 0) return ECADK_OK;
  average_halves_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(buf), n8);
  return check_launch("average_halves_kernel");
}

int ecadk_mask_bias(const float* mask, float* bias, int samples, int t, int t_pad, ecadk_stream_t stream) {
  ECADK_REQUIRE(mask && bias && samples > 0 && t > 0 && t_pad >= t, "mask_bias: bad args");
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, static_cast<cudaStream_t>(stream));
  const int n = samples * t_pad;
  mask_bias_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, bias, samples, t, t_pad);
  return check_launch("mask_bias_kernel");
}

int ecadk_final_layer(const float* x, const float* table, const float* emb, int emb_stride, const void* w_pad,
                      const float* bias, void* h_scratch, float* out, int samples, int hp, int wp, int dim,
                      int out_channels, float eps, ecadk_stream_t stream) {
  return ecadk_final_layer_padded(x, table, emb, emb_stride, w_pad, bias, h_scratch, out, samples, hp, wp, hp * wp, dim,
                                  out_channels, eps, stream);
}

int ecadk_final_layer_padded(const float* x, const float* table, const float* emb, int emb_stride, const void* w_pad,
                             const float* bias, void* h_scratch, float* out, int samples, int hp, int wp,
                             int tokens_pad, int dim, int out_channels, float eps, ecadk_stream_t stream) {
  ECADK_REQUIRE(x && table && emb && w_pad && bias && h_scratch && out, "final_layer: null pointer");
  ECADK_REQUIRE(4 * out_channels <= 128 && out_channels % 4 == 0, "final_layer: out_channels=%d", out_channels);
  ECADK_REQUIRE(tokens_pad >= hp * wp, "final_layer: tokens_pad=%d < %d tokens", tokens_pad, hp * wp);
  // rows [hp*wp, tokens_pad) of every sample are padding: normalised like the rest, dropped by the unpatchify epilogue
  const int tokens = tokens_pad, M = samples * tokens;
  // 1. h = LN(x) * (1 + scale) + shift with shift = table[0] + emb[s], scale = table[1] + emb[s]
  EcadkResidualLnArgs a;
  memset(&a, 0, sizeof(a));
  a.x = const_cast<float*>(x);
  a.h = h_scratch;
  a.rows = M;
  a.tokens = tokens;
  a.dim = dim;
  a.shift_table = table;
  a.scale_table = table + dim;
  a.shift_temb = emb;
  a.scale_temb = emb;
  a.temb_stride = emb_stride;
  a.eps = eps;
  int rc = launch_residual_ln(a, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  // 2. [M, dim] x [128, dim]^T on the tensor cores; the epilogue keeps the first p*p*C columns and unpatchifies
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = 128; p.K = dim;
  p.bias = bias;
  p.tokens = tokens;
  p.unp_out = out;
  p.unp_wp = wp;
  p.unp_hp = hp;
  p.unp_c = out_channels;
  p.unp_cols = 4 * out_channels;
  return launch_gemm<EPI_UNPATCHIFY>(h_scratch, w_pad, p, static_cast<cudaStream_t>(stream));
}

int ecadk_cfg_dpm_step(const float* noise, float* latents, float* x0_prev, int batch, int channels, int hw,
                       int has_cfg, float guidance, float sigma_s, float alpha_s, float c_x, float c_d0, float c_d1,
                       ecadk_stream_t stream) {
  ECADK_REQUIRE(noise && latents && x0_prev && batch > 0 && channels > 0 && hw > 0, "cfg_dpm_step: bad args");
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, static_cast<cudaStream_t>(stream));
  CfgDpmParams p{noise, latents, x0_prev, batch, channels, hw, has_cfg, guidance, sigma_s, alpha_s, c_x, c_d0, c_d1};
  const int n = batch * channels * hw;
  cfg_dpm_step_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("cfg_dpm_step_kernel");
}

int ecadk_gemm_bias(const void* a, const void* w, const float* bias, void* out, int m, int n, int k, int ldo,
                    int gelu, ecadk_stream_t stream) {
  ECADK_REQUIRE(out != nullptr && aligned16(out) && ldo % 8 == 0 && ldo >= n, "gemm_bias: bad output");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = m; p.N = n; p.K = k;
  p.bias = bias;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.tokens = 1;
  return gelu ? launch_gemm<EPI_BIAS_GELU>(a, w, p, static_cast<cudaStream_t>(stream))
              : launch_gemm<EPI_BIAS>(a, w, p, static_cast<cudaStream_t>(stream));
}

int ecadk_gemm_bias_gated_residual_cache(const void* a, const void* w, const float* bias, float* x, void* xb,
                                         void* cache, const float* gate_table, const float* gate_temb,
                                         int temb_stride, int tokens, int m, int n, int k, ecadk_stream_t stream) {
  ECADK_REQUIRE(x && aligned16(x) && (cache == nullptr || aligned16(cache)) && (xb == nullptr || aligned16(xb)),
                "gemm_gated_residual: bad x/cache/xb");
  ECADK_REQUIRE(tokens > 0 && tokens % 32 == 0, "gemm_gated_residual: tokens=%d must be a multiple of 32", tokens);
  ECADK_REQUIRE(gate_table == nullptr || gate_temb != nullptr || temb_stride == 0,
                "gemm_gated_residual: a gate table without a per-sample vector needs temb_stride = 0");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = m; p.N = n; p.K = k;
  p.bias = bias;
  p.x = x;
  p.xb = static_cast<__nv_bfloat16*>(xb);
  p.cache = static_cast<__nv_bfloat16*>(cache);
  p.gate_table = gate_table;
  p.gate_temb = gate_temb;
  p.temb_stride = temb_stride;
  p.tokens = tokens;
  return launch_gemm<EPI_GATED_RESIDUAL>(a, w, p, static_cast<cudaStream_t>(stream));
}

int ecadk_gemm2src_gated_residual_cache(const void* a, int k1, const void* a2, const void* w, const float* bias,
                                        float* x, void* cache, const float* gate_temb, int temb_stride, int tokens,
                                        int m, int n, int k, ecadk_stream_t stream) {
  ECADK_REQUIRE(x && aligned16(x) && (cache == nullptr || aligned16(cache)) && a2 != nullptr,
                "gemm2src_gated_residual: bad x/cache/a2");
  ECADK_REQUIRE(tokens > 0 && tokens % 32 == 0, "gemm2src_gated_residual: tokens=%d must be a multiple of 32", tokens);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = m; p.N = n; p.K = k;
  p.a2 = a2;
  p.k1 = k1;
  p.bias = bias;
  p.x = x;
  p.cache = static_cast<__nv_bfloat16*>(cache);
  p.gate_temb = gate_temb;
  p.temb_stride = temb_stride;
  p.tokens = tokens;
  return launch_gemm<EPI_GATED_RESIDUAL>(a, w, p, static_cast<cudaStream_t>(stream));
}

int ecadk_gemm_bias_dual(const void* a, const void* w, const float* bias, void* out_pre, void* out_gelu, int m, int n,
                         int k, int ldo_pre, int ldo_gelu, ecadk_stream_t stream) {
  ECADK_REQUIRE(out_gelu != nullptr && aligned16(out_gelu) && (out_pre == nullptr || aligned16(out_pre)),
                "gemm_bias_dual: bad outputs");
  ECADK_REQUIRE(ldo_gelu >= n && ldo_gelu % 8 == 0 && (out_pre == nullptr || (ldo_pre >= n && ldo_pre % 8 == 0)),
                "gemm_bias_dual: row pitches must be >= n and multiples of 8");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = m; p.N = n; p.K = k;
  p.bias = bias;
  p.out = static_cast<__nv_bfloat16*>(out_pre);
  p.ldo = ldo_pre;
  p.out2 = static_cast<__nv_bfloat16*>(out_gelu);
  p.ldo2 = ldo_gelu;
  return launch_gemm<EPI_BIAS_DUAL>(a, w, p, static_cast<cudaStream_t>(stream));
}

int ecadk_gemm_bias_headmajor(const void* a, const void* w, const float* bias, void* out0, void* out1, void* out2,
                              int n_parts, int heads, int tokens, int tokens_pad, int m, int k,
                              ecadk_stream_t stream) {
  ECADK_REQUIRE(n_parts >= 1 && n_parts <= 3 && out0, "gemm_headmajor: n_parts=%d", n_parts);
  ECADK_REQUIRE(tokens > 0 && tokens_pad >= tokens && heads > 0, "gemm_headmajor: bad token/head counts");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = m; p.N = n_parts * heads * kHeadDim; p.K = k;
  p.bias = bias;
  p.hm_out[0] = static_cast<__nv_bfloat16*>(out0);
  p.hm_out[1] = static_cast<__nv_bfloat16*>(out1);
  p.hm_out[2] = static_cast<__nv_bfloat16*>(out2);
  for (int i = 0; i < n_parts; ++i) ECADK_REQUIRE(p.hm_out[i] && aligned16(p.hm_out[i]), "gemm_headmajor: out%d", i);
  p.heads = heads;
  p.head_dim = kHeadDim;
  p.head_pad = kHeadPad;
  p.tokens = tokens;
  p.tokens_pad = tokens_pad;
  return launch_gemm<EPI_HEADMAJOR>(a, w, p, static_cast<cudaStream_t>(stream));
}

int ecadk_attention_ex(const void* q, int q_ld, const void* k, const void* v, int kv_ld, const float* bias, void* out,
                       int samples, int heads, int q_tokens, int n_keys, ecadk_stream_t stream) {
  return launch_attention_ex(q, q_ld, k, v, kv_ld, bias, out, samples, heads, q_tokens, n_keys,
                             static_cast<cudaStream_t>(stream));
}

int ecadk_attention(const void* q, const void* k, const void* v, const float* bias, void* out, int samples,
                    int heads, int q_tokens, int n_keys, ecadk_stream_t stream) {
  return launch_attention(q, k, v, bias, out, samples, heads, q_tokens, n_keys, static_cast<cudaStream_t>(stream));
}

int ecadk_flux_create(int device, const EcadkFluxDesc* desc, const EcadkFluxDoubleWeights* dbl,
                      const EcadkFluxSingleWeights* sgl, ecadk_flux_handle_t* out) {
  ECADK_REQUIRE(desc && dbl && sgl && out, "flux_create: null argument");
  ECADK_REQUIRE(desc->dim == desc->heads * 128, "flux_create: dim must be heads*128");
  ECADK_REQUIRE(desc->dim == 512 || desc->dim == 3072, "flux_create: dim=%d (supported: 3072, 512 for tests)", desc->dim);
  int rc = ecadk_device_check(device);
  if (rc) return rc;
  auto* h = new EcadkFluxHandle_();
  h->device = device;
  h->desc = *desc;
  h->dbl.assign(dbl, dbl + desc->num_layers);
  h->sgl.assign(sgl, sgl + desc->num_single_layers);
  *out = h;
  return ECADK_OK;
}

int ecadk_flux_destroy(ecadk_flux_handle_t h) {
  delete h;
  return ECADK_OK;
}

int ecadk_flux_blocks(ecadk_flux_handle_t h, const EcadkFluxArgs* a, const uint8_t* executed, int* n_launches,
                      ecadk_stream_t stream_) {
  ECADK_REQUIRE(h && a && executed, "flux_blocks: null argument");
  const EcadkFluxDesc& d = h->desc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int D = d.dim, H = d.heads, B = a->samples, N = a->img_tokens, T = a->txt_tokens, S = T + N, F = 4 * D;
  ECADK_REQUIRE(N % 32 == 0 && T % 32 == 0 && S % 256 == 0, "flux_blocks: tokens N=%d T=%d (N, T %% 32, N+T %% 256)", N, T);
  int launches = 0;
  int rc;
  const int MS = a->mod_stride;
  ECADK_REQUIRE(a->mod != nullptr && MS >= (d.num_layers * 12 + d.num_single_layers * 3) * D,
                "flux_blocks: mod_stride=%d too small", MS);
  NvtxRange nv_call("ecadk_flux_blocks B=%d N=%d T=%d", B, N, T);
  FluxStream img{a->x_img, B * N, N, D, MS, d.eps, stream, &launches, {}};
  FluxStream txt{a->x_txt, B * T, T, D, MS, d.eps, stream, &launches, {}};
  img.reset();
  txt.reset();

  // ------------------------------------------------------------------ double-stream blocks
  for (int b = 0; b < d.num_layers; ++b) {
    const EcadkFluxDoubleWeights& w = h->dbl[b];
    const float* mi = a->mod + static_cast<size_t>(b) * 12 * D;  // image stream modulation (row pitch MS)
    const float* mt = mi + 6 * D;                                  // text stream
    void* c_attn = a->cache_double[b * 4 + 0];
    void* c_ctx = a->cache_double[b * 4 + 1];
    void* c_ff = a->cache_double[b * 4 + 2];
    void* c_ffc = a->cache_double[b * 4 + 3];
    const bool ex_attn = executed[b * 3 + 0], ex_ff = executed[b * 3 + 1], ex_ffc = executed[b * 3 + 2];
    const uint8_t* dead = a->cache_dead;  // dead-store elimination (see EcadkFluxArgs.cache_dead)
    void* s_attn = (dead && dead[b * 3 + 0]) ? nullptr : c_attn;
    void* s_ctx = (dead && dead[b * 3 + 0]) ? nullptr : c_ctx;
    void* s_ff = (dead && dead[b * 3 + 1]) ? nullptr : c_ff;
    void* s_ffc = (dead && dead[b * 3 + 2]) ? nullptr : c_ffc;
    // joint attention (cached_flux_transformer_block.py:170-201,247-256): chunks shift_msa 0, scale_msa 1, gate_msa 2
    if (ex_attn) {
      NvtxRange nv("d%02d.attn", b);
      if ((rc = img.layer_norm(a->h_img, mi + 0 * D, mi + 1 * D))) return rc;
      if ((rc = txt.layer_norm(a->h_txt, mt + 0 * D, mt + 1 * D))) return rc;
      if ((rc = ecadk_gemm_bias_headmajor_ex(a->h_txt, w.w_qkv_ctx, w.b_qkv_ctx, a->q, a->k, a->v, 3, H, 128, 128, T, S,
                                             0, B * T, D, stream_)))
        return rc;
      if ((rc = ecadk_gemm_bias_headmajor_ex(a->h_img, w.w_qkv, w.b_qkv, a->q, a->k, a->v, 3, H, 128, 128, N, S, T,
                                             B * N, D, stream_)))
        return rc;
      if ((rc = ecadk_qk_norm_rope_batched(a->q, a->k, w.norm_q, w.norm_k, w.norm_added_q, w.norm_added_k, a->rope_cos,
                                           a->rope_sin, a->rope_sample_stride, B, H, S, T, d.eps, stream_)))
        return rc;
      if ((rc = launch_attention_d128(a->q, a->k, a->v, a->attn_img, D, a->attn_txt, T, B, H, S, S, stream))) return rc;
      if ((rc = ecadk_gemm_bias_gated_residual_cache(a->attn_img, w.w_out, w.b_out, a->x_img, nullptr, s_attn, nullptr,
                                                     mi + 2 * D, MS, N, B * N, D, D, stream_)))
        return rc;
      if ((rc = ecadk_gemm_bias_gated_residual_cache(a->attn_txt, w.w_out_ctx, w.b_out_ctx, a->x_txt, nullptr, s_ctx,
                                                     nullptr, mt + 2 * D, MS, T, B * T, D, D, stream_)))
        return rc;
      launches += 6;
    } else {
      if ((rc = img.push(c_attn, mi + 2 * D))) return rc;
      if ((rc = txt.push(c_ctx, mt + 2 * D))) return rc;
    }
    // image feed-forward (:258-268): shift_mlp 3, scale_mlp 4, gate_mlp 5
    if (ex_ff) {
      NvtxRange nv("d%02d.ff", b);
      if ((rc = img.layer_norm(a->h_img, mi + 3 * D, mi + 4 * D))) return rc;
      if ((rc = ecadk_gemm_bias(a->h_img, w.w_ff1, w.b_ff1, a->ffh, B * N, F, D, F, 1, stream_))) return rc;
      if ((rc = ecadk_gemm_bias_gated_residual_cache(a->ffh, w.w_ff2, w.b_ff2, a->x_img, nullptr, s_ff, nullptr,
                                                     mi + 5 * D, MS, N, B * N, D, F, stream_)))
        return rc;
      launches += 2;
    } else {
      if ((rc = img.push(c_ff, mi + 5 * D))) return rc;
    }
    // text feed-forward (:272-287)
    if (ex_ffc) {
      NvtxRange nv("d%02d.ff_context", b);
      if ((rc = txt.layer_norm(a->h_txt, mt + 3 * D, mt + 4 * D))) return rc;
      if ((rc = ecadk_gemm_bias(a->h_txt, w.w_ff1_ctx, w.b_ff1_ctx, a->ffh, B * T, F, D, F, 1, stream_))) return rc;
      if ((rc = ecadk_gemm_bias_gated_residual_cache(a->ffh, w.w_ff2_ctx, w.b_ff2_ctx, a->x_txt, nullptr, s_ffc,
                                                     nullptr, mt + 5 * D, MS, T, B * T, D, F, stream_)))
        return rc;
      launches += 2;
    } else {
      if ((rc = txt.push(c_ffc, mt + 5 * D))) return rc;
    }
  }
  if ((rc = img.flush())) return rc;
  if ((rc = txt.flush())) return rc;

  // ------------------------------------------------------------------ concat [text; image] (:205-207)
  const size_t row_bytes = static_cast<size_t>(D) * sizeof(float);
  ECADK_CHECK_CUDA(cudaMemcpy2DAsync(a->x_cat, S * row_bytes, a->x_txt, T * row_bytes, T * row_bytes, B,
                                     cudaMemcpyDeviceToDevice, stream));
  ECADK_CHECK_CUDA(cudaMemcpy2DAsync(a->x_cat + static_cast<size_t>(T) * D, S * row_bytes, a->x_img, N * row_bytes,
                                     N * row_bytes, B, cudaMemcpyDeviceToDevice, stream));

  // ------------------------------------------------------------------ single-stream blocks (:99-130)
  FluxStream cat{a->x_cat, B * S, S, D, MS, d.eps, stream, &launches, {}};
  cat.reset();
  for (int b = 0; b < d.num_single_layers; ++b) {
    const EcadkFluxSingleWeights& w = h->sgl[b];
    const float* ms = a->mod + (static_cast<size_t>(d.num_layers) * 12 + static_cast<size_t>(b) * 3) * D;  // shift|scale|gate
    void* c_attn = a->cache_single[b * 3 + 0];
    void* c_mlp = a->cache_single[b * 3 + 1];
    void* c_out = a->cache_single[b * 3 + 2];
    const int r = (d.num_layers + b) * 3;
    const bool ex_attn = executed[r + 0], ex_mlp = executed[r + 1], ex_out = executed[r + 2];
    if (ex_attn || ex_mlp) {
      if ((rc = cat.layer_norm(a->h_cat, ms + 0 * D, ms + 1 * D))) return rc;
    }
    const uint8_t* dead = a->cache_dead;
    if (ex_mlp) {
      NvtxRange nv("s%02d.proj_mlp", b);
      // proj_mlp: ONE GEMM writes the pre-activation into its cache slot (what the reference caches; skipped when
      // that store is dead) and GELU(tanh) of it into the operand buffer proj_out reads
      void* s_mlp = (dead && dead[r + 1]) ? nullptr : c_mlp;
      if ((rc = ecadk_gemm_bias_dual(a->h_cat, w.w_mlp, w.b_mlp, s_mlp, a->cat, B * S, F, D, F, F, stream_))) return rc;
      ++launches;
    }
    if (ex_attn) {
      NvtxRange nv("s%02d.attn", b);
      if ((rc = ecadk_gemm_bias_headmajor_ex(a->h_cat, w.w_qkv, w.b_qkv, a->q, a->k, a->v, 3, H, 128, 128, S, S, 0,
                                             B * S, D, stream_)))
        return rc;
      if ((rc = ecadk_qk_norm_rope_batched(a->q, a->k, w.norm_q, w.norm_k, nullptr, nullptr, a->rope_cos, a->rope_sin,
                                           a->rope_sample_stride, B, H, S, 0, d.eps, stream_)))
        return rc;
      if ((rc = launch_attention_d128(a->q, a->k, a->v, c_attn, D, nullptr, 0, B, H, S, S, stream))) return rc;
      launches += 3;
    }
    if (ex_out) {
      NvtxRange nv("s%02d.proj_out", b);
      // proj_out over [attn | GELU(proj_mlp)]: the A operand is read from the two buffers directly (split along K), no
      // concatenated copy.  A reused proj_mlp is re-activated from its pre-GELU cache.
      if (!ex_mlp) {
        if ((rc = ecadk_strided_unary(c_mlp, a->cat, B * S, F, F, F, 1, stream_))) return rc;
        ++launches;
      }
      if ((rc = cat.flush())) return rc;  // the epilogue below updates x in place: pending reuses must land first
      void* s_out = (dead && dead[r + 2]) ? nullptr : c_out;
      if ((rc = ecadk_gemm2src_gated_residual_cache(c_attn, D, a->cat, w.w_out, w.b_out, a->x_cat, s_out, ms + 2 * D, MS,
                                                    S, B * S, D, 5 * D, stream_)))
        return rc;
      ++launches;
    } else {
      if ((rc = cat.push(c_out, ms + 2 * D))) return rc;
    }
  }
  if ((rc = cat.flush())) return rc;
  // image rows of the concatenated stream -> x_img (the slice at flux_transformer_2d_edited.py:314)
  ECADK_CHECK_CUDA(cudaMemcpy2DAsync(a->x_img, N * row_bytes, a->x_cat + static_cast<size_t>(T) * D, S * row_bytes,
                                     N * row_bytes, B, cudaMemcpyDeviceToDevice, stream));
  if (n_launches) *n_launches = launches;
  return ECADK_OK;
}

#ifdef ECADK_ATTN_TIMING
// instrumented builds only: how fast can TMA stream head-major [rows, 80] bf16 tensors the way the attention kernels
// read them (64-column + 16-column boxes, or five 16-column boxes) versus fully contiguous boxes of the same bytes?
__global__ void __launch_bounds__(64, 1) tma_stream_kernel(const __grid_constant__ CUtensorMap t64,
                                                            const __grid_constant__ CUtensorMap t16,
                                                            const __grid_constant__ CUtensorMap tflat, int tensors,
                                                            int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kStages = 4, kStage = 256 * 160;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStage);
  uint64_t* empty = full + kStages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int t = blockIdx.x; t < tensors; t += gridDim.x, ++n) {
      const int st = n % kStages;
      mbar_wait(&empty[st], ((n / kStages) & 1) ^ 1);
      mbar_arrive_expect_tx(&full[st], kStage);
      uint8_t* d = smem + st * kStage;
      const int row = t * 256;
      if (mode == 0) {  // 64 + 16 columns (Q, K)
        tma_load_2d(d, &t64, &full[st], 0, row);
        tma_load_2d(d + 256 * 128, &t16, &full[st], 64, row);
      } else if (mode == 1) {  // five 16-column atoms (V)
        for (int a = 0; a < 5; ++a) tma_load_2d(d + a * 256 * 32, &t16, &full[st], a * 16, row);
      } else {  // the same 40 KB as one contiguous block: 320 rows of 128 B
        tma_load_2d(d, &tflat, &full[st], 0, t * 320);
        tma_load_2d(d + 160 * 128, &tflat, &full[st], 0, t * 320 + 160);
      }
    }
  } else if (threadIdx.x == 32) {
    int n = 0;
    for (int t = blockIdx.x; t < tensors; t += gridDim.x, ++n) {
      const int st = n % kStages;
      mbar_wait(&full[st], (n / kStages) & 1);
      mbar_arrive(&empty[st]);
    }
  }
}

int ecadk_debug_tma_stream(const void* buf, int tensors, int mode, float* ms_out) {
  CUtensorMap t64, t16, tflat;
  int rc;
  const uint64_t rows = static_cast<uint64_t>(tensors) * 256;
  if ((rc = make_tmap_bf16(&t64, buf, rows, 80, 80, 256, 64, 128))) return rc;
  if ((rc = make_tmap_bf16(&t16, buf, rows, 80, 80, 256, 16, 32))) return rc;
  if ((rc = make_tmap_bf16(&tflat, buf, static_cast<uint64_t>(tensors) * 320, 64, 64, 160, 64, 128))) return rc;
  const int smem = 4 * 256 * 160 + 1024 + 256;
  ECADK_CHECK_CUDA(cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b;
  ECADK_CHECK_CUDA(cudaEventCreate(&a));
  ECADK_CHECK_CUDA(cudaEventCreate(&b));
  tma_stream_kernel<<<num_sms(), 64, smem>>>(t64, t16, tflat, tensors, mode);
  ECADK_CHECK_CUDA(cudaEventRecord(a));
  for (int i = 0; i < 5; ++i) tma_stream_kernel<<<num_sms(), 64, smem>>>(t64, t16, tflat, tensors, mode);
  ECADK_CHECK_CUDA(cudaEventRecord(b));
  ECADK_CHECK_CUDA(cudaEventSynchronize(b));
  ECADK_CHECK_CUDA(cudaEventElapsedTime(ms_out, a, b));
  *ms_out /= 5;
  return check_launch("tma_stream_kernel");
}

// instrumented builds only: copies the phase-clock accumulators of the last attn_flash_kernel launch to the host
int ecadk_debug_attn_timing(unsigned int* out_host) {
  ECADK_CHECK_CUDA(cudaDeviceSynchronize());
  ECADK_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_attn_dbg, sizeof(unsigned int) * 148 * 32));
  return ECADK_OK;
}
#endif

int ecadk_set_nvtx(int on) {
  g_nvtx.store(on ? 1 : 0, std::memory_order_relaxed);
  return ECADK_OK;
}

long long ecadk_nvtx_ranges(void) { return g_nvtx_ranges.load(std::memory_order_relaxed); }

int ecadk_profile_start(void) {
  g_prof.used = 0;
  g_prof.recs.clear();
  g_prof.on = true;
  return ECADK_OK;
}

int ecadk_profile_stop(EcadkProfileRecord* out) {
  ECADK_REQUIRE(out != nullptr, "profile_stop: null output");
  g_prof.on = false;
  ECADK_CHECK_CUDA(cudaDeviceSynchronize());
  for (int c = 0; c < ECADK_PROF_CLASSES; ++c) out[c] = EcadkProfileRecord{0, 0.0, 0.0, 0.0};
  for (const ProfRec& r : g_prof.recs) {
    float ms = 0.f;
    ECADK_CHECK_CUDA(cudaEventElapsedTime(&ms, g_prof.events[r.ev], g_prof.events[r.ev + 1]));
    EcadkProfileRecord& o = out[r.cls];
    o.launches += 1;
    o.total_ms += ms;
    o.flops += r.flops;
    o.bytes += r.bytes;
  }
  g_prof.recs.clear();
  g_prof.used = 0;
  return ECADK_OK;
}

int ecadk_attention_d128(const void* q, const void* k, const void* v, void* out, int out_ld, void* out_lo,
                         int split_tokens, int samples, int heads, int q_tokens, int n_keys, ecadk_stream_t stream) {
  return launch_attention_d128(q, k, v, out, out_ld, out_lo, split_tokens, samples, heads, q_tokens, n_keys,
                               static_cast<cudaStream_t>(stream));
}

int ecadk_qk_norm_rope(void* q, void* k, const float* wq, const float* wk, const float* wq_add, const float* wk_add,
                       const float* rope_cos, const float* rope_sin, int samples, int heads, int seq, int split,
                       float eps, ecadk_stream_t stream) {
  return ecadk_qk_norm_rope_batched(q, k, wq, wk, wq_add, wk_add, rope_cos, rope_sin, 0, samples, heads, seq, split, eps,
                                    stream);
}

int ecadk_qk_norm_rope_batched(void* q, void* k, const float* wq, const float* wk, const float* wq_add,
                               const float* wk_add, const float* rope_cos, const float* rope_sin, int rope_sample_stride,
                               int samples, int heads, int seq, int split, float eps, ecadk_stream_t stream) {
  ECADK_REQUIRE(q && k && wq && wk && rope_cos && rope_sin, "qk_norm_rope: null pointer");
  ECADK_REQUIRE(split == 0 || (wq_add && wk_add), "qk_norm_rope: split > 0 needs the added-stream norm weights");
  ECADK_REQUIRE(rope_sample_stride == 0 || rope_sample_stride >= seq * 64,
                "qk_norm_rope: rope_sample_stride=%d must be 0 (shared table) or >= seq*64", rope_sample_stride);
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, static_cast<cudaStream_t>(stream));
  QkNormRopeParams p{static_cast<__nv_bfloat16*>(q), static_cast<__nv_bfloat16*>(k), wq, wk, wq_add, wk_add,
                     rope_cos, rope_sin, samples * heads * seq, seq, split, eps, heads * seq, rope_sample_stride};
  qk_norm_rope_kernel<<<(p.rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("qk_norm_rope_kernel");
}

int ecadk_strided_unary(const void* src, void* dst, int rows, int cols, int ld_src, int ld_dst, int op,
                        ecadk_stream_t stream) {
  ECADK_REQUIRE(src && dst && rows > 0 && cols > 0 && cols % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0,
                "strided_unary: cols / pitches must be multiples of 8");
  ECADK_REQUIRE(aligned16(src) && aligned16(dst) && (op == 0 || op == 1), "strided_unary: alignment / op");
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, static_cast<cudaStream_t>(stream));
  const size_t total = static_cast<size_t>(rows) * (cols / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > static_cast<size_t>(num_sms()) * 32) blocks = static_cast<size_t>(num_sms()) * 32;
  strided_unary_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), rows, cols / 8, ld_src, ld_dst, op);
  return check_launch("strided_unary_kernel");
}

int ecadk_axpy_f32(float* y, const float* x, float a, size_t n, ecadk_stream_t stream) {
  ECADK_REQUIRE(y && x, "axpy: null pointer");
  if (n == 0) return ECADK_OK;
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, static_cast<cudaStream_t>(stream));
  size_t blocks = (n + 255) / 256;
  if (blocks > static_cast<size_t>(num_sms()) * 16) blocks = static_cast<size_t>(num_sms()) * 16;
  axpy_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, x, a, n);
  return check_launch("axpy_kernel");
}

int ecadk_gemm_bias_f32(const void* a, const void* w, const float* bias, float* out, int m, int n, int k, int ldo,
                        int out_cols, ecadk_stream_t stream) {
  ECADK_REQUIRE(out != nullptr && aligned16(out) && ldo % 4 == 0 && out_cols > 0 && out_cols <= n && out_cols % 4 == 0 &&
                    ldo >= out_cols,
                "gemm_bias_f32: bad output");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = m; p.N = n; p.K = k;
  p.bias = bias;
  p.f32_out = out;
  p.f32_cols = out_cols;
  p.ldo = ldo;
  p.tokens = 1;
  return launch_gemm<EPI_BIAS_F32>(a, w, p, static_cast<cudaStream_t>(stream));
}

int ecadk_gemm_bias_headmajor_ex(const void* a, const void* w, const float* bias, void* out0, void* out1, void* out2,
                                 int n_parts, int heads, int head_dim, int head_pad, int tokens, int tokens_pad,
                                 int tok_offset, int m, int k, ecadk_stream_t stream) {
  ECADK_REQUIRE(n_parts >= 1 && n_parts <= 3 && out0, "gemm_headmajor_ex: n_parts=%d", n_parts);
  ECADK_REQUIRE(tokens > 0 && tok_offset >= 0 && tokens_pad >= tokens + tok_offset && heads > 0,
                "gemm_headmajor_ex: bad token counts");
  ECADK_REQUIRE(head_dim % 4 == 0 && head_pad >= head_dim && head_pad % 8 == 0, "gemm_headmajor_ex: bad head dims");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = m; p.N = n_parts * heads * head_dim; p.K = k;
  p.bias = bias;
  p.hm_out[0] = static_cast<__nv_bfloat16*>(out0);
  p.hm_out[1] = static_cast<__nv_bfloat16*>(out1);
  p.hm_out[2] = static_cast<__nv_bfloat16*>(out2);
  for (int i = 0; i < n_parts; ++i) ECADK_REQUIRE(p.hm_out[i] && aligned16(p.hm_out[i]), "gemm_headmajor_ex: out%d", i);
  p.heads = heads;
  p.head_dim = head_dim;
  p.head_pad = head_pad;
  p.tokens = tokens;
  p.tokens_pad = tokens_pad;
  p.hm_tok_off = tok_offset;
  return launch_gemm<EPI_HEADMAJOR>(a, w, p, static_cast<cudaStream_t>(stream));
}

int ecadk_create(int device, const EcadkModelDesc* desc, const EcadkBlockWeights* blocks, ecadk_handle_t* out) {
  ECADK_REQUIRE(desc && blocks && out, "create: null argument");
  ECADK_REQUIRE(desc->num_layers > 0 && desc->dim == desc->heads * kHeadDim, "create: dim must be heads*72");
  int rc = ecadk_device_check(device);
  if (rc) return rc;
  auto* h = new EcadkHandle_();
  h->device = device;
  h->desc = *desc;
  h->blocks.assign(blocks, blocks + desc->num_layers);
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess && cur == device) {
    // optional: without it the small-batch GEMMs simply do not split
    if (cudaMalloc(&h->split_ws, kSplitWsBytes) != cudaSuccess) {
      cudaGetLastError();
      h->split_ws = nullptr;
    }
  }
  *out = h;
  return ECADK_OK;
}

int ecadk_destroy(ecadk_handle_t h) {
  if (h != nullptr && h->split_ws != nullptr) cudaFree(h->split_ws);
  delete h;
  return ECADK_OK;
}

int ecadk_set_splitk_force(int bn, int split) {
  ECADK_REQUIRE((bn == 0 && split == 0) || ((bn == 64 || bn == 128) && split >= 2 && split <= kSplitMax),
                "set_splitk_force: bn=%d split=%d (bn 64|128, split 2..%d, or 0, 0)", bn, split, kSplitMax);
  g_split_force = SplitForce{bn, split};
  return ECADK_OK;
}

long long ecadk_splitk_launches(void) { return g_splitk_launches.load(std::memory_order_relaxed); }

int ecadk_set_splitk_workspace(void* workspace, size_t bytes) {
  ECADK_REQUIRE((workspace == nullptr) == (bytes == 0), "set_splitk_workspace: pointer and size must agree");
  ECADK_REQUIRE(workspace == nullptr || aligned16(workspace), "set_splitk_workspace: 16-byte alignment");
  g_split_ws = SplitWs{workspace, bytes};
  return ECADK_OK;
}

int ecadk_pixart_text_kv(ecadk_handle_t h, const void* enc, int samples, int text_tokens, int text_pad,
                         void* const* k2, void* const* v2, int* n_launches, ecadk_stream_t stream) {
  ECADK_REQUIRE(h && enc && k2 && v2, "text_kv: null argument");
  SplitWsScope split_scope(h->split_ws, kSplitWsBytes);
  const EcadkModelDesc& d = h->desc;
  int launches = 0;
  for (int b = 0; b < d.num_layers; ++b) {
    const EcadkBlockWeights& w = h->blocks[b];
    int rc = ecadk_gemm_bias_headmajor(enc, w.w_kv2, w.b_kv2, k2[b], v2[b], nullptr, 2, d.heads, text_tokens,
                                       text_pad, samples * text_tokens, d.dim, stream);
    if (rc) return rc;
    ++launches;
  }
  if (n_launches) *n_launches = launches;
  return ECADK_OK;
}

int ecadk_pixart_blocks(ecadk_handle_t h, const EcadkBlocksArgs* a, const uint8_t* executed, int* n_launches,
                        ecadk_stream_t stream_) {
  ECADK_REQUIRE(h, "pixart_blocks: null handle");
  return ecadk_pixart_blocks_range(h, a, executed, 0, h->desc.num_layers, n_launches, stream_);
}

int ecadk_pixart_blocks_range(ecadk_handle_t h, const EcadkBlocksArgs* a, const uint8_t* executed, int block_begin,
                              int block_end, int* n_launches, ecadk_stream_t stream_) {
  ECADK_REQUIRE(h && a && executed, "pixart_blocks: null argument");
  SplitWsScope split_scope(h->split_ws, kSplitWsBytes);  // small-batch GEMMs of this call may split along K
  const EcadkModelDesc& d = h->desc;
  ECADK_REQUIRE(block_begin >= 0 && block_begin <= block_end && block_end <= d.num_layers,
                "pixart_blocks: block range [%d, %d) outside [0, %d]", block_begin, block_end, d.num_layers);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int D = d.dim, M = a->samples * a->tokens, S6 = a->temb_stride;
  ECADK_REQUIRE(S6 == 0 || S6 == 6 * D, "pixart_blocks: temb_stride must be 0 or 6*dim");
  ECADK_REQUIRE(a->tokens > 0 && a->tokens % 256 == 0, "pixart_blocks: tokens=%d must be a multiple of 256", a->tokens);
  ECADK_REQUIRE(a->text_pad > 0 && a->text_pad % 128 == 0, "pixart_blocks: text_pad=%d must be a multiple of 128",
                a->text_pad);
  int launches = 0;
  int rc;
  NvtxRange nv_call("ecadk_pixart_blocks [%d,%d) rows=%d", block_begin, block_end, M);

  EcadkResidualLnArgs pend;  // pending cached-residual reuse, flushed lazily into the next kernel that reads x
  memset(&pend, 0, sizeof(pend));
  auto reset_pend = [&]() {
    memset(&pend, 0, sizeof(pend));
    pend.x = a->x;
    pend.rows = M;
    pend.tokens = a->tokens;
    pend.dim = D;
    pend.temb_stride = S6;
    pend.eps = d.norm_eps;
  };
  reset_pend();
  auto flush = [&]() -> int {
    if (pend.n_reuse == 0 && pend.h == nullptr && pend.xb == nullptr) return ECADK_OK;
    int r = launch_residual_ln(pend, stream);
    ++launches;
    reset_pend();
    return r;
  };
  auto push_reuse = [&](const void* cache, const float* gate_table, const float* gate_temb) -> int {
    if (pend.n_reuse == ECADK_MAX_REUSE) {
      int r = flush();
      if (r) return r;
    }
    pend.reuse[pend.n_reuse++] = EcadkReuse{cache, gate_table, gate_temb};
    return ECADK_OK;
  };

  for (int b = block_begin; b < block_end; ++b) {
    const EcadkBlockWeights& w = h->blocks[b];
    const float* tab = w.scale_shift_table;
    const bool ex1 = executed[b * 3 + 0], ex2 = executed[b * 3 + 1], ex3 = executed[b * 3 + 2];
    void* c1 = a->cache[b * 3 + 0];
    void* c2 = a->cache[b * 3 + 1];
    void* c3 = a->cache[b * 3 + 2];
    // dead-store elimination: an executed sub-block whose cache slot is overwritten before any step reads it
    const uint8_t* dead = a->cache_dead;
    void* s1 = (dead && dead[b * 3 + 0]) ? nullptr : c1;
    void* s2 = (dead && dead[b * 3 + 1]) ? nullptr : c2;
    void* s3 = (dead && dead[b * 3 + 2]) ? nullptr : c3;
    bool xb_valid = false;

    // ---- attn1 (cached_transformer_block.py:208-246)
    if (ex1) {
      NvtxRange nv("b%02d.attn1", b);
      pend.h = a->h;
      pend.shift_table = tab + 0 * D;
      pend.scale_table = tab + 1 * D;
      pend.shift_temb = a->temb6 + 0 * D;
      pend.scale_temb = a->temb6 + 1 * D;
      if ((rc = flush())) return rc;
      if (a->qkv != nullptr) {
        // plain [M, 3*dim] projection output; the attention kernel gathers its (sample, head) tiles from it through
        // 3-D tensor maps (no head-major scatter epilogue, no padded copies of Q/K/V in HBM)
        __nv_bfloat16* qkv = static_cast<__nv_bfloat16*>(a->qkv);
        if ((rc = ecadk_gemm_bias(a->h, w.w_qkv1, w.b_qkv1, qkv, M, 3 * D, D, 3 * D, 0, stream_))) return rc;
        if ((rc = launch_attention_ex(qkv, 3 * D, qkv + D, qkv + 2 * D, 3 * D, a->self_bias, a->attn_o, a->samples,
                                      d.heads, a->tokens, a->tokens, stream)))
          return rc;
      } else {
        if ((rc = ecadk_gemm_bias_headmajor(a->h, w.w_qkv1, w.b_qkv1, a->q, a->k, a->v, 3, d.heads, a->tokens,
                                            a->tokens, M, D, stream_)))
          return rc;
        if ((rc = launch_attention(a->q, a->k, a->v, a->self_bias, a->attn_o, a->samples, d.heads, a->tokens,
                                   a->tokens, stream)))
          return rc;
      }
      if ((rc = ecadk_gemm_bias_gated_residual_cache(a->attn_o, w.w_out1, w.b_out1, a->x, ex2 ? a->xb : nullptr, s1,
                                                     tab + 2 * D, a->temb6 + 2 * D, S6, a->tokens, M, D, D, stream_)))
        return rc;
      launches += 3;
      xb_valid = ex2;
    } else {
      if ((rc = push_reuse(c1, tab + 2 * D, a->temb6 + 2 * D))) return rc;
    }

    // ---- attn2 on the un-normalised stream (:264-289)
    if (ex2) {
      NvtxRange nv("b%02d.attn2", b);
      if (!xb_valid) {
        pend.xb = a->xb;
        if ((rc = flush())) return rc;
      }
      if (a->qkv != nullptr) {  // (256 padded text keys + bias: launch_attention_ex streams them, see `biased256`)
        __nv_bfloat16* q2 = static_cast<__nv_bfloat16*>(a->qkv);  // the query projection reuses the first dim columns
        if ((rc = ecadk_gemm_bias(a->xb, w.w_q2, w.b_q2, q2, M, D, D, 3 * D, 0, stream_))) return rc;
        if ((rc = launch_attention_ex(q2, 3 * D, a->k2[b], a->v2[b], 0, a->text_bias, a->attn_o, a->samples, d.heads,
                                      a->tokens, a->text_pad, stream)))
          return rc;
      } else {
        if ((rc = ecadk_gemm_bias_headmajor(a->xb, w.w_q2, w.b_q2, a->q, nullptr, nullptr, 1, d.heads, a->tokens,
                                            a->tokens, M, D, stream_)))
          return rc;
        if ((rc = launch_attention(a->q, a->k2[b], a->v2[b], a->text_bias, a->attn_o, a->samples, d.heads, a->tokens,
                                   a->text_pad, stream)))
          return rc;
      }
      if ((rc = ecadk_gemm_bias_gated_residual_cache(a->attn_o, w.w_out2, w.b_out2, a->x, nullptr, s2, nullptr,
                                                     nullptr, S6, a->tokens, M, D, D, stream_)))
        return rc;
      launches += 3;
    } else {
      if ((rc = push_reuse(c2, nullptr, nullptr))) return rc;
    }

    // ---- feed-forward (:306-320)
    if (ex3) {
      NvtxRange nv("b%02d.ff", b);
      pend.h = a->h;
      pend.shift_table = tab + 3 * D;
      pend.scale_table = tab + 4 * D;
      pend.shift_temb = a->temb6 + 3 * D;
      pend.scale_temb = a->temb6 + 4 * D;
      if ((rc = flush())) return rc;
      if ((rc = ecadk_gemm_bias(a->h, w.w_ff1, w.b_ff1, a->ffh, M, d.ff_dim, D, d.ff_dim, 1, stream_))) return rc;
      if ((rc = ecadk_gemm_bias_gated_residual_cache(a->ffh, w.w_ff2, w.b_ff2, a->x, nullptr, s3, tab + 5 * D,
                                                     a->temb6 + 5 * D, S6, a->tokens, M, D, d.ff_dim, stream_)))
        return rc;
      launches += 2;
    } else {
      if ((rc = push_reuse(c3, tab + 5 * D, a->temb6 + 5 * D))) return rc;
    }
  }
  if ((rc = flush())) return rc;
  if (n_launches) *n_launches = launches;
  return ECADK_OK;
}

// ---------------------------------------------------------------------------------------------------
// VAE decoder pieces (vae.cuh + the implicit-GEMM mode of the tcgen05 GEMMs)
// ---------------------------------------------------------------------------------------------------
int ecadk_conv_nhwc(const void* x, const void* w, const float* bias, const void* residual, void* out, int batch, int h,
                    int w_, int c_in, int c_out, int out_ld, int out_cols, int taps, ecadk_stream_t stream) {
  ECADK_REQUIRE(x && w && out, "conv_nhwc: null pointer");
  ECADK_REQUIRE(batch > 0 && h > 0 && w_ > 0, "conv_nhwc: bad image size %d x %d x %d", batch, h, w_);
  ECADK_REQUIRE(taps == 1 || taps == 9, "conv_nhwc: taps=%d (1 = 1x1, 9 = 3x3)", taps);
  ECADK_REQUIRE(c_in > 0 && c_in % 64 == 0, "conv_nhwc: c_in=%d must be a multiple of 64 (pad with zero channels)", c_in);
  ECADK_REQUIRE(c_out > 0 && c_out % 128 == 0, "conv_nhwc: c_out=%d must be a multiple of 128 (pad the weight rows)", c_out);
  ECADK_REQUIRE(out_cols > 0 && out_cols <= c_out && out_ld % 8 == 0 && out_ld >= ((out_cols + 31) / 32) * 32,
                "conv_nhwc: out_cols=%d out_ld=%d", out_cols, out_ld);
  ECADK_REQUIRE(aligned16(out) && (residual == nullptr || aligned16(residual)), "conv_nhwc: 16-byte alignment");
  const long long rows = static_cast<long long>(batch) * (h + 2) * (w_ + 2);
  ECADK_REQUIRE(rows < (1ll << 31), "conv_nhwc: %lld rows", rows);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = static_cast<int>(rows);
  p.N = c_out;
  p.K = taps * c_in;
  p.bias = bias;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = out_ld;
  p.out2 = static_cast<__nv_bfloat16*>(const_cast<void*>(residual));
  p.ldo2 = out_ld;
  p.tokens = 1;
  p.conv_taps = taps;
  p.conv_cblocks = c_in / kGemmBK;
  p.conv_pitch = w_ + 2;
  for (int t = 0; t < 9; ++t) p.conv_off[t] = taps == 9 ? (t / 3 - 1) * (w_ + 2) + (t % 3 - 1) : 0;
  p.conv_h = h;
  p.conv_w = w_;
  p.conv_cols = out_cols;
  return launch_gemm<EPI_CONV>(x, w, p, static_cast<cudaStream_t>(stream));
}

int ecadk_conv_up2x_nhwc(const void* x, const void* w4, const float* bias, void* out, int batch, int h, int w_, int c_in,
                         int c_out, ecadk_stream_t stream_) {
  ECADK_REQUIRE(x && w4 && out, "conv_up2x_nhwc: null pointer");
  ECADK_REQUIRE(batch > 0 && h > 0 && w_ > 0, "conv_up2x_nhwc: bad image size %d x %d x %d", batch, h, w_);
  ECADK_REQUIRE(c_in > 0 && c_in % 64 == 0 && c_out > 0 && c_out % 128 == 0, "conv_up2x_nhwc: c_in=%d c_out=%d", c_in, c_out);
  ECADK_REQUIRE(aligned16(out) && aligned16(w4), "conv_up2x_nhwc: 16-byte alignment");
  const long long rows = static_cast<long long>(batch) * (h + 2) * (w_ + 2);
  const long long rows_out = static_cast<long long>(batch) * (2 * h + 2) * (2 * w_ + 2);
  ECADK_REQUIRE(rows_out < (1ll << 31) && batch <= 65535, "conv_up2x_nhwc: %lld output rows", rows_out);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  {
    ZeroBorderParams z{static_cast<__nv_bfloat16*>(out), batch, 2 * h, 2 * w_, c_out};
    const int n = (2 * (2 * w_ + 2) + 2 * (2 * h)) * (c_out / 8);
    zero_border_kernel<<<dim3((n + 255) / 256, batch), 256, 0, stream>>>(z);
    int rc = check_launch("zero_border_kernel");
    if (rc) return rc;
  }
  for (int a = 0; a < 2; ++a) {
    for (int b = 0; b < 2; ++b) {
      GemmParams p;
      memset(&p, 0, sizeof(p));
      p.M = static_cast<int>(rows);
      p.N = c_out;
      p.K = 4 * c_in;
      p.bias = bias;
      p.out = static_cast<__nv_bfloat16*>(out);
      p.ldo = c_out;
      p.tokens = 1;
      p.conv_taps = 4;
      p.conv_cblocks = c_in / kGemmBK;
      p.conv_pitch = w_ + 2;
      // taps (ry, rx): source rows {y-1, y} for output parity 0, {y, y+1} for parity 1 (same for columns)
      for (int t = 0; t < 4; ++t) p.conv_off[t] = ((t >> 1) + a - 1) * (w_ + 2) + ((t & 1) + b - 1);
      p.conv_h = h;
      p.conv_w = w_;
      p.conv_cols = c_out;
      p.conv_up = 1;
      p.conv_up_a = a;
      p.conv_up_b = b;
      const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(w4) + static_cast<size_t>(a * 2 + b) * c_out * 4 * c_in;
      int rc = launch_gemm<EPI_CONV>(x, wp, p, stream);
      if (rc) return rc;
    }
  }
  return ECADK_OK;
}

size_t ecadk_groupnorm_scratch_bytes(int batch, int h, int w_, int groups) {
  const long long plane = static_cast<long long>(h + 2) * (w_ + 2);
  const long long blocks = (plane + kGnPixelsPerBlock - 1) / kGnPixelsPerBlock;
  return static_cast<size_t>(batch) * groups * sizeof(float2) * static_cast<size_t>(blocks + 1);
}

int ecadk_groupnorm_nhwc(const void* x, const float* gamma, const float* beta, void* out, void* scratch, int batch, int h,
                         int w_, int c, int groups, float eps, int silu, int unpadded_out, ecadk_stream_t stream_) {
  ECADK_REQUIRE(x && gamma && beta && out && scratch, "groupnorm_nhwc: null pointer");
  ECADK_REQUIRE(batch > 0 && h > 0 && w_ > 0 && c > 0 && groups > 0 && c % groups == 0, "groupnorm_nhwc: bad shape");
  ECADK_REQUIRE(unpadded_out == 0 || unpadded_out >= h * w_, "groupnorm_nhwc: unpadded_out=%d must be 0 or >= h*w",
                unpadded_out);
  ECADK_REQUIRE(static_cast<long long>(h + 2) * (w_ + 2) * (c / 8) < (1ll << 31) && batch <= 65535,
                "groupnorm_nhwc: image too large");
  const int cpg = c / groups;
  ECADK_REQUIRE(cpg == 4 || cpg == 8 || cpg == 16, "groupnorm_nhwc: %d channels per group (supported: 4, 8, 16)", cpg);
  ECADK_REQUIRE(c % 8 == 0 && c <= 2048 && 256 % (c / 8) == 0, "groupnorm_nhwc: c=%d", c);
  ECADK_REQUIRE(aligned16(x) && aligned16(out) && aligned16(scratch) && aligned16(gamma) && aligned16(beta),
                "groupnorm_nhwc: 16-byte alignment");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long plane = static_cast<long long>(h + 2) * (w_ + 2);
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 3.0 * batch * h * w_ * c * 2.0, stream);
  const int blocks = static_cast<int>((plane + kGnPixelsPerBlock - 1) / kGnPixelsPerBlock);
  GroupNormParams p;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.gamma = gamma;
  p.beta = beta;
  p.mean_rstd = static_cast<float2*>(scratch);
  p.partial = p.mean_rstd + static_cast<size_t>(batch) * groups;
  p.B = batch; p.H = h; p.W = w_; p.C = c; p.G = groups;
  p.blocks = blocks;
  p.eps = eps;
  p.silu = silu;
  p.unpadded_out = unpadded_out;
  dim3 sgrid(blocks, batch);
  gn_stats_kernel<<<sgrid, 256, 256 * 4 * sizeof(float), stream>>>(p);
  gn_finalize_kernel<<<(batch * groups + 255) / 256, 256, 0, stream>>>(p);
  dim3 agrid(static_cast<unsigned>(((plane + kGnApplyPix - 1) / kGnApplyPix * (c / 8) + 255) / 256), batch);
  gn_apply_kernel<<<agrid, 256, 0, stream>>>(p);
  return check_launch("groupnorm_nhwc");
}

int ecadk_upsample2x_nhwc(const void* x, void* out, int batch, int h, int w_, int c, ecadk_stream_t stream_) {
  ECADK_REQUIRE(x && out && batch > 0 && h > 0 && w_ > 0 && c > 0 && c % 8 == 0, "upsample2x_nhwc: bad arguments");
  ECADK_REQUIRE(aligned16(x) && aligned16(out), "upsample2x_nhwc: 16-byte alignment");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 5.0 * batch * h * w_ * c * 2.0, stream);
  UpsampleParams p{static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), batch, h, w_, c};
  const long long per_sample = static_cast<long long>(2 * h + 2) * (2 * w_ + 2) * (c / 8);
  ECADK_REQUIRE(per_sample < (1ll << 31), "upsample2x_nhwc: image too large");
  upsample2x_kernel<<<dim3(static_cast<unsigned>((per_sample + 255) / 256), batch), 256, 0, stream>>>(p);
  return check_launch("upsample2x_kernel");
}

int ecadk_softmax_rows(const float* scores, void* probs, int rows, int cols, int valid_cols, float scale,
                       ecadk_stream_t stream_) {
  ECADK_REQUIRE(scores && probs && rows > 0 && cols > 0 && cols % 4 == 0, "softmax_rows: bad arguments");
  ECADK_REQUIRE(valid_cols > 0 && valid_cols <= cols && valid_cols % 4 == 0, "softmax_rows: valid_cols=%d of %d", valid_cols,
                cols);
  ECADK_REQUIRE(aligned16(scores) && aligned16(probs), "softmax_rows: 16-byte alignment");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 6.0 * rows * cols, stream);
  softmax_rows_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(scores, static_cast<__nv_bfloat16*>(probs), rows, cols,
                                                         valid_cols, scale * 1.4426950408889634f);
  return check_launch("softmax_rows_kernel");
}

int ecadk_vae_prepare_latents(const float* z, const float* pq_w, const float* pq_b, float inv_scaling, float shift,
                              void* out, int batch, int latent_channels, int h, int w_, ecadk_stream_t stream_) {
  ECADK_REQUIRE(z && out && batch > 0 && h > 0 && w_ > 0, "vae_prepare_latents: bad arguments");
  ECADK_REQUIRE(latent_channels == 4 || latent_channels == 16, "vae_prepare_latents: latent_channels=%d (4 or 16)",
                latent_channels);
  ECADK_REQUIRE((pq_w == nullptr) == (pq_b == nullptr), "vae_prepare_latents: post_quant_conv weight and bias go together");
  ECADK_REQUIRE(aligned16(out), "vae_prepare_latents: 16-byte alignment");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, stream);
  VaePrepParams p{z, pq_w, pq_b, static_cast<__nv_bfloat16*>(out), batch, h, w_, latent_channels, inv_scaling, shift};
  const long long total = static_cast<long long>(batch) * (h + 2) * (w_ + 2) * 8;
  vae_prepare_latents_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(p);
  return check_launch("vae_prepare_latents_kernel");
}

int ecadk_vae_add_tokens(const void* x, const void* tokens, int tokens_per_sample, void* out, int batch, int h, int w_,
                         int c, ecadk_stream_t stream_) {
  ECADK_REQUIRE(x && tokens && out && batch > 0 && h > 0 && w_ > 0 && c > 0 && c % 8 == 0, "vae_add_tokens: bad arguments");
  ECADK_REQUIRE(tokens_per_sample >= h * w_, "vae_add_tokens: tokens_per_sample=%d < %d", tokens_per_sample, h * w_);
  ECADK_REQUIRE(aligned16(x) && aligned16(tokens) && aligned16(out), "vae_add_tokens: 16-byte alignment");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, stream);
  AddTokensParams p{static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(tokens),
                    static_cast<__nv_bfloat16*>(out), batch, h, w_, c, tokens_per_sample};
  const long long per_sample = static_cast<long long>(h + 2) * (w_ + 2) * (c / 8);
  ECADK_REQUIRE(per_sample < (1ll << 31), "vae_add_tokens: image too large");
  vae_add_tokens_kernel<<<dim3(static_cast<unsigned>((per_sample + 255) / 256), batch), 256, 0, stream>>>(p);
  return check_launch("vae_add_tokens_kernel");
}

int ecadk_vae_finish(const void* y, float* image, int batch, int h, int w_, int denormalize, ecadk_stream_t stream_) {
  ECADK_REQUIRE(y && image && batch > 0 && h > 0 && w_ > 0, "vae_finish: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ProfScope prof(ECADK_PROF_OTHER, 0.0, 0.0, stream);
  VaeFinishParams p{static_cast<const __nv_bfloat16*>(y), image, batch, h, w_, denormalize};
  const long long total = static_cast<long long>(batch) * h * w_;
  vae_finish_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(p);
  return check_launch("vae_finish_kernel");
}

}  // extern "C"
