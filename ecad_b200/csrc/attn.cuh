// Single-pass tcgen05 attention tile for short key sequences (NK <= 256 keys: PixArt 256x256 self-attention,
// N = 256, and cross-attention against T5 tokens padded to 128).  One CTA = 128 queries x NK keys of one
// (sample, head); the whole score row lives in TMEM, so there is no online-softmax rescaling:
//
//   TMA   : Q[128 x 80], K[NK x 80], V[NK x 80] tiles (head_dim 72 zero-padded to 80 = one 64-wide
//           128B-swizzled chunk + one 16-wide 32B-swizzled chunk)
//   MMA 1 : S[128 x NK]  = Q K^T            (tcgen05, fp32 in TMEM, 5 K-steps of 16)
//   warps : softmax over the TMEM row (scale, additive key bias, exp2), P -> bf16 -> smem (K-major, SW128)
//   MMA 2 : O[128 x 80]  = P V              (V consumed MN-major straight from its [key, d] TMA tile)
//   warps : O / rowsum -> bf16 -> [sample, token, head*72 + e]
//
// Two CTAs fit per SM (104 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's TMA/MMA.
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace ecadk {

constexpr int kHeadDim = 72;
constexpr int kHeadPad = 80;
constexpr int kAttnBM = 128;
constexpr int kAttnThreads = 160;

struct AttnParams {
  int heads;
  int q_tokens;       // queries per (sample, head) in the Q layout (multiple of 128)
  int out_ld;         // row pitch of the output in elements (= heads * 72)
  float scale_log2e;  // (1/sqrt(d)) * log2(e)
  const float* bias;  // [samples, NK] additive key bias in natural-log units (0 / -10000 / -inf), or null
  __nv_bfloat16* out;  // [samples, q_tokens, out_ld]
  // optional split of the query sequence (FLUX double-stream blocks: text tokens first): rows q < split_tokens go to
  // out_lo [samples, split_tokens, out_ld], the others to out [samples, q_tokens - split_tokens, out_ld]
  __nv_bfloat16* out_lo;
  int split_tokens;
  // Operand layout seen through the (3-D) tensor maps: 0 = head-major [sample*heads + head][token][80] (the middle
  // coordinate is always 0), 1 = ROW-major [sample*tokens + token][head][72] - the plain output of the projection GEMM,
  // head = middle coordinate; the 8 padding columns of the 80-wide shared-memory tiles are the tensor map's
  // out-of-bounds zero fill.  attn_pair2_kernel / attn_flash_kernel<72> only.
  int q_rowmajor;
  int kv_rowmajor;
};

// Phase timing of the flash kernel (instrumented builds only: -DECADK_ATTN_TIMING; tools/micro/attn_phase_timing.py)
#ifdef ECADK_ATTN_TIMING
__device__ unsigned int g_attn_dbg[148 * 32];
#define ATTN_T(var) const unsigned int var = clock()
#define ATTN_ACC(slot, a, b) dbg_acc[slot] += (b) - (a)
#else
#define ATTN_T(var)
#define ATTN_ACC(slot, a, b)
#endif

template <int NK>
struct AttnCfg {
  // shared-memory map (bytes); every chunk base is a multiple of 1024
  static constexpr int kQ64 = 0;                         // 128 rows x 128 B  (SW128)
  static constexpr int kQ16 = kQ64 + kAttnBM * 128;      // 128 rows x  32 B  (SW32)
  static constexpr int kK64 = kQ16 + kAttnBM * 32;       // NK rows x 128 B
  static constexpr int kK16 = kK64 + NK * 128;           // NK rows x 32 B
  static constexpr int kQKEnd = kK16 + NK * 32;
  static constexpr int kPBytes = kAttnBM * NK * 2;       // P aliases Q/K once S is complete
  static constexpr int kV64 = (kQKEnd > kPBytes ? kQKEnd : kPBytes);
  static constexpr int kV16 = kV64 + NK * 128;
  static constexpr int kBars = kV16 + NK * 32;
  static constexpr int kSmemBytes = kBars + 64 + 1024;
  static constexpr int kTmemCols = NK < 128 ? 128 : NK;  // S needs NK fp32 columns; O reuses columns [0,80)
  static constexpr uint32_t kBytesQK = (kAttnBM + NK) * kHeadPad * 2;
  static constexpr uint32_t kBytesV = NK * kHeadPad * 2;
};

template <int NK, bool HAS_BIAS>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_tile_kernel(const __grid_constant__ CUtensorMap tm_q64, const __grid_constant__ CUtensorMap tm_q16,
                 const __grid_constant__ CUtensorMap tm_k64, const __grid_constant__ CUtensorMap tm_k16,
                 const __grid_constant__ CUtensorMap tm_v64, const __grid_constant__ CUtensorMap tm_v16,
                 const AttnParams p) {
  using Cfg = AttnCfg<NK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + Cfg::kBars);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_p = bar_qk + 3;
  uint64_t* bar_o = bar_qk + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int sh = blockIdx.y;  // sample * heads + head
  const int sample = sh / p.heads;
  const int head = sh - sample * p.heads;

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  griddep_launch_dependents();  // programmatic dependent launch: see gemm_bf16_kernel
  griddep_wait();

  if (warp == 4) {
    if (lane == 0) {
      // ---- TMA: Q + K on one barrier, V on another so S = QK^T can start before V lands
      const int q_row = sh * p.q_tokens + q_tile * kAttnBM;
      const int k_row = sh * NK;
      mbar_arrive_expect_tx(bar_qk, Cfg::kBytesQK);
      tma_load_2d(smem + Cfg::kQ64, &tm_q64, bar_qk, 0, q_row);
      tma_load_2d(smem + Cfg::kQ16, &tm_q16, bar_qk, 64, q_row);
      tma_load_2d(smem + Cfg::kK64, &tm_k64, bar_qk, 0, k_row);
      tma_load_2d(smem + Cfg::kK16, &tm_k16, bar_qk, 64, k_row);
      mbar_arrive_expect_tx(bar_v, Cfg::kBytesV);
      tma_load_2d(smem + Cfg::kV64, &tm_v64, bar_v, 0, k_row);
      tma_load_2d(smem + Cfg::kV16, &tm_v16, bar_v, 64, k_row);

      // ---- MMA 1: S = Q K^T  (both operands K-major; K extent 80 = 4 x 16 (SW128) + 1 x 16 (SW32))
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      {
        constexpr uint32_t idesc = make_idesc_bf16(kAttnBM, NK);
        const uint32_t sbase = smem_u32(smem);
        const uint64_t dq = make_smem_desc(sbase + Cfg::kQ64, 0, 1024, kLayoutSW128);
        const uint64_t dk = make_smem_desc(sbase + Cfg::kK64, 0, 1024, kLayoutSW128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, dq + 2 * k, dk + 2 * k, idesc, k != 0);
        const uint64_t dq2 = make_smem_desc(sbase + Cfg::kQ16, 16, 256, kLayoutSW32);
        const uint64_t dk2 = make_smem_desc(sbase + Cfg::kK16, 16, 256, kLayoutSW32);
        umma_bf16_ss(tmem, dq2, dk2, idesc, 1);
        umma_commit(bar_s);
      }

      // ---- MMA 2: O = P V  (A = P K-major SW128 in 64-key chunks; B = V MN-major, N = 64 + 16)
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      {
        constexpr uint32_t idesc64 = make_idesc_bf16(kAttnBM, 64, 0, 1);
        constexpr uint32_t idesc16 = make_idesc_bf16(kAttnBM, 16, 0, 1);
        const uint32_t sbase = smem_u32(smem);
#pragma unroll
        for (int ks = 0; ks < NK / 16; ++ks) {
          const int chunk = ks >> 2, kin = ks & 3;
          const uint64_t dp =
              make_smem_desc(sbase + chunk * (kAttnBM * 128) + kin * 32, 0, 1024, kLayoutSW128);
          // V tile rows are keys: 16 keys per K-step = 16 rows of 128 B (SW128 part) / 32 B (SW32 part)
          const uint64_t dv64 = make_smem_desc(sbase + Cfg::kV64 + ks * 16 * 128, NK * 128, 1024, kLayoutSW128);
          const uint64_t dv16 = make_smem_desc(sbase + Cfg::kV16 + ks * 16 * 32, NK * 32, 256, kLayoutSW32);
          umma_bf16_ss(tmem, dp, dv64, idesc64, ks != 0);
          umma_bf16_ss(tmem + 64, dp, dv16, idesc16, ks != 0);
        }
        umma_commit(bar_o);
      }
    }
    __syncwarp();
  } else {
    // ---- softmax + epilogue warps: thread = one query row = one TMEM lane
    const int row = warp * 32 + lane;
    const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const float* bias = HAS_BIAS ? p.bias + static_cast<size_t>(sample) * NK : nullptr;
    constexpr float kLog2e = 1.4426950408889634f;

    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < NK / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float t = __uint_as_float(v[j]) * p.scale_log2e;
        if constexpr (HAS_BIAS) t += __ldg(bias + c * 32 + j) * kLog2e;
        mx = fmaxf(mx, t);
      }
    }
    float sum = 0.f;
    uint8_t* prow = smem + row * 128;
#pragma unroll 1
    for (int c = 0; c < NK / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + c * 32, v);
      tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float t = __uint_as_float(v[j]) * p.scale_log2e;
        if constexpr (HAS_BIAS) t += __ldg(bias + c * 32 + j) * kLog2e;
        e[j] = exp2f(t - mx);
        sum += e[j];
      }
      // keys [c*32, c*32+32) live in 64-key chunk (c>>1), 16-byte groups g0..g0+3 of the 128-byte row
      uint8_t* pchunk = prow + (c >> 1) * (kAttnBM * 128);
      const int g0 = (c & 1) * 4;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16x2(e[g * 8 + 0], e[g * 8 + 1]);
        o.y = pack_bf16x2(e[g * 8 + 2], e[g * 8 + 3]);
        o.z = pack_bf16x2(e[g * 8 + 4], e[g * 8 + 5]);
        o.w = pack_bf16x2(e[g * 8 + 6], e[g * 8 + 7]);
        *reinterpret_cast<uint4*>(pchunk + (((g0 + g) ^ (row & 7)) << 4)) = o;
      }
    }
    fence_proxy_async_smem();  // generic-proxy P writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);

    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    const int q = q_tile * kAttnBM + row;
    __nv_bfloat16* dst = p.out + (static_cast<size_t>(sample) * p.q_tokens + q) * p.out_ld + head * kHeadDim;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * inv, __uint_as_float(v[g * 8 + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = o;
      }
    }
    {
      uint32_t v[16];
      tmem_ld_32x16(t_row + 64, v);  // columns 64..79; 72..79 are padding
      tmem_ld_wait();
      uint4 o;
      o.x = pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
      o.y = pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
      o.z = pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
      o.w = pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
      *reinterpret_cast<uint4*>(dst + 64) = o;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::kTmemCols);
  }
}

}  // namespace ecadk

namespace ecadk {

// ---- softmax building blocks shared by attn_pair_kernel / attn_flash_kernel -------------------------------------
// One 32-column chunk of a score row (thread = row).  `bias_a` = shared-space address of this warp's key-bias slice,
// already multiplied by log2(e).  Packed fp32x2 math: one FFMA2 / FADD2 / FMNMX3 per two scores.
template <bool HAS_BIAS>
__device__ __forceinline__ float softmax_chunk_max(const uint32_t (&v)[32], float mx, float scale_log2e, uint32_t bias_a) {
  if constexpr (!HAS_BIAS) {
    // scale > 0: the maximum of the raw scores is taken here and scaled once by the caller.  Two independent chains:
    // 16 dependent FMNMX3 per chunk were ~100 clocks of pure latency in front of every block's exchange
    float m1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      mx = fmax3(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    }
    mx = fmaxf(mx, m1);
  } else {
    const uint64_t sc2 = pack_f2(scale_log2e, scale_log2e);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = lds_f4(bias_a + i * 4);
      float t0, t1, t2, t3;
      unpack_f2(ffma2(pack_f2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, pack_f2(b4.x, b4.y)), t0, t1);
      unpack_f2(ffma2(pack_f2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), sc2, pack_f2(b4.z, b4.w)), t2, t3);
      mx = fmax3(mx, t0, t1);
      mx = fmax3(mx, t2, t3);
    }
  }
  return mx;
}

// p = exp2(score * scale + bias - m); accumulates the row sum as a packed pair; writes 16 packed bf16x2 words
template <bool HAS_BIAS>
__device__ __forceinline__ void softmax_chunk_exp(const uint32_t (&v)[32], uint32_t (&pk)[16], uint64_t& sum2, float m,
                                                  float scale_log2e, uint32_t bias_a) {
  const uint64_t sc2 = pack_f2(scale_log2e, scale_log2e);
  const uint64_t nm2 = pack_f2(-m, -m);
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    uint64_t c01 = nm2, c23 = nm2;
    if constexpr (HAS_BIAS) {
      const float4 b4 = lds_f4(bias_a + i * 4);
      c01 = fadd2(pack_f2(b4.x, b4.y), nm2);
      c23 = fadd2(pack_f2(b4.z, b4.w), nm2);
    }
    float t0, t1, t2, t3;
    unpack_f2(ffma2(pack_f2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, c01), t0, t1);
    unpack_f2(ffma2(pack_f2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), sc2, c23), t2, t3);
    const float e0 = fast_exp2(t0), e1 = fast_exp2(t1), e2 = fast_exp2(t2), e3 = fast_exp2(t3);
    sum2 = fadd2(sum2, pack_f2(e0, e1));
    sum2 = fadd2(sum2, pack_f2(e2, e3));
    pk[i / 2] = pack_bf16x2(e0, e1);
    pk[i / 2 + 1] = pack_bf16x2(e2, e3);
  }
}

// Register-resident rows (the whole score row of a thread stays in registers between the two steps):
// prep: with a key bias the scaled+biased score replaces the raw one in place (so the bias is read once); without, the
// raw maximum is taken and scaled by the caller.  exp: p = exp2(t - m) resp. exp2(score * scale - m).
template <bool HAS_BIAS>
__device__ __forceinline__ float softmax_chunk_prep(uint32_t (&v)[32], float mx, float scale_log2e, uint32_t bias_a) {
  if constexpr (!HAS_BIAS) {
    float m1 = -INFINITY;  // two independent chains (see softmax_chunk_max)
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      mx = fmax3(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    }
    mx = fmaxf(mx, m1);
  } else {
    const uint64_t sc2 = pack_f2(scale_log2e, scale_log2e);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = lds_f4(bias_a + i * 4);
      float t0, t1, t2, t3;
      unpack_f2(ffma2(pack_f2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, pack_f2(b4.x, b4.y)), t0, t1);
      unpack_f2(ffma2(pack_f2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), sc2, pack_f2(b4.z, b4.w)), t2, t3);
      v[i] = __float_as_uint(t0);
      v[i + 1] = __float_as_uint(t1);
      v[i + 2] = __float_as_uint(t2);
      v[i + 3] = __float_as_uint(t3);
      mx = fmax3(mx, t0, t1);
      mx = fmax3(mx, t2, t3);
    }
  }
  return mx;
}

template <bool HAS_BIAS>
__device__ __forceinline__ void softmax_chunk_exp_reg(const uint32_t (&v)[32], uint32_t (&pk)[16], uint64_t& sum2, float m,
                                                      float scale_log2e) {
  const uint64_t sc2 = pack_f2(scale_log2e, scale_log2e);
  const uint64_t nm2 = pack_f2(-m, -m);
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float t0, t1, t2, t3;
    if constexpr (HAS_BIAS) {
      unpack_f2(fadd2(pack_f2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), nm2), t0, t1);
      unpack_f2(fadd2(pack_f2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), nm2), t2, t3);
    } else {
      unpack_f2(ffma2(pack_f2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2), t0, t1);
      unpack_f2(ffma2(pack_f2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), sc2, nm2), t2, t3);
    }
    const float e0 = fast_exp2(t0), e1 = fast_exp2(t1), e2 = fast_exp2(t2), e3 = fast_exp2(t3);
    sum2 = fadd2(sum2, pack_f2(e0, e1));
    sum2 = fadd2(sum2, pack_f2(e2, e3));
    pk[i / 2] = pack_bf16x2(e0, e1);
    pk[i / 2 + 1] = pack_bf16x2(e2, e3);
  }
}

// =====================================================================================================
// Persistent pipelined variant: one work item = all 256 queries of one (sample, head) against NK keys.
//
//   warp 0       TMA producer: Q (2 tiles) + K of item i+1 are prefetched as soon as item i's QK^T has retired;
//                V is double-buffered.  K/V are read ONCE per (sample, head).  (Measured: double-buffering Q/K as well
//                does not help, 110 -> 114 us at the config-2 shape - the ~2000 clk the MMA thread waits for Q/K per
//                item are the HBM transfer itself, not buffer availability.)
//   warp 1       MMA issuer:  S_t = Q_t K^T (SS), then O_t = P_t V with P_t read straight from TMEM (TS form).
//   warps 2..9   softmax + epilogue of query tile 0;  warps 10..17 of query tile 1 (two threads per query row).
//
// TMEM map (512 columns): tile t owns columns [256t, 256t+256): S_t fp32 in [0,NK); P_t (bf16 pairs) and O_t are
// written over consumed S columns (AttnPairCfg::kPHi / kO).  No shared memory is spent on P.
// =====================================================================================================
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

constexpr int kAttnPairThreads = 576;  // warp 0 TMA, warp 1 MMA, 16 softmax warps (two threads per score row)

template <int NK>
struct AttnPairCfg {
  static constexpr int kQ64 = 0;                    // 256 rows x 128 B (tile 1 at +16 KB)
  static constexpr int kQ16 = kQ64 + 256 * 128;     // 256 rows x 32 B  (tile 1 at +4 KB)
  static constexpr int kK64 = kQ16 + 256 * 32;
  static constexpr int kK16 = kK64 + NK * 128;
  static constexpr int kV = kK16 + NK * 32;         // 2 buffers of (NK*128 + NK*32)
  static constexpr int kVBuf = NK * 160;
  static constexpr int kBias = kV + 2 * kVBuf;      // 16 warps x NK/2 floats
  static constexpr int kXch = kBias + 8 * NK * 4;   // row-maximum / row-sum exchange between partner warps: 2 x [16][32]
  static constexpr int kBars = kXch + 2 * 16 * 32 * 4;
  // TMEM columns inside a tile's 256: with 256 keys S fills all of them, so P and O live over consumed S columns -
  // P of keys 0..127 in [0,64) (written by the threads that own S[0,128)), O in [64,144), P of keys 128..255 in
  // [144,208) (written by the threads that own S[128,256), always behind their own reads).  With 128 keys the row is
  // register-resident, S is dead after the exchange barrier, and the layout is simply P [0,64), O [128,208).
  static constexpr int kPHi = NK == 256 ? 144 : 32;  // column of the P half written by the "half 1" threads
  static constexpr int kO = NK == 256 ? 64 : 128;
  // the kernel allocates all 512 TMEM columns, so it must be alone on its SM: ask for more than half the smem
  static constexpr int kSmemBytes = (kBars + 256 + 1024) > 120 * 1024 ? (kBars + 256 + 1024) : 120 * 1024;
  static constexpr uint32_t kBytesQK = (256 + NK) * kHeadPad * 2;
  static constexpr uint32_t kBytesV = NK * kHeadPad * 2;
};

template <int NK, bool HAS_BIAS>
__global__ void __launch_bounds__(kAttnPairThreads, 1)
attn_pair_kernel(const __grid_constant__ CUtensorMap tm_q64, const __grid_constant__ CUtensorMap tm_q16,
                 const __grid_constant__ CUtensorMap tm_k64, const __grid_constant__ CUtensorMap tm_k16,
                 const __grid_constant__ CUtensorMap tm_v64, const __grid_constant__ CUtensorMap tm_v16,
                 const AttnParams p, const int num_items) {
  using Cfg = AttnPairCfg<NK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBars);
  uint64_t* qk_full = bars + 0;
  uint64_t* qk_empty = bars + 1;
  uint64_t* v_full = bars + 2;    // [2]
  uint64_t* v_empty = bars + 4;   // [2]
  uint64_t* s_full = bars + 6;    // [2] per query tile
  uint64_t* p_full = bars + 8;    // [2]
  uint64_t* o_full = bars + 10;   // [2]
  uint64_t* s_empty = bars + 12;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(qk_full, 1);
    mbar_init(qk_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  griddep_launch_dependents();  // programmatic dependent launch: see gemm_bf16_kernel
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int q_row = item * 256;
        const int k_row = item * NK;
        mbar_wait(qk_empty, (n & 1) ^ 1);
        mbar_arrive_expect_tx(qk_full, Cfg::kBytesQK);
        tma_load_2d(smem + Cfg::kQ64, &tm_q64, qk_full, 0, q_row);
        tma_load_2d(smem + Cfg::kQ16, &tm_q16, qk_full, 64, q_row);
        tma_load_2d(smem + Cfg::kK64, &tm_k64, qk_full, 0, k_row);
        tma_load_2d(smem + Cfg::kK16, &tm_k16, qk_full, 64, k_row);
        const int vb = n & 1;
        mbar_wait(&v_empty[vb], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[vb], Cfg::kBytesV);
        uint8_t* vdst = smem + Cfg::kV + vb * Cfg::kVBuf;
        // V lands as FIVE 16-column (32-byte-swizzled) atoms so that the whole 80-column head is one MN-major operand:
        // P V is ONE N = 80 instruction per 16-key step instead of an N = 64 plus an N = 16 one, and a narrow
        // tcgen05.mma costs about as much as a wide one (measured with ECADK_ATTN_TIMING)
#pragma unroll
        for (int a = 0; a < kHeadPad / 16; ++a) tma_load_2d(vdst + a * (NK * 32), &tm_v16, &v_full[vb], a * 16, k_row);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_s = make_idesc_bf16(kAttnBM, NK);
      constexpr uint32_t idesc_o80 = make_idesc_bf16(kAttnBM, kHeadPad, 0, 1);
      const uint32_t sbase = smem_u32(smem);
      // Issue order is software-pipelined so that the two query tiles run in anti-phase: while tile 0's warps are in
      // softmax(i), the tensor core serves tile 1's PV(i-1) and QK^T(i), and vice versa.  Each tile's chain
      // (softmax -> PV -> read-out -> next QK^T) is serial, so overlapping the two chains is what hides it.
      auto issue_qk = [&](int t) {
        const uint32_t d = tmem + 256 * t;
        const uint64_t dk = make_smem_desc(sbase + Cfg::kK64, 16, 1024, kLayoutSW128);
        const uint64_t dk2 = make_smem_desc(sbase + Cfg::kK16, 16, 256, kLayoutSW32);
        const uint64_t dq = make_smem_desc(sbase + Cfg::kQ64 + t * (kAttnBM * 128), 16, 1024, kLayoutSW128);
        const uint64_t dq2 = make_smem_desc(sbase + Cfg::kQ16 + t * (kAttnBM * 32), 16, 256, kLayoutSW32);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(d, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_bf16_ss(d, dq2, dk2, idesc_s, 1);
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int vb) {
        const uint32_t vbase = sbase + Cfg::kV + vb * Cfg::kVBuf;
        const uint32_t p_tmem = tmem + 256 * t;  // bf16 pairs: 8 columns per 16-key step
        const uint32_t o_tmem = tmem + 256 * t + Cfg::kO;
#pragma unroll
        for (int ks = 0; ks < NK / 16; ++ks) {
          const uint64_t dv = make_smem_desc(vbase + ks * 16 * 32, NK * 32, 256, kLayoutSW32);
          const uint32_t pa = ks < NK / 32 ? p_tmem + ks * 8 : p_tmem + Cfg::kPHi + (ks - NK / 32) * 8;
          umma_bf16_ts(o_tmem, pa, dv, idesc_o80, ks != 0);
        }
        umma_commit(&o_full[t]);
      };
      int n = 0;
#ifdef ECADK_ATTN_TIMING
      unsigned int dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const unsigned int t_begin = clock();
#endif
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const uint32_t par = n & 1;
        ATTN_T(m0);
        mbar_wait(qk_full, par);
        ATTN_T(m1);
        // tile 0: S(i)
        mbar_wait(&s_empty[0], par ^ 1);  // previous item's O_0 has been read out of TMEM
        tc_fence_after();
        ATTN_T(m2);
        issue_qk(0);
        ATTN_T(m3);
        // tile 1: PV(i-1)  (V(i-1) sits in buffer (n-1)&1 and was waited for in the previous iteration)
        if (n > 0) {
          mbar_wait(&p_full[1], par ^ 1);
          tc_fence_after();
          issue_pv(1, (n - 1) & 1);
          umma_commit(&v_empty[(n - 1) & 1]);
        }
        ATTN_T(m4);
        // tile 1: S(i)
        mbar_wait(&s_empty[1], par ^ 1);
        tc_fence_after();
        issue_qk(1);
        umma_commit(qk_empty);  // Q/K smem may be refilled with the next item
        ATTN_T(m5);
        // tile 0: PV(i)
        mbar_wait(&v_full[n & 1], (n >> 1) & 1);
        mbar_wait(&p_full[0], par);
        tc_fence_after();
        ATTN_T(m6);
        issue_pv(0, n & 1);
        ATTN_T(m7);
        ATTN_ACC(0, m0, m1);  // wait Q/K
        ATTN_ACC(1, m1, m2);  // wait O_0 read out
        ATTN_ACC(2, m2, m3);  // issue QK0
        ATTN_ACC(3, m3, m4);  // wait P1 + issue PV1
        ATTN_ACC(4, m4, m5);  // wait O_1 read out + issue QK1
        ATTN_ACC(5, m5, m6);  // wait V + P0
        ATTN_ACC(6, m6, m7);  // issue PV0
      }
#ifdef ECADK_ATTN_TIMING
      dbg_acc[7] = n;
      for (int i = 0; i < 8; ++i) g_attn_dbg[blockIdx.x * 32 + i] = dbg_acc[i];
      g_attn_dbg[blockIdx.x * 32 + 24] = clock() - t_begin;
#endif
      if (n > 0) {  // drain: tile 1's PV of the last item
        mbar_wait(&p_full[1], (n - 1) & 1);
        tc_fence_after();
        issue_pv(1, (n - 1) & 1);
        umma_commit(&v_empty[(n - 1) & 1]);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax + epilogue: warps 2..9 -> tile 0, warps 10..17 -> tile 1 =====================
    // Every score row is shared by two threads (the two warps of a tile that own the same TMEM lane quarter): "half 0"
    // takes keys [0, NK/2) and output columns 0..47, "half 1" keys [NK/2, NK) and columns 48..71.  The per-thread
    // exponential chain (~17 clk each) bounded this kernel; halving the row halves it.
    const int sw = warp - 2;
    const int t = sw >> 3;
    const int half = (sw >> 2) & 1;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t t_row = tmem + 256 * t + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t s_col = t_row + half * (NK / 2);                    // this thread's score columns
    const uint32_t p_col = t_row + (half ? Cfg::kPHi : 0);             // ... and where its P goes
    const uint32_t bias_a = smem_u32(smem + Cfg::kBias) + sw * (NK / 2) * 4;  // per-warp bias slice (shared space)
    const uint32_t xch = smem_u32(smem + Cfg::kXch);
    const uint32_t my_slot = xch + (sw * 32 + lane) * 4;
    const uint32_t peer_slot = xch + ((sw ^ 4) * 32 + lane) * 4;
    const int bar_id = 1 + t * 4 + quarter;
    constexpr int NC = NK / 64;  // 32-column chunks per thread
    constexpr float kLog2e = 1.4426950408889634f;
    // key bias of the NEXT item is fetched one item ahead (registers), so its global latency is off the chain
    float breg[NC];
    auto fetch_bias = [&](int it) {
      const float* b = p.bias + static_cast<size_t>(it / p.heads) * NK + half * (NK / 2);
#pragma unroll
      for (int j = 0; j < NC; ++j) breg[j] = __ldg(b + lane + 32 * j) * kLog2e;
    };
    if constexpr (HAS_BIAS) {
      if (static_cast<int>(blockIdx.x) < num_items) fetch_bias(blockIdx.x);
    }
    int n = 0;
#ifdef ECADK_ATTN_TIMING
    unsigned int dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
      const uint32_t par = n & 1;
      const int sample = item / p.heads;
      const int head = item - sample * p.heads;
      if constexpr (HAS_BIAS) {
#pragma unroll
        for (int j = 0; j < NC; ++j) sts_f1(bias_a + (lane + 32 * j) * 4, breg[j]);
        __syncwarp();
        if (item + static_cast<int>(gridDim.x) < num_items) fetch_bias(item + gridDim.x);
      }
      ATTN_T(s0);
      mbar_wait(&s_full[t], par);
      tc_fence_after();
      ATTN_T(s1);
      float mx = -INFINITY;
      uint64_t sum2 = pack_f2(0.f, 0.f);
      if constexpr (NK == 128) {
        // 64 scores per thread stay in registers: one TMEM read pass
        uint32_t v[NC][32];
#pragma unroll
        for (int c = 0; c < NC; ++c) tmem_ld_32x32(s_col + c * 32, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < NC; ++c) mx = softmax_chunk_prep<HAS_BIAS>(v[c], mx, p.scale_log2e, bias_a + c * 128);
        if constexpr (!HAS_BIAS) mx *= p.scale_log2e;
        sts_f1(my_slot, mx);
        named_bar_sync(bar_id, 64);  // also: the partner has finished reading its S columns
        mx = fmaxf(mx, lds_f1(peer_slot));
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          uint32_t pk[16];
          softmax_chunk_exp_reg<HAS_BIAS>(v[c], pk, sum2, mx, p.scale_log2e);
          tmem_st_32x16(p_col + c * 16, pk);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(s_col + c * 32, v);
          tmem_ld_wait();
          mx = softmax_chunk_max<HAS_BIAS>(v, mx, p.scale_log2e, bias_a + c * 128);
        }
        if constexpr (!HAS_BIAS) mx *= p.scale_log2e;
        sts_f1(my_slot, mx);
        named_bar_sync(bar_id, 64);
        mx = fmaxf(mx, lds_f1(peer_slot));
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(s_col + c * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
          softmax_chunk_exp<HAS_BIAS>(v, pk, sum2, mx, p.scale_log2e, bias_a + c * 128);
          // P chunk c of this thread -> packed columns [16c, 16c+16) of ITS P region: always behind its own S reads
          tmem_st_32x16(p_col + c * 16, pk);
        }
      }
      float sum_lo, sum_hi;
      unpack_f2(sum2, sum_lo, sum_hi);
      const float sum = sum_lo + sum_hi;
      sts_f1(my_slot + 16 * 32 * 4, sum);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      ATTN_T(s2);
      mbar_wait(&o_full[t], par);
      tc_fence_after();
      named_bar_sync(bar_id, 64);
      ATTN_T(s3);
      const float inv = 1.0f / (sum + lds_f1(peer_slot + 16 * 32 * 4));
      const int q = t * kAttnBM + row;
      __nv_bfloat16* dst = p.out + (static_cast<size_t>(sample) * p.q_tokens + q) * p.out_ld + head * kHeadDim;
      auto store8 = [&](const uint32_t* w, __nv_bfloat16* d) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(w[0]) * inv, __uint_as_float(w[1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(w[2]) * inv, __uint_as_float(w[3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(w[4]) * inv, __uint_as_float(w[5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(w[6]) * inv, __uint_as_float(w[7]) * inv);
        *reinterpret_cast<uint4*>(d) = o;
      };
      {
        // (measured: handing the TMEM columns back BEFORE the row-strided global stores - O parked in registers -
        // is slower for the d = 72 kernels, 110 -> 138 us at the config-2 shape, although it helps the d = 128 one)
        uint32_t v[32];
        tmem_ld_32x32(t_row + Cfg::kO + half * 48, v);  // half 0: columns 0..31, half 1: 48..79 (72..79 are padding)
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 3; ++g) store8(v + g * 8, dst + half * 48 + g * 8);
        if (half == 0) {
          store8(v + 24, dst + 24);
          uint32_t w[16];
          tmem_ld_32x16(t_row + Cfg::kO + 32, w);
          tmem_ld_wait();
          store8(w, dst + 32);
          store8(w + 8, dst + 40);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[t]);  // TMEM columns of tile t are free for the next item
      ATTN_T(s4);
      ATTN_ACC(0, s0, s1);  // wait S
      ATTN_ACC(1, s1, s2);  // softmax (max, exchange, exp, P store)
      ATTN_ACC(2, s2, s3);  // wait O
      ATTN_ACC(3, s3, s4);  // O read-out + global stores
    }
#ifdef ECADK_ATTN_TIMING
    if (lane == 0 && (sw == 0 || sw == 12)) {
      dbg_acc[7] = n;
      for (int i = 0; i < 8; ++i) g_attn_dbg[blockIdx.x * 32 + (sw == 0 ? 8 : 16) + i] = dbg_acc[i];
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}


// =====================================================================================================
// Second-generation 256-query kernel (round 2): same math, TMEM map and softmax as attn_pair_kernel, but every stage
// of an item is decoupled so that HBM streams continuously while MUFU and the tensor core stay busy:
//
//   * K and V are double-buffered PER ITEM, the two 128-query Q tiles have their own slots that are refilled as soon
//     as their S = Q K^T has retired - every TMA load is issued a full item period before it is consumed (the first
//     kernel waited ~2000 clk per item for Q/K: its single Q/K buffer could only be refilled half an item ahead);
//   * the output leaves through shared memory and ONE TMA store per tile (warp 18) instead of row-strided 16-byte
//     stores from the softmax threads: a warp store used to touch 32 different 128-byte lines (32 LSU passes per
//     instruction, ~2000 clk of read-out per tile per item), now the softmax warps write conflict-free 16-byte pieces
//     of dense 144-byte rows into smem, hand the tile's TMEM columns back at once and go on to the next item;
//   * ONE staged output tile (18 KB) is shared by the two query tiles, which run in anti-phase; the store warp hands it
//     back as soon as the TMA store has read it.  (First version: with 256 keys the staged tiles lived in the retired
//     K buffer of the same item - that made the NEXT K load wait for the item's stores, which queue behind the loads
//     in the TMA unit; the MMA thread then waited ~900 clk per item for K.)
// =====================================================================================================
constexpr int kPair2Threads = 608;  // warp 0 TMA loads, warp 1 MMA, warps 2..17 softmax, warp 18 TMA stores

template <int NK, bool HAS_BIAS>
struct AttnPair2Cfg {
  static constexpr int kQTile = kAttnBM * 160;        // [128 x 128 B SW128][128 x 32 B SW32]
  static constexpr int kQ = 0;                        // 2 tile slots
  static constexpr int kKBuf = NK * 160;              // [NK x 128 B SW128][NK x 32 B SW32]
  static constexpr int kK = kQ + 2 * kQTile;          // 2 buffers
  static constexpr int kVBuf = NK * 160;              // five [NK x 32 B] SW32 atoms
  static constexpr int kV = kK + 2 * kKBuf;           // 2 buffers
  static constexpr int kOTile = kAttnBM * kHeadDim * 2;  // 18432: dense [128 x 72] bf16
  static constexpr int kO = kV + 2 * kVBuf;           // ONE staged output tile, shared by both query tiles
  static constexpr int kBias = kO + kOTile;           // 16 warps x NK/2 floats (only with a key bias)
  static constexpr int kXch = kBias + (HAS_BIAS ? 8 * NK * 4 : 0);
  static constexpr int kBars = kXch + 2 * 16 * 32 * 4;
  static constexpr int kPHi = NK == 256 ? 144 : 32;   // TMEM map: see AttnPairCfg
  static constexpr int kOCol = NK == 256 ? 64 : 128;
  static constexpr int kSmemNeed = kBars + 256 + 1024;
  static constexpr int kSmemBytes = kSmemNeed > 120 * 1024 ? kSmemNeed : 120 * 1024;  // alone on its SM (512 TMEM columns)
  static constexpr uint32_t kBytesQ = kAttnBM * kHeadPad * 2;
  static constexpr uint32_t kBytesKV = NK * kHeadPad * 2;
  static_assert(kSmemBytes <= 227 * 1024, "attn_pair2_kernel: shared memory");
  static_assert(kOTile % 1024 == 0 && kKBuf % 1024 == 0 && kQTile % 1024 == 0, "alignment");
};

template <int NK, bool HAS_BIAS>
__global__ void __launch_bounds__(kPair2Threads, 1)
attn_pair2_kernel(const __grid_constant__ CUtensorMap tm_q64, const __grid_constant__ CUtensorMap tm_q16,
                  const __grid_constant__ CUtensorMap tm_k64, const __grid_constant__ CUtensorMap tm_k16,
                  const __grid_constant__ CUtensorMap tm_v16, const __grid_constant__ CUtensorMap tm_o,
                  const AttnParams p, const int num_items) {
  using Cfg = AttnPair2Cfg<NK, HAS_BIAS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBars);
  uint64_t* q_full = bars + 0;     // [2] per query-tile slot, one phase per item
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* k_full = bars + 4;     // [2] per buffer, one phase per two items
  uint64_t* k_empty = bars + 6;    // [2]
  uint64_t* v_full = bars + 8;     // [2]
  uint64_t* v_empty = bars + 10;   // [2]
  uint64_t* s_full = bars + 12;    // [2] per query tile, one phase per item
  uint64_t* p_full = bars + 14;    // [2]
  uint64_t* o_full = bars + 16;    // [2]
  uint64_t* s_empty = bars + 18;   // [2]
  uint64_t* o_staged = bars + 20;  // [2] per query tile, one phase per item: the bf16 output tile is in shared memory
  // [2] per query tile, one phase per item: the tile's store has read the staging tile.  ONE barrier with a phase per
  // store (the tiles alternate) is ambiguous: a tile that tests "store k - 1 done" by parity also passes when store k - 2
  // (its OWN previous store) is still pending - which happens when the bulk store queues behind an item's worth of
  // strided operand gathers (row-major operands, 200 samples): tile 1 then staged too early and its rows received
  // tile 0's output (tools/micro/determinism_stress.py: 11 % of the launches, and every launch once the MMA issue
  // got faster).  With one barrier per tile the tested store is never more than one phase away.
  uint64_t* stage_free = bars + 22;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  // (the shuffles tell ptxas that the values are warp-uniform - see the MMA issuer of attn_flash_kernel)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q64);
    tma_prefetch_desc(&tm_k64);
    tma_prefetch_desc(&tm_v16);
    tma_prefetch_desc(&tm_o);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_empty[i], 8);
      mbar_init(&o_staged[i], 8);
    }
    mbar_init(&stage_free[0], 1);
    mbar_init(&stage_free[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  griddep_launch_dependents();  // programmatic dependent launch: see gemm_bf16_kernel
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer: runs one item ahead of the MMA warp =====================
      int n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int b = n & 1;
        const uint32_t ph2 = ((n >> 1) & 1) ^ 1;  // "slot is free" parity of the per-buffer barriers
        const int sample = item / p.heads, hd = item - sample * p.heads;
        const int q_mid = p.q_rowmajor ? hd : 0, kv_mid = p.kv_rowmajor ? hd : 0;
        const int q_row = (p.q_rowmajor ? sample : item) * 256;
        const int k_row = (p.kv_rowmajor ? sample : item) * NK;
        mbar_wait(&k_empty[b], ph2);
        uint8_t* kd = smem + Cfg::kK + b * Cfg::kKBuf;
        mbar_arrive_expect_tx(&k_full[b], Cfg::kBytesKV);
        tma_load_3d(kd, &tm_k64, &k_full[b], 0, kv_mid, k_row);
        tma_load_3d(kd + NK * 128, &tm_k16, &k_full[b], 64, kv_mid, k_row);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&q_empty[t], (n & 1) ^ 1);
          uint8_t* qd = smem + Cfg::kQ + t * Cfg::kQTile;
          mbar_arrive_expect_tx(&q_full[t], Cfg::kBytesQ);
          tma_load_3d(qd, &tm_q64, &q_full[t], 0, q_mid, q_row + t * kAttnBM);
          tma_load_3d(qd + kAttnBM * 128, &tm_q16, &q_full[t], 64, q_mid, q_row + t * kAttnBM);
        }
        mbar_wait(&v_empty[b], ph2);
        uint8_t* vd = smem + Cfg::kV + b * Cfg::kVBuf;
        mbar_arrive_expect_tx(&v_full[b], Cfg::kBytesKV);
        // V lands as FIVE 16-column (32-byte-swizzled) atoms: the 80-column head is one MN-major operand (N = 80)
#pragma unroll
        for (int a = 0; a < kHeadPad / 16; ++a) tma_load_3d(vd + a * (NK * 32), &tm_v16, &v_full[b], a * 16, kv_mid, k_row);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      // ===================== MMA issuer (whole warp, elected-lane tcgen05 instructions: see attn_flash_kernel) =======
      // Static anti-phase order QK0(i), PV1(i-1), QK1(i), PV0(i) with blocking (hardware-suspended) barrier waits.
      // Measured and rejected (round 2): a dynamic order in which this thread polls both tiles' barriers and issues
      // whatever is ready - 128 keys 83 -> 88 us, 256 keys 99 -> 120 us: the probe loop shares its SM sub-partition's
      // issue port with four softmax warps (with a nanosleep back-off it is no better); starting P V on the first half
      // of the probabilities (128 keys only - with 256 keys O lives over S columns that are consumed last) and issuing
      // the next S = Q K^T before the read-out: no gain either.
      constexpr uint32_t idesc_s = make_idesc_bf16(kAttnBM, NK);
      constexpr uint32_t idesc_o80 = make_idesc_bf16(kAttnBM, kHeadPad, 0, 1);
      const uint32_t sbase = smem_u32(smem);
      auto issue_qk = [&](int t, int b) {
        const uint32_t d = tmem + 256 * t;
        const uint32_t kb = sbase + Cfg::kK + b * Cfg::kKBuf;
        const uint32_t qb = sbase + Cfg::kQ + t * Cfg::kQTile;
        const uint64_t dk = make_smem_desc(kb, 16, 1024, kLayoutSW128);
        const uint64_t dk2 = make_smem_desc(kb + NK * 128, 16, 256, kLayoutSW32);
        const uint64_t dq = make_smem_desc(qb, 16, 1024, kLayoutSW128);
        const uint64_t dq2 = make_smem_desc(qb + kAttnBM * 128, 16, 256, kLayoutSW32);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(d, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          umma_bf16_ss(d, dq2, dk2, idesc_s, 1);
          umma_commit(&s_full[t]);
          umma_commit(&q_empty[t]);
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
      };
      auto issue_pv = [&](int t, int b) {
        const uint32_t vbase = sbase + Cfg::kV + b * Cfg::kVBuf;
        const uint32_t p_tmem = tmem + 256 * t;  // bf16 pairs: 8 columns per 16-key step
        const uint32_t o_tmem = tmem + 256 * t + Cfg::kOCol;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < NK / 16; ++ks) {
            const uint64_t dv = make_smem_desc(vbase + ks * 16 * 32, NK * 32, 256, kLayoutSW32);
            const uint32_t pa = ks < NK / 32 ? p_tmem + ks * 8 : p_tmem + Cfg::kPHi + (ks - NK / 32) * 8;
            umma_bf16_ts(o_tmem, pa, dv, idesc_o80, ks != 0);
          }
          umma_commit(&o_full[t]);
        }
        __syncwarp();
      };
      int n = 0;
#ifdef ECADK_ATTN_TIMING
      unsigned int dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const unsigned int t_begin = clock();
#endif
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int b = n & 1;
        const uint32_t par = n & 1;
        const uint32_t ph2 = (n >> 1) & 1;
        ATTN_T(m0);
        mbar_wait(&k_full[b], ph2);
        mbar_wait(&q_full[0], par);
        ATTN_T(m1);
        mbar_wait(&s_empty[0], par ^ 1);  // previous item's O_0 has left TMEM
        tc_fence_after();
        ATTN_T(m2);
        issue_qk(0, b);
        ATTN_T(m3);
        if (n > 0) {  // tile 1: PV of the previous item (anti-phase with tile 0)
          mbar_wait(&p_full[1], par ^ 1);
          tc_fence_after();
          issue_pv(1, b ^ 1);
          commit(&v_empty[b ^ 1]);
        }
        ATTN_T(m4);
        mbar_wait(&q_full[1], par);
        mbar_wait(&s_empty[1], par ^ 1);
        tc_fence_after();
        issue_qk(1, b);
        commit(&k_empty[b]);
        ATTN_T(m5);
        mbar_wait(&v_full[b], ph2);
        mbar_wait(&p_full[0], par);
        tc_fence_after();
        ATTN_T(m6);
        issue_pv(0, b);
        ATTN_T(m7);
        ATTN_ACC(0, m0, m1);  // wait K + Q0
        ATTN_ACC(1, m1, m2);  // wait O_0 read out
        ATTN_ACC(2, m2, m3);  // issue QK0
        ATTN_ACC(3, m3, m4);  // wait P1 + issue PV1
        ATTN_ACC(4, m4, m5);  // wait Q1 + O_1 read out + issue QK1
        ATTN_ACC(5, m5, m6);  // wait V + P0
        ATTN_ACC(6, m6, m7);  // issue PV0
      }
#ifdef ECADK_ATTN_TIMING
      dbg_acc[7] = n;
      if (lane == 0) {
        for (int i = 0; i < 8; ++i) g_attn_dbg[blockIdx.x * 32 + i] = dbg_acc[i];
        g_attn_dbg[blockIdx.x * 32 + 24] = clock() - t_begin;
      }
#endif
      if (n > 0) {  // drain: tile 1's PV of the last item
        const int b = (n - 1) & 1;
        mbar_wait(&p_full[1], (n - 1) & 1);
        tc_fence_after();
        issue_pv(1, b);
        commit(&v_empty[b]);
      }
    }
    __syncwarp();
  } else if (warp == 18) {
    if (lane == 0) {
      // ===================== TMA stores of the staged output tiles =====================
      // store k = 2 n + t (the MMA order makes the tiles alternate: O_0(n), O_1(n), O_0(n+1), ...)
      int n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int sample = item / p.heads;
        const int head = item - sample * p.heads;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&o_staged[t], n & 1);
          tma_store_2d(&tm_o, smem + Cfg::kO, head * kHeadDim, sample * 256 + t * kAttnBM);
          tma_store_commit();
          tma_store_wait_read0();  // the staging tile has been read: the other query tile may overwrite it
          mbar_arrive(&stage_free[t]);
        }
      }
      tma_store_wait_all0();
    }
    __syncwarp();
  } else {
    // ===================== softmax + read-out: warps 2..9 -> tile 0, warps 10..17 -> tile 1 =====================
    // (two threads per score row, see attn_pair_kernel)
    const int sw = warp - 2;
    const int t = sw >> 3;
    const int half = (sw >> 2) & 1;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t t_row = tmem + 256 * t + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t s_col = t_row + half * (NK / 2);
    const uint32_t p_col = t_row + (half ? Cfg::kPHi : 0);
    const uint32_t bias_a = smem_u32(smem + Cfg::kBias) + sw * (NK / 2) * 4;
    const uint32_t xch = smem_u32(smem + Cfg::kXch);
    const uint32_t my_slot = xch + (sw * 32 + lane) * 4;
    const uint32_t peer_slot = xch + ((sw ^ 4) * 32 + lane) * 4;
    const int bar_id = 1 + t * 4 + quarter;
    constexpr int NC = NK / 64;  // 32-column chunks per thread
    constexpr float kLog2e = 1.4426950408889634f;
    float breg[NC];
    auto fetch_bias = [&](int it) {
      const float* bp = p.bias + static_cast<size_t>(it / p.heads) * NK + half * (NK / 2);
#pragma unroll
      for (int j = 0; j < NC; ++j) breg[j] = __ldg(bp + lane + 32 * j) * kLog2e;
    };
    if constexpr (HAS_BIAS) {
      if (static_cast<int>(blockIdx.x) < num_items) fetch_bias(blockIdx.x);
    }
    int n = 0;
#ifdef ECADK_ATTN_TIMING
    unsigned int dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
      const uint32_t par = n & 1;
      if constexpr (HAS_BIAS) {
#pragma unroll
        for (int j = 0; j < NC; ++j) sts_f1(bias_a + (lane + 32 * j) * 4, breg[j]);
        __syncwarp();
        if (item + static_cast<int>(gridDim.x) < num_items) fetch_bias(item + gridDim.x);
      }
      ATTN_T(s0);
      mbar_wait(&s_full[t], par);
      tc_fence_after();
      ATTN_T(s1);
      float mx = -INFINITY;
      uint64_t sum2 = pack_f2(0.f, 0.f);
      if constexpr (NK == 128) {
        uint32_t v[NC][32];
#pragma unroll
        for (int c = 0; c < NC; ++c) tmem_ld_32x32(s_col + c * 32, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < NC; ++c) mx = softmax_chunk_prep<HAS_BIAS>(v[c], mx, p.scale_log2e, bias_a + c * 128);
        if constexpr (!HAS_BIAS) mx *= p.scale_log2e;
        sts_f1(my_slot, mx);
        named_bar_sync(bar_id, 64);  // also: the partner has finished reading its S columns
        mx = fmaxf(mx, lds_f1(peer_slot));
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          uint32_t pk[16];
          softmax_chunk_exp_reg<HAS_BIAS>(v[c], pk, sum2, mx, p.scale_log2e);
          tmem_st_32x16(p_col + c * 16, pk);
        }
      } else {
        // 256 keys: 128 scores per thread do not fit in registers - two passes over TMEM, two 32-column loads in flight
        // per wait (halves the exposed TMEM latencies of the chunk-at-a-time loop: 95 -> 93 us at the config-2 shape)
#pragma unroll 1
        for (int c = 0; c < NC; c += 2) {
          uint32_t v[2][32];
          tmem_ld_32x32(s_col + c * 32, v[0]);
          tmem_ld_32x32(s_col + c * 32 + 32, v[1]);
          tmem_ld_wait();
          mx = softmax_chunk_max<HAS_BIAS>(v[0], mx, p.scale_log2e, bias_a + c * 128);
          mx = softmax_chunk_max<HAS_BIAS>(v[1], mx, p.scale_log2e, bias_a + c * 128 + 128);
        }
        if constexpr (!HAS_BIAS) mx *= p.scale_log2e;
        sts_f1(my_slot, mx);
        named_bar_sync(bar_id, 64);
        mx = fmaxf(mx, lds_f1(peer_slot));
#pragma unroll 1
        for (int c = 0; c < NC; c += 2) {
          uint32_t v[2][32];
          tmem_ld_32x32(s_col + c * 32, v[0]);
          tmem_ld_32x32(s_col + c * 32 + 32, v[1]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t pk[16];
            softmax_chunk_exp<HAS_BIAS>(v[j], pk, sum2, mx, p.scale_log2e, bias_a + (c + j) * 128);
            tmem_st_32x16(p_col + (c + j) * 16, pk);  // always behind this thread's own S reads
          }
        }
      }
      float sum_lo, sum_hi;
      unpack_f2(sum2, sum_lo, sum_hi);
      const float sum = sum_lo + sum_hi;
      sts_f1(my_slot + 16 * 32 * 4, sum);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      ATTN_T(s2);
      mbar_wait(&o_full[t], par);
      tc_fence_after();
      named_bar_sync(bar_id, 64);
      ATTN_T(s3);
      const float inv = 1.0f / (sum + lds_f1(peer_slot + 16 * 32 * 4));
      // O_t leaves TMEM into registers; the tile's columns go straight back to the MMA warp
      uint32_t v[32], w[16];
      tmem_ld_32x32(t_row + Cfg::kOCol + half * 48, v);  // half 0: columns 0..31, half 1: 48..79 (72..79 are padding)
      if (half == 0) tmem_ld_32x16(t_row + Cfg::kOCol + 32, w);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[t]);
      ATTN_T(s3a);
      // the shared staging tile is free once the OTHER tile's previous store has read it: tile 1 follows tile 0's store
      // of this item, tile 0 follows tile 1's store of the previous item
      mbar_wait(&stage_free[t ^ 1], t == 1 ? par : (par ^ 1));
      ATTN_T(s3b);
      const uint32_t srow = smem_u32(smem + Cfg::kO) + row * (kHeadDim * 2) + half * 96;
      auto stage8 = [&](const uint32_t* x, uint32_t addr) {
        sts_u4(addr, pack_bf16x2(__uint_as_float(x[0]) * inv, __uint_as_float(x[1]) * inv),
               pack_bf16x2(__uint_as_float(x[2]) * inv, __uint_as_float(x[3]) * inv),
               pack_bf16x2(__uint_as_float(x[4]) * inv, __uint_as_float(x[5]) * inv),
               pack_bf16x2(__uint_as_float(x[6]) * inv, __uint_as_float(x[7]) * inv));
      };
#pragma unroll
      for (int g = 0; g < 3; ++g) stage8(v + g * 8, srow + g * 16);
      if (half == 0) {
        stage8(v + 24, srow + 48);
        stage8(w, srow + 64);
        stage8(w + 8, srow + 80);
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA store (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_staged[t]);
      ATTN_T(s4);
      ATTN_ACC(0, s0, s1);  // wait S
      ATTN_ACC(1, s1, s2);  // softmax (max, exchange, exp, P store)
      ATTN_ACC(2, s2, s3);  // wait O
      ATTN_ACC(3, s3, s3a);   // O out of TMEM (tile handed back)
      ATTN_ACC(4, s3a, s3b);  // wait for the staging tile
      ATTN_ACC(5, s3b, s4);   // staging writes + proxy fence
    }
#ifdef ECADK_ATTN_TIMING
    if (lane == 0 && (sw == 0 || sw == 12)) {
      dbg_acc[7] = n;
      for (int i = 0; i < 8; ++i) g_attn_dbg[blockIdx.x * 32 + (sw == 0 ? 8 : 16) + i] = dbg_acc[i];
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace ecadk

namespace ecadk {

// =====================================================================================================
// General flash-style variant for long key sequences (PixArt 512x512 / 1024x1024 self-attention, N = 1024 / 4096 keys,
// and PixArt-sigma cross-attention, 300 text tokens padded to 384): keys are streamed in blocks of 128 with an online
// softmax.  Same skeleton as attn_pair_kernel - one work item = 256 queries (two anti-phase tiles) of one
// (sample, head); S_t (128 fp32 columns), P_t (bf16, in place over S_t) and O_t (80 columns) live in TMEM; K and V
// blocks flow through 2-stage TMA rings.
//
// Rescaling of O is LAZY: the running reference maximum m_used only moves when the block maximum exceeds it by more
// than 8 (log2 units), so P entries stay <= 2^8 and the TMEM round trip  O *= 2^(m_old - m_new)  is rare.
// =====================================================================================================
constexpr int kFlashKB = 128;  // keys per block
// warp 0 TMA, warp 1 MMA, warps 2..17 softmax: EIGHT warps per query tile - every score row is shared by two threads
// (the two warps that own the same TMEM lane quarter), each taking 64 of the 128 keys of a block.  A tile's softmax is
// a serial chain per thread (~17 clk per exponential per warp), and the chain - not MUFU throughput - bounds the kernel:
// halving the per-thread row halves the chain.
constexpr int kFlashThreads = 576;

// HD = real head dim: 72 (PixArt; stored padded to 80 = 64 + 16 columns) or 128 (FLUX; 64 + 64 columns).  The head is
// always staged as a 64-column 128B-swizzled chunk plus a second chunk of kC2 columns (32B- or 128B-swizzled).
template <int HD>
struct AttnFlashCfg {
  static constexpr int kPad = HD == 72 ? 80 : HD;     // columns per head in the Q/K/V layout
  static constexpr int kC2 = kPad - 64;               // columns of the second chunk (16 or 64)
  static constexpr int kRow2 = kC2 * 2;               // bytes per row of the second chunk (32 or 128)
  static constexpr uint32_t kSBO2 = 8 * kRow2;        // 8-row group stride of the second chunk
  static constexpr uint64_t kLayout2 = kC2 == 16 ? kLayoutSW32 : kLayoutSW128;
  static constexpr int kQ64 = 0;                      // 256 rows x 128 B
  static constexpr int kQ2 = kQ64 + 256 * 128;        // 256 rows x kRow2
  static constexpr int kKStage = kFlashKB * (128 + kRow2);
  static constexpr int kK = kQ2 + 256 * kRow2;        // 2 stages
  static constexpr int kV = kK + 2 * kKStage;         // 2 stages
  static constexpr int kBias = kV + 2 * kKStage;      // 16 warps x 64 floats
  static constexpr int kXch = kBias + 8 * kFlashKB * 4;  // row-maximum exchange [2 buffers][16 warps][32] + row-sum [16][32]
  static constexpr int kBars = kXch + 3 * 16 * 32 * 4;
  static constexpr int kSmemBytes = kBars + 256 + 1024;
  static constexpr uint32_t kBytesQ = 256 * kPad * 2;
  static constexpr uint32_t kBytesKV = kFlashKB * kPad * 2;
  static_assert(HD == 72 || HD == 128, "head dims built: 72 (PixArt), 128 (FLUX)");
  static_assert(kSmemBytes > 114 * 1024 && kSmemBytes <= 227 * 1024,
                "attn_flash_kernel must be alone on its SM (it owns all 512 TMEM columns) and fit in shared memory");
};

template <int HD, bool HAS_BIAS>
__global__ void __launch_bounds__(kFlashThreads, 1)
attn_flash_kernel(const __grid_constant__ CUtensorMap tm_q64, const __grid_constant__ CUtensorMap tm_q16,
                  const __grid_constant__ CUtensorMap tm_k64, const __grid_constant__ CUtensorMap tm_k16,
                  const __grid_constant__ CUtensorMap tm_v64, const __grid_constant__ CUtensorMap tm_v16,
                  const AttnParams p, const int n_keys, const int num_items) {
  using Cfg = AttnFlashCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBars);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;    // [2]
  uint64_t* k_empty = bars + 4;   // [2]
  uint64_t* v_full = bars + 6;    // [2]
  uint64_t* v_empty = bars + 8;   // [2]
  uint64_t* s_full = bars + 10;   // [2] per query tile, one phase per key block
  uint64_t* p_full = bars + 12;   // [2]
  uint64_t* o_full = bars + 14;   // [2] one phase per item
  uint64_t* s_empty = bars + 16;  // [2]
  uint64_t* p_half = bars + 18;   // [2] the first 32 keys of both row halves of P are in TMEM (P V can start on them)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  // (the shuffles tell ptxas that the values are warp-uniform - see the MMA issuer)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nkb = n_keys / kFlashKB;            // key blocks per item
  const int pairs = p.q_tokens / 256;           // 256-query pairs per (sample, head)

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&p_half[i], 8);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  griddep_launch_dependents();  // programmatic dependent launch: see gemm_bf16_kernel
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int n = 0, nb = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int sh = item / pairs, pr = item - sh * pairs;
        const int smp = sh / p.heads, hd = sh - smp * p.heads;
        const int q_mid = p.q_rowmajor ? hd : 0, kv_mid = p.kv_rowmajor ? hd : 0;
        const int q_row = (p.q_rowmajor ? smp : sh) * p.q_tokens + pr * 256;
        mbar_wait(q_empty, (n & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, Cfg::kBytesQ);
        tma_load_3d(smem + Cfg::kQ64, &tm_q64, q_full, 0, q_mid, q_row);
        tma_load_3d(smem + Cfg::kQ2, &tm_q16, q_full, 64, q_mid, q_row);
        for (int j = 0; j < nkb; ++j, ++nb) {
          const int st = nb & 1;
          const uint32_t ph = ((nb >> 1) & 1) ^ 1;
          const int k_row = (p.kv_rowmajor ? smp : sh) * n_keys + j * kFlashKB;
          uint8_t* kd = smem + Cfg::kK + st * Cfg::kKStage;
          uint8_t* vd = smem + Cfg::kV + st * Cfg::kKStage;
          mbar_wait(&k_empty[st], ph);
          mbar_arrive_expect_tx(&k_full[st], Cfg::kBytesKV);
          tma_load_3d(kd, &tm_k64, &k_full[st], 0, kv_mid, k_row);
          tma_load_3d(kd + kFlashKB * 128, &tm_k16, &k_full[st], 64, kv_mid, k_row);
          mbar_wait(&v_empty[st], ph);
          mbar_arrive_expect_tx(&v_full[st], Cfg::kBytesKV);
          if constexpr (HD == 128) {
            tma_load_3d(vd, &tm_v64, &v_full[st], 0, kv_mid, k_row);
            tma_load_3d(vd + kFlashKB * 128, &tm_v16, &v_full[st], 64, kv_mid, k_row);
          } else {  // five 16-column SW32 atoms: the 80-column head is one MN-major operand (see attn_pair_kernel)
#pragma unroll
            for (int a = 0; a < Cfg::kPad / 16; ++a)
              tma_load_3d(vd + a * (kFlashKB * 32), &tm_v16, &v_full[st], a * 16, kv_mid, k_row);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      // ===================== MMA issuer =====================
      // The WHOLE warp runs this loop; only the tcgen05 instructions are predicated on one elected lane.  Issued from
      // inside an `if (lane == 0)` branch ptxas cannot prove the operands warp-uniform and moves each one into a uniform
      // register through an ELECT / R2UR.BROADCAST / BRA.U.ANY loop - ~125 clocks of the issuing thread per tcgen05.mma,
      // which for head dim 72 (26 MMAs of 40-64 clocks per key block) was the whole period (tools/micro/
      // mma_issue_bench.cu, profiles/r2_flash2_phase_timing.txt)
      constexpr uint32_t idesc_s = make_idesc_bf16(kAttnBM, kFlashKB);
      constexpr uint32_t idesc_o80 = make_idesc_bf16(kAttnBM, 80, 0, 1);
      constexpr uint32_t idesc_o128 = make_idesc_bf16(kAttnBM, 128, 0, 1);
      const uint32_t sbase = smem_u32(smem);
      auto issue_qk = [&](int t, int st) {
        const uint32_t d = tmem + 256 * t;
        const uint32_t kb = sbase + Cfg::kK + st * Cfg::kKStage;
        const uint64_t dk = make_smem_desc(kb, 16, 1024, kLayoutSW128);
        const uint64_t dk2 = make_smem_desc(kb + kFlashKB * 128, 16, Cfg::kSBO2, Cfg::kLayout2);
        const uint64_t dq = make_smem_desc(sbase + Cfg::kQ64 + t * (kAttnBM * 128), 16, 1024, kLayoutSW128);
        const uint64_t dq2 = make_smem_desc(sbase + Cfg::kQ2 + t * (kAttnBM * Cfg::kRow2), 16, Cfg::kSBO2, Cfg::kLayout2);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(d, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < Cfg::kC2 / 16; ++k) umma_bf16_ss(d, dq2 + 2 * k, dk2 + 2 * k, idesc_s, 1);
          umma_commit(&s_full[t]);
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
      };
      // P V of one key block in two parts: part 0 covers the 16-key steps whose P is written FIRST by the softmax warps
      // (keys 0..31 by the "half 0" threads, 64..95 by the "half 1" threads), part 1 the rest - so the tensor core
      // starts on a tile's P V while the second half of its exponentials is still being computed
      auto issue_pv = [&](int t, int st, bool accumulate, int part) {
        const uint32_t vb = sbase + Cfg::kV + st * Cfg::kKStage;
        const uint32_t p_tmem = tmem + 256 * t;
        const uint32_t o_tmem = tmem + 256 * t + 128;
        if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ks = (i >> 1) * 4 + part * 2 + (i & 1);  // part 0: 0,1,4,5   part 1: 2,3,6,7
          const uint32_t acc = (accumulate || part != 0 || i != 0) ? 1u : 0u;
          if constexpr (HD == 128) {
            // both 64-column swizzle atoms of V in ONE N = 128 instruction (LBO = atom stride along the head dim):
            // two N = 64 instructions cost about as much as two N = 128 ones (measured with ECADK_ATTN_TIMING)
            const uint64_t dv64 = make_smem_desc(vb + ks * 16 * 128, kFlashKB * 128, 1024, kLayoutSW128);
            umma_bf16_ts(o_tmem, p_tmem + ks * 8, dv64, idesc_o128, acc);
          } else {
            const uint64_t dv = make_smem_desc(vb + ks * 16 * 32, kFlashKB * 32, 256, kLayoutSW32);
            umma_bf16_ts(o_tmem, p_tmem + ks * 8, dv, idesc_o80, acc);
          }
        }
        }
        __syncwarp();
      };
      // tile 1's PV of block nb-1 is deferred by one block so the two tiles run in anti-phase
      bool pend = false, pend_last = false, pend_acc = false;
      int pend_st = 0;
      uint32_t pend_par = 0;
      auto flush_pending = [&]() {
        if (!pend) return;
        mbar_wait(&p_half[1], pend_par);
        tc_fence_after();
        issue_pv(1, pend_st, pend_acc, 0);
        mbar_wait(&p_full[1], pend_par);
        tc_fence_after();
        issue_pv(1, pend_st, true, 1);
        commit(&v_empty[pend_st]);
        if (pend_last) commit(&o_full[1]);
        pend = false;
      };
      int n = 0, nb = 0;
#ifdef ECADK_ATTN_TIMING
      unsigned int dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const unsigned int t_begin = clock();
#endif
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        mbar_wait(q_full, n & 1);
        for (int j = 0; j < nkb; ++j, ++nb) {
          const int st = nb & 1;
          const uint32_t ring_ph = (nb >> 1) & 1;
          const uint32_t blk_par = nb & 1;
          ATTN_T(m0);
          mbar_wait(&k_full[st], ring_ph);
          if (j == 0) mbar_wait(&s_empty[0], (n & 1) ^ 1);  // previous item's O_0 read out
          tc_fence_after();
          ATTN_T(m1);
          issue_qk(0, st);
          ATTN_T(m2);
          flush_pending();  // tile 1: PV of the previous block (its QK^T for this block comes next)
          ATTN_T(m3);
          if (j == 0) mbar_wait(&s_empty[1], (n & 1) ^ 1);
          tc_fence_after();
          issue_qk(1, st);
          commit(&k_empty[st]);
          if (j == nkb - 1) commit(q_empty);
          ATTN_T(m4);
          mbar_wait(&v_full[st], ring_ph);
          ATTN_T(m5);
          mbar_wait(&p_half[0], blk_par);
          tc_fence_after();
          issue_pv(0, st, j != 0, 0);
          mbar_wait(&p_full[0], blk_par);
          tc_fence_after();
          ATTN_T(m6);
          ATTN_ACC(0, m0, m1);  // wait K (+ s_empty at item start)
          ATTN_ACC(1, m1, m2);  // issue QK0
          ATTN_ACC(2, m2, m3);  // wait P1 + issue PV1
          ATTN_ACC(3, m3, m4);  // issue QK1
          ATTN_ACC(4, m4, m5);  // wait V
          ATTN_ACC(5, m5, m6);  // wait P0
          issue_pv(0, st, true, 1);
          if (j == nkb - 1) commit(&o_full[0]);
          pend = true;
          pend_st = st;
          pend_par = blk_par;
          pend_acc = j != 0;
          pend_last = j == nkb - 1;
        }
      }
      flush_pending();
#ifdef ECADK_ATTN_TIMING
      dbg_acc[6] = clock() - t_begin;
      dbg_acc[7] = nb;
      if (lane == 0) {
        for (int i = 0; i < 8; ++i) g_attn_dbg[blockIdx.x * 32 + i] = dbg_acc[i];
      }
#endif
    }
    __syncwarp();
  } else {
    // ===================== softmax + correction + epilogue =====================
    const int sw = warp - 2;
    const int t = sw >> 3;            // query tile
    const int half = (sw >> 2) & 1;   // which 64 keys of a block / which output columns this thread owns
    const int quarter = warp & 3;     // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t t_row = tmem + 256 * t + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t bias_a = smem_u32(smem + Cfg::kBias) + sw * (kFlashKB / 2) * 4;
    // exchange slots of this thread and of its partner (same tile, same quarter, other half)
    const uint32_t xch = smem_u32(smem + Cfg::kXch);
    const uint32_t my_slot = xch + (sw * 32 + lane) * 4;
    const uint32_t peer_slot = xch + ((sw ^ 4) * 32 + lane) * 4;
    const int bar_id = 1 + t * 4 + quarter;  // named barrier of the two partner warps
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr int kOSplit = Cfg::kPad == 128 ? 64 : 48;  // O columns [0, kOSplit) -> half 0, [kOSplit, kPad) -> half 1
    float breg[2];
    auto fetch_bias = [&](int sample, int j) {
      const float* b = p.bias + static_cast<size_t>(sample) * n_keys + j * kFlashKB + half * 64;
      breg[0] = __ldg(b + lane) * kLog2e;
      breg[1] = __ldg(b + lane + 32) * kLog2e;
    };
    int n = 0, nb = 0;
#ifdef ECADK_ATTN_TIMING
    unsigned int dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned int t_begin = clock();
#endif
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
      const int sh = item / pairs, pr = item - sh * pairs;
      const int sample = sh / p.heads, head = sh - sample * p.heads;
      float m_used = -INFINITY, l_sum = 0.f;  // l_sum: this thread's 64-key share of the row sum
      if constexpr (HAS_BIAS) fetch_bias(sample, 0);
      for (int j = 0; j < nkb; ++j, ++nb) {
        if constexpr (HAS_BIAS) {
          sts_f1(bias_a + lane * 4, breg[0]);
          sts_f1(bias_a + (lane + 32) * 4, breg[1]);
          __syncwarp();
          if (j + 1 < nkb) fetch_bias(sample, j + 1);
        }
        ATTN_T(s0);
        mbar_wait(&s_full[t], nb & 1);
        tc_fence_after();
        ATTN_T(s1);
        // this thread's 64 scores stay in registers: ONE TMEM read pass, both loads in flight
        uint32_t v[2][32];
        tmem_ld_32x32(t_row + half * 64, v[0]);
        tmem_ld_32x32(t_row + half * 64 + 32, v[1]);
        tmem_ld_wait();
        ATTN_T(s2);
        float mx = softmax_chunk_prep<HAS_BIAS>(v[0], -INFINITY, p.scale_log2e, bias_a);
        mx = softmax_chunk_prep<HAS_BIAS>(v[1], mx, p.scale_log2e, bias_a + 128);
        if constexpr (!HAS_BIAS) mx *= p.scale_log2e;
        // row maximum over both halves; the barrier also orders "partner has read its S columns" before any P write
        const uint32_t buf = (nb & 1) * (16 * 32 * 4);
        sts_f1(my_slot + buf, mx);
        named_bar_sync(bar_id, 64);
        mx = fmaxf(mx, lds_f1(peer_slot + buf));
        ATTN_T(s3);
        // lazy rescale: move the reference maximum only when the block maximum exceeds it by > 8 (log2 units).  Both
        // partners see the same mx / m_used, so they take the same decisions; each rescales its own O columns.
        // tcgen05.ld/st are warp-collective: the TMEM round trip runs for the whole warp as soon as ANY row needs it.
        const bool need = mx > m_used + 8.0f;
        const bool rescale = need && j > 0 && m_used != -INFINITY;
        const float alpha = rescale ? fast_exp2(m_used - mx) : 1.0f;
        if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll
          for (int c = (half ? kOSplit : 0); c < (half ? Cfg::kPad : kOSplit); c += 16) {
            uint32_t o[16];
            tmem_ld_32x16(t_row + 128 + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x16(t_row + 128 + c, o);
          }
        }
        l_sum *= alpha;
        if (need) m_used = mx;
        const float m_eff = m_used == -INFINITY ? 0.f : m_used;  // a fully masked prefix must not produce NaN
        uint64_t sum2 = pack_f2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t pk[16];
          softmax_chunk_exp_reg<HAS_BIAS>(v[c], pk, sum2, m_eff, p.scale_log2e);
          tmem_st_32x16(t_row + half * 32 + c * 16, pk);  // P of keys [64*half + 32c, +32) as bf16 pairs
          if (c == 0) {  // first 32 keys of this half are in TMEM: let the MMA warp start P V on them
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_half[t]);
          }
        }
        {
          float s_lo, s_hi;
          unpack_f2(sum2, s_lo, s_hi);
          l_sum += s_lo + s_hi;
        }
        ATTN_T(s4);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        ATTN_T(s5);
        ATTN_ACC(0, s0, s1);  // wait S
        ATTN_ACC(1, s1, s2);  // TMEM load
        ATTN_ACC(2, s2, s3);  // max + exchange
        ATTN_ACC(3, s3, s4);  // exp + P store issue
        ATTN_ACC(4, s4, s5);  // st wait + fence + arrive
      }
      // ---- item done: O_t / l -> bf16 -> global; each partner writes its own output columns
      sts_f1(my_slot + 2 * (16 * 32 * 4), l_sum);
      mbar_wait(&o_full[t], n & 1);
      tc_fence_after();
      named_bar_sync(bar_id, 64);
      const float inv = 1.0f / (l_sum + lds_f1(peer_slot + 2 * (16 * 32 * 4)));
      const int q = pr * 256 + t * kAttnBM + row;
      __nv_bfloat16* dst;
      if (p.split_tokens > 0) {
        dst = q < p.split_tokens
                  ? p.out_lo + (static_cast<size_t>(sample) * p.split_tokens + q) * p.out_ld
                  : p.out + (static_cast<size_t>(sample) * (p.q_tokens - p.split_tokens) + (q - p.split_tokens)) * p.out_ld;
      } else {
        dst = p.out + (static_cast<size_t>(sample) * p.q_tokens + q) * p.out_ld;
      }
      dst += head * HD;
      auto store8 = [&](const uint32_t* w, __nv_bfloat16* d) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(w[0]) * inv, __uint_as_float(w[1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(w[2]) * inv, __uint_as_float(w[3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(w[4]) * inv, __uint_as_float(w[5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(w[6]) * inv, __uint_as_float(w[7]) * inv);
        *reinterpret_cast<uint4*>(d) = o;
      };
      // O leaves TMEM into registers and the tile's columns go back to the MMA warp BEFORE the (slow, row-strided)
      // global stores, so the next item's QK^T does not wait for them
      if constexpr (HD == 128) {
        uint32_t v[2][32];
        tmem_ld_32x32(t_row + 128 + half * 64, v[0]);
        tmem_ld_32x32(t_row + 128 + half * 64 + 32, v[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[t]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int g = 0; g < 4; ++g) store8(v[c] + g * 8, dst + half * 64 + c * 32 + g * 8);
        }
      } else {  // HD = 72 stored as 80 columns: half 0 -> columns 0..47, half 1 -> 48..71 (72..79 are padding)
        uint32_t v[32];
        tmem_ld_32x32(t_row + 128 + half * 48, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 3; ++g) store8(v + g * 8, dst + half * 48 + g * 8);
        if (half == 0) {
          store8(v + 24, dst + 24);
          uint32_t w[16];
          tmem_ld_32x16(t_row + 128 + 32, w);
          tmem_ld_wait();
          store8(w, dst + 32);
          store8(w + 8, dst + 40);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[t]);
      }
    }
#ifdef ECADK_ATTN_TIMING
    if (lane == 0 && (sw == 0 || sw == 12)) {
      dbg_acc[6] = clock() - t_begin;
      dbg_acc[7] = nb;
      for (int i = 0; i < 8; ++i) g_attn_dbg[blockIdx.x * 32 + (sw == 0 ? 8 : 16) + i] = dbg_acc[i];
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =====================================================================================================
// Second-generation streaming kernel for head dim 72 (PixArt 512 / 1024 px self-attention, sigma cross-attention).
//
// attn_flash_kernel's chain per query tile and key block is  softmax -> P V -> next Q K^T -> softmax: P overwrites S in
// place, so the next S can only be produced after P V has consumed P, and the softmax warps wait for it (~1100-1500
// clocks of a ~3300-3700 clock period, profiles/r1_attention_times.txt).  Here the chain is cut:
//   * a thread keeps its 64 scores in registers, so the S columns are FREE as soon as the tcgen05.ld has landed
//     (`s_free`): the MMA thread issues the NEXT block's Q K^T right then, one whole softmax ahead of its use;
//   * P therefore cannot live over S.  With an 80-column O the two tiles leave 96 TMEM columns = 96 of a block's 128
//     keys as packed bf16 (A operand from TMEM); the last 32 keys of every row go through shared memory as one
//     128B-swizzled K-major slab (A operand from shared memory).  TMEM per tile: S [0,128) O [128,208) P [208,256);
//   * the P buffers are single: a block's P is written only after the previous block's P V has retired (`pv_done`),
//     which by then is one softmax old.
// Each query tile has its OWN MMA-issuing thread (two warps in different sub-partitions): a single thread issuing for
// both tiles in a static order measured ~100 clocks per tcgen05.mma next to four busy softmax warps, i.e. the issue
// loop itself was the period.  K and V run through 3-stage rings (K is consumed one block early).
// Shared-memory traffic per pair of tiles and key block: operands 136 KB + P 16 KB + TMA 40 KB = 1536 clk at 128 B/clk,
// tensor pipe 1328 clk, MUFU 32768 exp = ~1700 clk at the measured 19.6/clk - the exponentials bound it.
// =====================================================================================================
struct AttnFlash2Cfg {
  static constexpr int kStages = 3;
  static constexpr int kPad = 80;
  static constexpr int kQ64 = 0;                           // 256 rows x 128 B
  static constexpr int kQ2 = kQ64 + 256 * 128;             // 256 rows x 32 B
  static constexpr int kKStage = kFlashKB * (128 + 32);    // 20 KB
  static constexpr int kK = kQ2 + 256 * 32;
  static constexpr int kV = kK + kStages * kKStage;
  static constexpr int kP = kV + kStages * kKStage;        // 2 tiles x [128 rows x 128 B] (64 B of every row used)
  static constexpr int kPTile = kAttnBM * 128;
  static constexpr int kBias = kP + 2 * kPTile;            // 16 warps x 64 floats
  static constexpr int kXch = kBias + 8 * kFlashKB * 4;    // row-maximum exchange [2][16][32] + row sums [16][32]
  static constexpr int kBars = kXch + 3 * 16 * 32 * 4;
  static constexpr int kSmemBytes = kBars + 256 + 1024;
  static constexpr uint32_t kBytesQ = 256 * kPad * 2;
  static constexpr uint32_t kBytesKV = kFlashKB * kPad * 2;
  static constexpr int kTmemO = 128, kTmemP = 208;         // column offsets inside a tile's 256 columns
  static constexpr int kKeysTmem = 96;                     // keys of a block whose P goes through TMEM
  static_assert(kP % 1024 == 0, "the P slabs are addressed with 128B-swizzle descriptors");
  static_assert(kSmemBytes > 114 * 1024 && kSmemBytes <= 227 * 1024, "one CTA per SM (it owns all 512 TMEM columns)");
};

constexpr int kFlash2Warps = 19;  // warp 0 TMA, warp 1 MMA of tile 0, warps 2..17 softmax, warp 18 MMA of tile 1
constexpr int kFlash2Threads = kFlash2Warps * 32;

template <bool HAS_BIAS>
__global__ void __launch_bounds__(kFlash2Threads, 1)
attn_flash2_kernel(const __grid_constant__ CUtensorMap tm_q64, const __grid_constant__ CUtensorMap tm_q16,
                   const __grid_constant__ CUtensorMap tm_k64, const __grid_constant__ CUtensorMap tm_k16,
                   const __grid_constant__ CUtensorMap tm_v16, const AttnParams p, const int n_keys,
                   const int num_items) {
  using Cfg = AttnFlash2Cfg;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBars);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;     // [3]
  uint64_t* k_empty = bars + 5;    // [3]
  uint64_t* v_full = bars + 8;     // [3]
  uint64_t* v_empty = bars + 11;   // [3]
  uint64_t* s_full = bars + 14;    // [2] per query tile, one phase per key block
  uint64_t* s_free = bars + 16;    // [2] every softmax warp of the tile holds its scores in registers
  uint64_t* p_full = bars + 18;    // [2]
  uint64_t* pv_done = bars + 20;   // [2] the tile's P V of a block has retired (P buffers and O may be touched)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  // (the shuffle tells ptxas that the value is warp-uniform - see the MMA issuers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nkb = n_keys / kFlashKB;
  const int pairs = p.q_tokens / 256;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);  // the "empty" barriers: one commit per MMA thread
    for (int i = 0; i < S; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 8);
      mbar_init(&p_full[i], 8);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  griddep_launch_dependents();
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int n = 0, nb = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int sh = item / pairs, pr = item - sh * pairs;
        const int smp = sh / p.heads, hd = sh - smp * p.heads;
        const int q_mid = p.q_rowmajor ? hd : 0, kv_mid = p.kv_rowmajor ? hd : 0;
        const int q_row = (p.q_rowmajor ? smp : sh) * p.q_tokens + pr * 256;
        mbar_wait(q_empty, (n & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, Cfg::kBytesQ);
        tma_load_3d(smem + Cfg::kQ64, &tm_q64, q_full, 0, q_mid, q_row);
        tma_load_3d(smem + Cfg::kQ2, &tm_q16, q_full, 64, q_mid, q_row);
        for (int j = 0; j < nkb; ++j, ++nb) {
          const int st = nb % S;
          const uint32_t ph = ((nb / S) & 1) ^ 1;
          const int k_row = (p.kv_rowmajor ? smp : sh) * n_keys + j * kFlashKB;
          uint8_t* kd = smem + Cfg::kK + st * Cfg::kKStage;
          uint8_t* vd = smem + Cfg::kV + st * Cfg::kKStage;
          if (j == (nkb > 2 ? nkb - 3 : 0) && item + static_cast<int>(gridDim.x) < num_items) {
            // the next item's Q can only be LOADED once this item's last Q K^T has retired, with the MMA warps about to
            // wait for it: bring it into L2 a few key blocks early so that the load is an L2 hit
            const int item2 = item + gridDim.x;
            const int sh2 = item2 / pairs, pr2 = item2 - sh2 * pairs;
            const int smp2 = sh2 / p.heads, hd2 = sh2 - smp2 * p.heads;
            const int q_row2 = (p.q_rowmajor ? smp2 : sh2) * p.q_tokens + pr2 * 256;
            tma_prefetch_3d(&tm_q64, 0, p.q_rowmajor ? hd2 : 0, q_row2);
            tma_prefetch_3d(&tm_q16, 64, p.q_rowmajor ? hd2 : 0, q_row2);
          }
          mbar_wait(&k_empty[st], ph);
          mbar_arrive_expect_tx(&k_full[st], Cfg::kBytesKV);
          tma_load_3d(kd, &tm_k64, &k_full[st], 0, kv_mid, k_row);
          tma_load_3d(kd + kFlashKB * 128, &tm_k16, &k_full[st], 64, kv_mid, k_row);
          mbar_wait(&v_empty[st], ph);
          mbar_arrive_expect_tx(&v_full[st], Cfg::kBytesKV);
#pragma unroll
          for (int a = 0; a < Cfg::kPad / 16; ++a)  // five 16-column SW32 atoms: one MN-major operand
            tma_load_3d(vd + a * (kFlashKB * 32), &tm_v16, &v_full[st], a * 16, kv_mid, k_row);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == kFlash2Warps - 1) {
    {
      // ===================== MMA issuers: ONE WARP PER QUERY TILE (warp 1: tile 0, last warp: tile 1) =============
      // The WHOLE warp runs this loop and only the tcgen05 instructions are predicated on one elected lane.  Issued from
      // inside an `if (lane == 0)` branch, every tcgen05.mma cost ~125 clocks of the issuing thread: ptxas cannot prove
      // the operands warp-uniform there and moves each one into a uniform register through an ELECT / R2UR.BROADCAST /
      // BRA.U.ANY loop (7 R2UR per MMA in the SASS) - 26 MMAs x 125 clk was the whole period of a key block, for every
      // attention kernel of this file.  Warp-uniform control flow keeps the descriptors in uniform registers.
      // A single issuing thread shares its sub-partition's issue port with four busy softmax warps and took ~100 clocks
      // per tcgen05.mma (26 per key block = the whole period); the tiles are independent, so each gets its own thread
      // in a different sub-partition, and there is no issue order between the tiles left to get wrong.  K / V / Q go
      // back to the producer when BOTH threads have committed (those barriers count 2).
      const int t = warp == 1 ? 0 : 1;
      constexpr uint32_t idesc_s = make_idesc_bf16(kAttnBM, kFlashKB);
      constexpr uint32_t idesc_o = make_idesc_bf16(kAttnBM, 80, 0, 1);
      constexpr uint32_t kStageDesc = Cfg::kKStage >> 4;  // descriptor start-address units (16 B) per ring stage
      const uint32_t sbase = smem_u32(smem);
      const int my_items = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / gridDim.x;
      const int G = my_items * nkb;  // key blocks this CTA runs, over all of its items
      const uint32_t s_tmem = tmem + 256 * t;
      const uint32_t o_tmem = s_tmem + Cfg::kTmemO;
      const uint32_t p_tmem = s_tmem + Cfg::kTmemP;
      // loop-invariant descriptors (stage 0; a stage adds kStageDesc to the start-address field)
      const uint64_t dq = make_smem_desc(sbase + Cfg::kQ64 + t * (kAttnBM * 128), 16, 1024, kLayoutSW128);
      const uint64_t dq2 = make_smem_desc(sbase + Cfg::kQ2 + t * (kAttnBM * 32), 16, 256, kLayoutSW32);
      const uint64_t dk0 = make_smem_desc(sbase + Cfg::kK, 16, 1024, kLayoutSW128);
      const uint64_t dk20 = make_smem_desc(sbase + Cfg::kK + kFlashKB * 128, 16, 256, kLayoutSW32);
      const uint64_t dv0 = make_smem_desc(sbase + Cfg::kV, kFlashKB * 32, 256, kLayoutSW32);
      const uint64_t dp = make_smem_desc(sbase + Cfg::kP + t * Cfg::kPTile, 16, 1024, kLayoutSW128);
      // operands of block g are in shared memory: K(g), and Q of its item when g opens one
      auto operands_ready = [&](int g) {
        if (g % nkb == 0) mbar_wait(q_full, (g / nkb) & 1);
        mbar_wait(&k_full[g % S], (g / S) & 1);
        tc_fence_after();
      };
      auto issue_qk = [&](int g) {
        const int st = g % S;
        const uint64_t dk = dk0 + st * kStageDesc, dk2 = dk20 + st * kStageDesc;
        const bool last = g % nkb == nkb - 1;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(s_tmem, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          umma_bf16_ss(s_tmem, dq2, dk2, idesc_s, 1);
          umma_commit(&s_full[t]);
          umma_commit(&k_empty[st]);       // this tile is done with K(g) ...
          if (last) umma_commit(q_empty);  // ... and with Q after an item's last block
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g) {
        const int st = g % S;
        const uint64_t dv = dv0 + st * kStageDesc;
        const uint32_t acc0 = (g % nkb != 0) ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            if (ks < Cfg::kKeysTmem / 16) {
              umma_bf16_ts(o_tmem, p_tmem + ks * 8, dv + ks * 32, idesc_o, ks != 0 ? 1u : acc0);
            } else {
              umma_bf16_ss(o_tmem, dp + 2 * (ks - Cfg::kKeysTmem / 16), dv + ks * 32, idesc_o, 1);
            }
          }
          umma_commit(&pv_done[t]);
          umma_commit(&v_empty[st]);
        }
        __syncwarp();
      };
#ifdef ECADK_ATTN_TIMING
      unsigned int dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
      operands_ready(0);
      issue_qk(0);
      for (int g = 0; g < G; ++g) {
        const uint32_t par = g & 1;
        ATTN_T(m0);
        // the tile holds S(g) in registers -> its next Q K^T, one whole softmax ahead of its use
        mbar_wait(&s_free[t], par);
        ATTN_T(m1);
        if (g + 1 < G) {
          operands_ready(g + 1);
          ATTN_T(m1b);
          ATTN_ACC(1, m1, m1b);  // wait K (+ Q)
          issue_qk(g + 1);
        }
        ATTN_T(m2);
        mbar_wait(&v_full[g % S], (g / S) & 1);
        ATTN_T(m3);
        mbar_wait(&p_full[t], par);
        tc_fence_after();
        ATTN_T(m4);
        issue_pv(g);
        ATTN_T(m5);
        ATTN_ACC(0, m0, m1);  // wait s_free
        ATTN_ACC(2, m1, m2);  // wait K + issue QK
        ATTN_ACC(3, m2, m3);  // wait V
        ATTN_ACC(4, m3, m4);  // wait P
        ATTN_ACC(5, m4, m5);  // issue PV
      }
#ifdef ECADK_ATTN_TIMING
      if (t == 0 && lane == 0) {  // (slots 6 / 7 are written by the softmax warps: their turn waits)
        for (int i = 0; i < 6; ++i) g_attn_dbg[blockIdx.x * 32 + i] = dbg_acc[i];
      }
#endif
    }
    __syncwarp();
  } else {
    // ===================== softmax + correction + epilogue =====================
    const int sw = warp - 2;
    const int t = sw >> 3;            // query tile
    const int half = (sw >> 2) & 1;   // which 64 keys of a block / which output columns this thread owns
    const int quarter = warp & 3;     // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t t_row = tmem + 256 * t + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t bias_a = smem_u32(smem + Cfg::kBias) + sw * (kFlashKB / 2) * 4;
    const uint32_t xch = smem_u32(smem + Cfg::kXch);
    const uint32_t my_slot = xch + (sw * 32 + lane) * 4;
    const uint32_t peer_slot = xch + ((sw ^ 4) * 32 + lane) * 4;
    const int bar_id = 1 + t * 4 + quarter;  // named barrier of the two partner warps
    // this row's 128-byte line of the tile's P slab; its 16-byte chunks are XOR-swizzled with the row (SW128)
    const uint32_t p_line = smem_u32(smem + Cfg::kP) + t * Cfg::kPTile + (row >> 3) * 1024 + (row & 7) * 128;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr int kOSplit = 48;  // O columns [0, 48) -> half 0, [48, 80) -> half 1
    float breg[2];
    auto fetch_bias = [&](int sample, int j) {
      const float* b = p.bias + static_cast<size_t>(sample) * n_keys + j * kFlashKB + half * 64;
      breg[0] = __ldg(b + lane) * kLog2e;
      breg[1] = __ldg(b + lane + 32) * kLog2e;
    };
    int nb = 0;
#ifdef ECADK_ATTN_TIMING
    unsigned int dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned int t_begin = clock();
#endif
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int sh = item / pairs, pr = item - sh * pairs;
      const int sample = sh / p.heads, head = sh - sample * p.heads;
      float m_used = -INFINITY, l_sum = 0.f;  // l_sum: this thread's 64-key share of the row sum
      if constexpr (HAS_BIAS) fetch_bias(sample, 0);
      for (int j = 0; j < nkb; ++j, ++nb) {
        if constexpr (HAS_BIAS) {
          sts_f1(bias_a + lane * 4, breg[0]);
          sts_f1(bias_a + (lane + 32) * 4, breg[1]);
          __syncwarp();
          if (j + 1 < nkb) fetch_bias(sample, j + 1);
        }
        ATTN_T(s0);
        mbar_wait(&s_full[t], nb & 1);
        tc_fence_after();
        ATTN_T(s1);
        uint32_t v[2][32];
        tmem_ld_32x32(t_row + half * 64, v[0]);
        tmem_ld_32x32(t_row + half * 64 + 32, v[1]);
        tmem_ld_wait();
        ATTN_T(s2);
        // the scores are in registers: the tile's S columns may take the next block's Q K^T
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        float mx = softmax_chunk_prep<HAS_BIAS>(v[0], -INFINITY, p.scale_log2e, bias_a);
        mx = softmax_chunk_prep<HAS_BIAS>(v[1], mx, p.scale_log2e, bias_a + 128);
        if constexpr (!HAS_BIAS) mx *= p.scale_log2e;
        const uint32_t buf = (nb & 1) * (16 * 32 * 4);
        sts_f1(my_slot + buf, mx);
        named_bar_sync(bar_id, 64);
        mx = fmaxf(mx, lds_f1(peer_slot + buf));
        ATTN_T(s3);
        // lazy rescale (see attn_flash_kernel); O may only be touched once the previous block's P V has retired
        const bool need = mx > m_used + 8.0f;
        const bool rescale = need && j > 0 && m_used != -INFINITY;
        const float alpha = rescale ? fast_exp2(m_used - mx) : 1.0f;
        if (__any_sync(0xffffffffu, rescale)) {
          mbar_wait(&pv_done[t], (nb - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = (half ? kOSplit : 0); c < (half ? Cfg::kPad : kOSplit); c += 16) {
            uint32_t o[16];
            tmem_ld_32x16(t_row + Cfg::kTmemO + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x16(t_row + Cfg::kTmemO + c, o);
          }
        }
        l_sum *= alpha;
        if (need) m_used = mx;
        const float m_eff = m_used == -INFINITY ? 0.f : m_used;  // a fully masked prefix must not produce NaN
        uint64_t sum2 = pack_f2(0.f, 0.f);
        uint32_t pk[2][16];
        // (measured and rejected: the two tiles taking strict TURNS on the exponentials - a tile's two warps per
        // sub-partition alone reach only ~13 exp/clk/SM (dependency-bound, 1300 clk per block) against ~21 with all four
        // warps queueing on the MUFU: 1563 -> 1902 us at the 4096-token shape)
        softmax_chunk_exp_reg<HAS_BIAS>(v[0], pk[0], sum2, m_eff, p.scale_log2e);
        softmax_chunk_exp_reg<HAS_BIAS>(v[1], pk[1], sum2, m_eff, p.scale_log2e);
        {
          float s_lo, s_hi;
          unpack_f2(sum2, s_lo, s_hi);
          l_sum += s_lo + s_hi;
        }
        // the single P buffers: the previous block's P V (issued one softmax ago) must have read them
        ATTN_T(s4);
        if (nb > 0) {
          mbar_wait(&pv_done[t], (nb - 1) & 1);
          tc_fence_after();
        }
        ATTN_T(s5);
        if (half == 0) {  // keys 0..63 -> packed TMEM columns [0, 32)
          tmem_st_32x16(t_row + Cfg::kTmemP, pk[0]);
          tmem_st_32x16(t_row + Cfg::kTmemP + 16, pk[1]);
        } else {          // keys 64..95 -> TMEM columns [32, 48); keys 96..127 -> the shared-memory slab
          tmem_st_32x16(t_row + Cfg::kTmemP + 32, pk[0]);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            sts_u4(p_line + ((c ^ (row & 7)) << 4), pk[1][4 * c], pk[1][4 * c + 1], pk[1][4 * c + 2], pk[1][4 * c + 3]);
          fence_proxy_async_smem();
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        ATTN_T(s6);
        ATTN_ACC(0, s0, s1);  // wait S
        ATTN_ACC(1, s1, s2);  // TMEM load
        ATTN_ACC(2, s2, s3);  // s_free arrive + max + exchange
        ATTN_ACC(3, s3, s4);  // (rescale) + exp
        ATTN_ACC(4, s4, s5);  // wait pv_done
        ATTN_ACC(5, s5, s6);  // P store + fences + arrive
      }
      // ---- item done: O_t / l -> bf16 -> global; each partner writes its own output columns.  The next item's first
      // P V (which overwrites O_t) is only issued behind this warp's next p_full arrival, i.e. after this read-out.
      sts_f1(my_slot + 2 * (16 * 32 * 4), l_sum);
      mbar_wait(&pv_done[t], (nb - 1) & 1);
      tc_fence_after();
      named_bar_sync(bar_id, 64);
      const float inv = 1.0f / (l_sum + lds_f1(peer_slot + 2 * (16 * 32 * 4)));
      const int q = pr * 256 + t * kAttnBM + row;
      __nv_bfloat16* dst = p.out + (static_cast<size_t>(sample) * p.q_tokens + q) * p.out_ld + head * 72;
      auto store8 = [&](const uint32_t* w, __nv_bfloat16* d) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(w[0]) * inv, __uint_as_float(w[1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(w[2]) * inv, __uint_as_float(w[3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(w[4]) * inv, __uint_as_float(w[5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(w[6]) * inv, __uint_as_float(w[7]) * inv);
        *reinterpret_cast<uint4*>(d) = o;
      };
      // 72 real columns stored as 80: half 0 -> columns 0..47, half 1 -> 48..71 (72..79 are padding)
      uint32_t w0[32];
      tmem_ld_32x32(t_row + Cfg::kTmemO + half * 48, w0);
      tmem_ld_wait();
#pragma unroll
      for (int g8 = 0; g8 < 3; ++g8) store8(w0 + g8 * 8, dst + half * 48 + g8 * 8);
      if (half == 0) {
        store8(w0 + 24, dst + 24);
        uint32_t w1[16];
        tmem_ld_32x16(t_row + Cfg::kTmemO + 32, w1);
        tmem_ld_wait();
        store8(w1, dst + 32);
        store8(w1 + 8, dst + 40);
      }
    }
#ifdef ECADK_ATTN_TIMING
    if (lane == 0 && (sw == 0 || sw == 12 || sw == 4)) {
      dbg_acc[6] = clock() - t_begin;
      dbg_acc[7] = nb;
      for (int i = 0; i < 8; ++i) g_attn_dbg[blockIdx.x * 32 + (sw == 0 ? 8 : (sw == 12 ? 16 : 24)) + i] = dbg_acc[i];
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}


}  // namespace ecadk
