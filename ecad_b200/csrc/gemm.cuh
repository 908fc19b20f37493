// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T (+ bias, + fused epilogue)
//
//   warp 0 (1 thread)  : TMA producer   - cp.async.bulk.tensor A/W tiles into a STAGES-deep smem ring
//   warp 1 (1 thread)  : MMA issuer     - tcgen05.mma (M=128, N=BN, K=16) into one of two TMEM accumulators
//   warps 2..5         : epilogue       - tcgen05.ld the finished accumulator, fused epilogue, global stores,
//                                         overlapped with the next tile's main loop (TMEM double buffering)
//
// Both operands are K-major (PyTorch nn.Linear weight is [out,in] = [N,K]), staged with the 128-byte TMA/UMMA swizzle.
// Epilogues cover every GEMM on the PixArt hot path (K3,K5,K6,K8,K10,K11,K15).
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace ecadk {

enum GemmEpilogue : int {
  EPI_BIAS = 0,            // out = bf16(acc + bias)                                   row-major [M, ldo]
  EPI_BIAS_GELU = 1,       // out = bf16(gelu_tanh(acc + bias))                        row-major [M, ldo]
  EPI_GATED_RESIDUAL = 2,  // o = acc + bias; cache = bf16(o); x += gate * o; (xb = bf16(x))
  EPI_HEADMAJOR = 3,       // out[part][sample][head][token][head_pad] = bf16(acc + bias)   (Q/K/V scatter)
};

struct GemmParams {
  int M, N, K;
  const float* bias;  // [N] fp32 (may be null)
  // EPI_BIAS / EPI_BIAS_GELU
  __nv_bfloat16* out;
  int ldo;
  // EPI_GATED_RESIDUAL
  float* x;                 // [M,N] fp32 residual stream, updated in place
  __nv_bfloat16* xb;        // optional bf16 shadow of the updated stream (feeds the next projection), or null
  __nv_bfloat16* cache;     // [M,N] un-gated sub-block output (the reference's cached_*_output)
  const float* gate_table;  // [N] row of the block's scale_shift_table, or null for "no gate" (attn2)
  const float* gate_temb;   // [samples, temb_stride] pointer already offset to the gate chunk
  int temb_stride;
  int tokens;  // rows per sample (maps a row to its sample for gate / head-major addressing)
  // EPI_HEADMAJOR
  __nv_bfloat16* hm_out[3];
  int heads, head_dim, head_pad, tokens_pad;
};

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kStageA = kGemmBM * kGemmBK * 2;  // 16 KB
  static constexpr int kStageB = BN * kGemmBK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (BN <= 128) ? 6 : (BN <= 192 ? 5 : 4);
  static constexpr int kAccStride = 256;  // TMEM column offset between the two accumulators
  static constexpr int kTmemCols = 512;
  static constexpr int kSmemBytes = kStages * kStage + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStage);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (p.M + kGemmBM - 1) / kGemmBM;
  const int num_n = p.N / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = p.K / kGemmBK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {  // whole warp: TMEM allocation
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * kGemmBM;
        const int n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStage;
          uint8_t* sb = sa + Cfg::kStageA;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStage);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kGemmBK, m0);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kGemmBK, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_bf16(kGemmBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStage);
          const uint32_t sb = sa + Cfg::kStageA;
          const uint64_t da = make_smem_desc(sa, 0, 1024, kLayoutSW128);
          const uint64_t db = make_smem_desc(sb, 0, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
            umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // smem slot is free once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * kGemmBM;
      const int n0 = (tile % num_n) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(quarter * 32) << 16);
      int sample = 0, tok = 0;
      if constexpr (EPI == EPI_GATED_RESIDUAL || EPI == EPI_HEADMAJOR) {
        sample = row / p.tokens;
        tok = row - sample * p.tokens;
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
            v[j + 0] = __float_as_uint(__uint_as_float(v[j + 0]) + b.x);
            v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + b.y);
            v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + b.z);
            v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + b.w);
          }
        }
        if (row_ok) {
          if constexpr (EPI == EPI_BIAS || EPI == EPI_BIAS_GELU) {
            __nv_bfloat16* dst = p.out + static_cast<size_t>(row) * p.ldo + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                f[e] = __uint_as_float(v[j + e]);
                if constexpr (EPI == EPI_BIAS_GELU) f[e] = gelu_tanh(f[e]);
              }
              uint4 o;
              o.x = pack_bf16x2(f[0], f[1]);
              o.y = pack_bf16x2(f[2], f[3]);
              o.z = pack_bf16x2(f[4], f[5]);
              o.w = pack_bf16x2(f[6], f[7]);
              *reinterpret_cast<uint4*>(dst + j) = o;
            }
          } else if constexpr (EPI == EPI_GATED_RESIDUAL) {
            const size_t off = static_cast<size_t>(row) * p.N + col0;
            float* xrow = p.x + off;
            __nv_bfloat16* crow = p.cache + off;
            const float* gt = p.gate_table ? p.gate_table + col0 : nullptr;
            const float* ge = p.gate_table ? p.gate_temb + static_cast<size_t>(sample) * p.temb_stride + col0 : nullptr;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float o[8], xn[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[j + e]);
              uint4 cv;
              cv.x = pack_bf16x2(o[0], o[1]);
              cv.y = pack_bf16x2(o[2], o[3]);
              cv.z = pack_bf16x2(o[4], o[5]);
              cv.w = pack_bf16x2(o[6], o[7]);
              *reinterpret_cast<uint4*>(crow + j) = cv;
              const float4 x0 = *reinterpret_cast<const float4*>(xrow + j);
              const float4 x1 = *reinterpret_cast<const float4*>(xrow + j + 4);
              float g[8];
              if (gt != nullptr) {
                const float4 a0 = __ldg(reinterpret_cast<const float4*>(gt + j));
                const float4 a1 = __ldg(reinterpret_cast<const float4*>(gt + j + 4));
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(ge + j));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(ge + j + 4));
                g[0] = a0.x + b0.x; g[1] = a0.y + b0.y; g[2] = a0.z + b0.z; g[3] = a0.w + b0.w;
                g[4] = a1.x + b1.x; g[5] = a1.y + b1.y; g[6] = a1.z + b1.z; g[7] = a1.w + b1.w;
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) g[e] = 1.0f;
              }
              xn[0] = fmaf(g[0], o[0], x0.x); xn[1] = fmaf(g[1], o[1], x0.y);
              xn[2] = fmaf(g[2], o[2], x0.z); xn[3] = fmaf(g[3], o[3], x0.w);
              xn[4] = fmaf(g[4], o[4], x1.x); xn[5] = fmaf(g[5], o[5], x1.y);
              xn[6] = fmaf(g[6], o[6], x1.z); xn[7] = fmaf(g[7], o[7], x1.w);
              *reinterpret_cast<float4*>(xrow + j) = make_float4(xn[0], xn[1], xn[2], xn[3]);
              *reinterpret_cast<float4*>(xrow + j + 4) = make_float4(xn[4], xn[5], xn[6], xn[7]);
              if (p.xb != nullptr) {
                uint4 xv;
                xv.x = pack_bf16x2(xn[0], xn[1]);
                xv.y = pack_bf16x2(xn[2], xn[3]);
                xv.z = pack_bf16x2(xn[4], xn[5]);
                xv.w = pack_bf16x2(xn[6], xn[7]);
                *reinterpret_cast<uint4*>(p.xb + off + j) = xv;
              }
            }
          } else {  // EPI_HEADMAJOR
            const int part_cols = p.heads * p.head_dim;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const int col = col0 + j;
              const int part = col / part_cols;
              const int rem = col - part * part_cols;
              const int h = rem / p.head_dim;
              const int e0 = rem - h * p.head_dim;
              __nv_bfloat16* dst = p.hm_out[part] +
                                   ((static_cast<size_t>(sample) * p.heads + h) * p.tokens_pad + tok) * p.head_pad + e0;
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]));
              o.y = pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              o.z = pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
              o.w = pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
              *reinterpret_cast<uint4*>(dst) = o;
            }
          }
        }
      }
      // accumulator drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace ecadk
