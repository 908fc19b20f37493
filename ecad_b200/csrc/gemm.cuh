// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T (+ bias, + fused epilogue)
//
//   warp 0 (1 thread)  : TMA producer   - cp.async.bulk.tensor A/W tiles into a STAGES-deep smem ring
//   warp 1 (1 thread)  : MMA issuer     - tcgen05.mma (M=128, N=BN, K=16) into one of two TMEM accumulators
//   warps 2..9         : epilogue       - tcgen05.ld the finished accumulator, fused epilogue, global stores,
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
  EPI_UNPATCHIFY = 4,      // fp32 out[s][c][2i+p][2j+q] = acc + bias for the first unp_cols columns (final layer)
  EPI_BIAS_F32 = 5,        // fp32 out[row, col] = acc + bias for col < f32_cols, row pitch ldo (embedders, FLUX head)
  EPI_BIAS_DUAL = 6,       // o = acc + bias; out = bf16(o) (optional: the pre-activation cache); out2 = bf16(gelu_tanh(o))
  EPI_CONV = 7,            // out = bf16(acc + bias (+ residual)) on interior pixels of a zero-bordered NHWC image, 0 on the border
};

struct GemmParams {
  int M, N, K;
  const float* bias;  // [N] fp32 (may be null)
  // A operand split along K: k-blocks [0, kb_split) come from the first tensor map, the rest from a second matrix
  // (FLUX single-stream proj_out reads [attn | GELU(mlp)] from two buffers instead of a concatenated copy)
  int kb_split;
  const void* a2;  // host side only: second A matrix [M, K - k1] (row pitch K - k1), or null
  int k1;          // host side only: columns taken from the first A matrix
  // EPI_BIAS / EPI_BIAS_GELU / EPI_BIAS_DUAL
  __nv_bfloat16* out;
  int ldo;
  __nv_bfloat16* out2;  // EPI_BIAS_DUAL: GELU(tanh) of the same values
  int ldo2;
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
  int hm_tok_off;  // token offset inside the head-major sequence (FLUX: image tokens follow the text tokens)
  // EPI_BIAS_F32
  float* f32_out;
  int f32_cols;
  // Implicit-GEMM convolution (VAE decoder): A is a zero-bordered NHWC image [batch, H+2, W+2, C] seen as a matrix
  // [M = batch*(H+2)*(W+2), C]; k-block kb covers tap kb / conv_cblocks (ky*3 + kx) and channels 64*(kb % conv_cblocks),
  // its A rows are the output rows shifted by (ky-1)*conv_pitch + (kx-1) - a plain 2-D TMA load at a shifted row
  // coordinate (rows outside the matrix are zero-filled).  W is [N, taps*C], tap-major.  conv_taps <= 1: ordinary GEMM.
  int conv_taps, conv_cblocks, conv_pitch;
  int conv_off[9];  // row offset of every tap (3x3: (ky-1)*conv_pitch + (kx-1))
  // EPI_CONV: border mask from (conv_h, conv_w); optional bf16 residual [M, ldo2] in out2; only chunks that start
  // below conv_cols are stored (conv_out: 3 real output channels in a 32-column buffer)
  int conv_h, conv_w, conv_cols;
  // conv_up != 0: the convolution of a nearest-2x-upsampled image, computed on the ORIGINAL image - this launch produces
  // the output pixels of parity (conv_up_a, conv_up_b): input pixel (y, x) -> output pixel (2y-1+a, 2x-1+b) of the
  // bordered [batch, 2h+2, 2w+2, N] output (bordered coordinates; see ecadk_conv_up2x_nhwc)
  int conv_up, conv_up_a, conv_up_b;
  // EPI_UNPATCHIFY (column o = (p*2+q)*C + c of token n = i*Wp + j of sample s)
  float* unp_out;
  int unp_wp, unp_hp, unp_c, unp_cols;
};

// timing experiments only (results are wrong): -DECADK_EPI_DBG=1 no residual-stream loads, 2 no stores, 3 neither
#ifndef ECADK_EPI_DBG
#define ECADK_EPI_DBG 0
#endif
constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kEpiWarps = 8;
constexpr int kEpiPitch = 36;                           // floats; see epilogue_chunk
constexpr int kEpiStageBytes = kEpiWarps * 32 * kEpiPitch * 4;  // one 32x32 fp32 transpose tile per epilogue warp
// Gated-residual epilogue with TMA stores: per warp one 128B-swizzled fp32 tile (transpose tile AND staging of the updated
// residual chunk) + two 64B-swizzled bf16 tiles (cache, bf16 shadow) = 8 KB
#ifndef ECADK_EPI_TMA_STORE
#define ECADK_EPI_TMA_STORE 1
#endif
constexpr int kEpiTmaWarpBytes = 8192;
constexpr int kEpiTmaStageBytes = kEpiWarps * kEpiTmaWarpBytes;
constexpr int kSmemLimit = 227 * 1024;
__host__ __device__ constexpr int epi_stage_bytes(int epi) {
  return (epi == 2 /*EPI_GATED_RESIDUAL*/ && ECADK_EPI_TMA_STORE) ? kEpiTmaStageBytes : kEpiStageBytes;
}
__host__ __device__ constexpr int fit_stages(int want, int stage_bytes, int epi) {
  const int room = (kSmemLimit - epi_stage_bytes(epi) - 1024 - 256) / stage_bytes;
  return want < room ? want : room;
}

// TMA coordinates (column, row) of the A tile of k-block kb for output rows starting at m0 (see GemmParams::conv_taps)
__device__ __forceinline__ void a_tile_coords(const GemmParams& p, const int kb, const int m0, int& col, int& row) {
  col = kb * kGemmBK;
  row = m0;
  if (p.conv_taps > 1) {
    const int tap = kb / p.conv_cblocks;
    col = (kb - tap * p.conv_cblocks) * kGemmBK;
    row = m0 + p.conv_off[tap];
  }
}

template <int BN, int EPI = 0>
struct GemmCfg {
  static constexpr int kStageA = kGemmBM * kGemmBK * 2;  // 16 KB
  static constexpr int kStageB = BN * kGemmBK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kEpiBytes = epi_stage_bytes(EPI);
  static constexpr int kStages = fit_stages((BN <= 128) ? 5 : (BN <= 192 ? 4 : 3), kStage, EPI);
  static constexpr int kAccStride = 256;  // TMEM column offset between the two accumulators
  static constexpr int kTmemCols = 512;
  static constexpr int kSmemBytes = kStages * kStage + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(kStages >= 3, "GEMM pipeline depth");
};


// Epilogue of one 128 x BN accumulator tile, executed by 8 warps: warp e covers TMEM lane quarter (e & 3) - the
// rows 32*(e&3) .. +31 of the tile - and every second 32-column chunk (parity e >> 2).
//
// tcgen05.ld hands each thread one ROW (32 consecutive columns); storing that way makes every global access touch 32
// different rows.  Each chunk is therefore transposed through a padded smem tile (pitch 36 floats: conflict-free for
// 128-bit accesses in both directions) so that one warp instruction covers 4 rows x 32 columns = 4 full 128-byte
// lines of the fp32 stream; bias / gate vectors are loaded once per chunk, row -> (sample, token) maps once per tile,
// and the residual-stream loads of a chunk are issued before its transpose so their latency overlaps it.
template <int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const uint32_t t_row, float* stage, const int lane,
                                              const int parity, const int row_base, const int n0, const int n_chunks) {
  const int cg = lane & 7, rs = lane >> 3;
  const uint32_t stage_s = smem_u32(stage);  // explicit shared-space address of this warp's transpose tile
  // per-tile row bookkeeping for the 8 rows this lane stores (r = 4*i + rs)
  int row_off[8];  // EPI_HEADMAJOR: element offset of (sample, token) inside one head-major tensor, head 0
  int sample0 = 0;
  if constexpr (EPI == EPI_GATED_RESIDUAL) sample0 = row_base / p.tokens;  // tokens % 32 == 0: one sample per chunk
  uint32_t interior = 0;  // EPI_CONV: bit i = row i*4 + rs of this lane is an interior pixel of its image
  if constexpr (EPI == EPI_CONV) {
    const int plane = (p.conv_h + 2) * (p.conv_w + 2);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row_base + i * 4 + rs;
      const int bimg = row / plane;
      const int r = row - bimg * plane;
      const int y = r / (p.conv_w + 2), x = r - y * (p.conv_w + 2);
      if (y >= 1 && y <= p.conv_h && x >= 1 && x <= p.conv_w) interior |= 1u << i;
      // the row this lane WRITES: itself, or its parity's pixel of the upsampled output
      row_off[i] = p.conv_up ? (bimg * (2 * p.conv_h + 2) + (2 * y - 1 + p.conv_up_a)) * (2 * p.conv_w + 2) +
                                   (2 * x - 1 + p.conv_up_b)
                             : row;
    }
  }
  if constexpr (EPI == EPI_HEADMAJOR) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row_base + i * 4 + rs;
      const int sample = row / p.tokens;
      const int tok = row - sample * p.tokens;
      row_off[i] = (sample * p.heads * p.tokens_pad + tok + p.hm_tok_off) * p.head_pad;
    }
  }
  // residual-stream prefetch: the x values of this warp's NEXT chunk (c+2) are requested before chunk c is processed.
  // The two register sets ping-pong (the chunk loop is unrolled by two): copying "next" into "current" at the end of
  // an iteration would make the warp wait for the prefetch right there and expose the whole DRAM latency per chunk.
  // The chunk's bias / gate vectors travel with it: loaded inside process() they sat on the critical path of every
  // chunk (ncu source page, round 1: 52 % of the kernel's stall samples were long-scoreboard waits, the top one the
  // FADD that combines the two gate loads).
  struct Vecs {
    float4 b4, g4, t4;
  };
  auto load_vecs = [&](int c, Vecs& v) {
    const int col = n0 + c * 32 + cg * 4;
    v.b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    v.g4 = make_float4(1.f, 1.f, 1.f, 1.f);
    v.t4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias != nullptr) v.b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    if constexpr (EPI == EPI_GATED_RESIDUAL) {
      if (p.gate_table != nullptr || p.gate_temb != nullptr) {
        // the two halves of the gate are summed in process(): an FADD here would wait for both loads on the spot
        v.g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        v.t4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.gate_table != nullptr) v.g4 = __ldg(reinterpret_cast<const float4*>(p.gate_table + col));
        if (p.gate_temb != nullptr)
          v.t4 = __ldg(reinterpret_cast<const float4*>(p.gate_temb + static_cast<size_t>(sample0) * p.temb_stride + col));
      }
    }
  };
  auto load_x = [&](int c, float4 (&dst)[8]) {
    const int col = n0 + c * 32 + cg * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row_base + i * 4 + rs;
      dst[i] = (row < p.M && !(ECADK_EPI_DBG & 1)) ? *reinterpret_cast<const float4*>(p.x + static_cast<size_t>(row) * p.N + col)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto process = [&](const int c, const float4 (&xin)[8], const Vecs& vec, const uint32_t (&v)[32]) {
    const int col = n0 + c * 32 + cg * 4;
    const float4 b4 = vec.b4;
    uint2 conv_res[8];  // EPI_CONV: the residual values of this lane's 8 rows, requested before the transpose (the
                        // residual never aliases `out`; loaded one by one between the stores they sat on the critical path)
    if constexpr (EPI == EPI_CONV) {
      if (p.out2 != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = row_base + i * 4 + rs;
          conv_res[i] = row < p.M ? __ldg(reinterpret_cast<const uint2*>(p.out2 + static_cast<size_t>(row) * p.ldo2 + col))
                                  : make_uint2(0u, 0u);
        }
      }
    }
    const float4 g4 = make_float4(vec.g4.x + vec.t4.x, vec.g4.y + vec.t4.y, vec.g4.z + vec.t4.z, vec.g4.w + vec.t4.w);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sts_u4(stage_s + (lane * kEpiPitch + 4 * j) * 4, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    __syncwarp();
    int hm_off = 0;
    __nv_bfloat16* hm_base = nullptr;
    if constexpr (EPI == EPI_HEADMAJOR) {
      const int part_cols = p.heads * p.head_dim;
      const int part = col / part_cols;
      const int rem = col - part * part_cols;
      const int head = rem / p.head_dim;
      hm_off = head * p.tokens_pad * p.head_pad + (rem - head * p.head_dim);
      hm_base = part == 0 ? p.hm_out[0] : (part == 1 ? p.hm_out[1] : p.hm_out[2]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + rs;
      const int row = row_base + r;
      float4 o = lds_f4(stage_s + (r * kEpiPitch + cg * 4) * 4);
      o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
      if (row < p.M) {
        if constexpr (EPI == EPI_BIAS || EPI == EPI_BIAS_GELU) {
          if constexpr (EPI == EPI_BIAS_GELU) {
            o.x = gelu_tanh(o.x); o.y = gelu_tanh(o.y); o.z = gelu_tanh(o.z); o.w = gelu_tanh(o.w);
          }
          uint2 w;
          w.x = pack_bf16x2(o.x, o.y);
          w.y = pack_bf16x2(o.z, o.w);
          *reinterpret_cast<uint2*>(p.out + static_cast<size_t>(row) * p.ldo + col) = w;
        } else if constexpr (EPI == EPI_BIAS_DUAL) {
          uint2 w;
          if (p.out != nullptr) {
            w.x = pack_bf16x2(o.x, o.y);
            w.y = pack_bf16x2(o.z, o.w);
            *reinterpret_cast<uint2*>(p.out + static_cast<size_t>(row) * p.ldo + col) = w;
          }
          w.x = pack_bf16x2(gelu_tanh(o.x), gelu_tanh(o.y));
          w.y = pack_bf16x2(gelu_tanh(o.z), gelu_tanh(o.w));
          *reinterpret_cast<uint2*>(p.out2 + static_cast<size_t>(row) * p.ldo2 + col) = w;
        } else if constexpr (EPI == EPI_GATED_RESIDUAL) {
          const size_t off = static_cast<size_t>(row) * p.N + col;
          if ((ECADK_EPI_DBG & 2) && o.x != 12345.678f) continue;
          if (p.cache != nullptr) {  // null: the caller knows this slot is overwritten before anyone reads it
            uint2 cv;
            cv.x = pack_bf16x2(o.x, o.y);
            cv.y = pack_bf16x2(o.z, o.w);
            *reinterpret_cast<uint2*>(p.cache + off) = cv;
          }
          float4 x = xin[i];
          x.x = fmaf(g4.x, o.x, x.x); x.y = fmaf(g4.y, o.y, x.y);
          x.z = fmaf(g4.z, o.z, x.z); x.w = fmaf(g4.w, o.w, x.w);
          *reinterpret_cast<float4*>(p.x + off) = x;
          if (p.xb != nullptr) {
            uint2 xv;
            xv.x = pack_bf16x2(x.x, x.y);
            xv.y = pack_bf16x2(x.z, x.w);
            *reinterpret_cast<uint2*>(p.xb + off) = xv;
          }
        } else if constexpr (EPI == EPI_HEADMAJOR) {
          uint2 w;
          w.x = pack_bf16x2(o.x, o.y);
          w.y = pack_bf16x2(o.z, o.w);
          *reinterpret_cast<uint2*>(hm_base + row_off[i] + hm_off) = w;
        } else if constexpr (EPI == EPI_CONV) {
          if (p.out2 != nullptr) {  // residual branch of the ResNet block, same layout
            const uint2 rv = conv_res[i];
            const float2 r01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rv.x));
            const float2 r23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rv.y));
            o.x += r01.x; o.y += r01.y; o.z += r23.x; o.w += r23.y;
          }
          uint2 w = make_uint2(0u, 0u);  // border pixels stay zero: they are the next convolution's padding
          const bool inside = (interior >> i) & 1u;
          if (inside) {
            w.x = pack_bf16x2(o.x, o.y);
            w.y = pack_bf16x2(o.z, o.w);
          }
          // (upsampling form: border rows of the input map to nothing - the output's border is zeroed by the caller)
          if (inside || !p.conv_up)
            *reinterpret_cast<uint2*>(p.out + static_cast<size_t>(row_off[i]) * p.ldo + col) = w;
        } else if constexpr (EPI == EPI_BIAS_F32) {
          if (col < p.f32_cols) *reinterpret_cast<float4*>(p.f32_out + static_cast<size_t>(row) * p.ldo + col) = o;
        } else {  // EPI_UNPATCHIFY: 4 consecutive columns = 4 channels of one (p, q) sub-pixel
          const int s_ = row / p.tokens, n_ = row - s_ * p.tokens;  // rows [hp*wp, tokens) of a sample are padding
          if (col < p.unp_cols && n_ < p.unp_hp * p.unp_wp) {
            const int i_h = n_ / p.unp_wp, j_w = n_ - i_h * p.unp_wp;
            const int pq = col / p.unp_c, ch = col - pq * p.unp_c;
            const int hout = 2 * p.unp_hp, wout = 2 * p.unp_wp;
            float* dst = p.unp_out + ((static_cast<size_t>(s_) * p.unp_c + ch) * hout + (2 * i_h + (pq >> 1))) * wout +
                         2 * j_w + (pq & 1);
            const size_t plane = static_cast<size_t>(hout) * wout;
            dst[0] = o.x;
            dst[plane] = o.y;
            dst[2 * plane] = o.z;
            dst[3 * plane] = o.w;
          }
        }
      }
    }
    __syncwarp();  // the stage tile is overwritten by the next chunk
  };
  int limit = n_chunks;  // chunks past the real output columns of the zero-padded heads carry nothing to store
  if constexpr (EPI == EPI_UNPATCHIFY) limit = min(n_chunks, (p.unp_cols - n0 + 31) / 32);
  if constexpr (EPI == EPI_BIAS_F32) limit = min(n_chunks, (p.f32_cols - n0 + 31) / 32);
  if constexpr (EPI == EPI_CONV) limit = max(0, min(n_chunks, (p.conv_cols - n0 + 31) / 32));
  // Chunk assignment: the two warps of a lane quarter split the tile's columns into two CONTIGUOUS halves (ECADK_EPI_
  // INTERLEAVE=1 at build time restores the every-second-chunk split): a warp then walks adjacent 128-byte pieces of
  // the same 32 rows, which keeps its DRAM pages open across iterations.
#ifdef ECADK_EPI_INTERLEAVE
  const int c_begin = parity, c_step = 2, c_end = limit;
#else
  const int half_chunks = (n_chunks + 1) / 2;
  const int c_begin = parity * half_chunks, c_step = 1;
  const int c_end = min(limit, c_begin + half_chunks);
#endif
  // (measured and rejected, round 2: requesting accumulator chunk c+1 from TMEM while chunk c is processed - two more
  // register sets - changes nothing for the bf16-output epilogues and makes the gated-residual one slower, 197 -> 254 us)
  uint32_t ta[32];
  auto acc_ready = [&](const int c) {
    tmem_ld_32x32(t_row + c * 32, ta);
    tmem_ld_wait();
  };
  if constexpr (EPI == EPI_GATED_RESIDUAL) {
    float4 xa[8], xb_[8];
    Vecs va, vb;
    if (c_begin < c_end) {
      load_vecs(c_begin, va);
      load_x(c_begin, xa);
    }
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 2 * c_step) {
      const bool more = c + c_step < c_end;
      if (more) {
        load_vecs(c + c_step, vb);
        load_x(c + c_step, xb_);
      }
      acc_ready(c);
      process(c, xa, va, ta);
      if (more) {
        const bool more2 = c + 2 * c_step < c_end;
        if (more2) {
          load_vecs(c + 2 * c_step, va);
          load_x(c + 2 * c_step, xa);
        }
        acc_ready(c + c_step);
        process(c + c_step, xb_, vb, ta);
      }
    }
  } else {
    const float4 none[8] = {};
    Vecs va, vb;
    if (c_begin < c_end) load_vecs(c_begin, va);
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 2 * c_step) {
      const bool more = c + c_step < c_end;
      if (more) load_vecs(c + c_step, vb);
      acc_ready(c);
      process(c, none, va, ta);
      if (more) {
        const bool more2 = c + 2 * c_step < c_end;
        if (more2) load_vecs(c + 2 * c_step, va);
        acc_ready(c + c_step);
        process(c + c_step, none, vb, ta);
      }
    }
  }
}

// Gated-residual epilogue, TMA-store form (round 2).
//
// tools/micro/epi2_sensitivity.py (out-projection, M = 51200, N = K = 1152) showed where the old epilogue's time went:
// MMA + operands alone 95 us; + the residual-stream reads 118 us; + the 472 MB of per-thread stores 148 us; both 195 us -
// the side traffic added to the MMA time instead of hiding under it, and the stores were the larger part: 3072 row-
// strided STG transactions per tile (a warp store touches 4 rows) drained at ~21 GB/s per SM.  Here every 32 x 32
// chunk leaves through shared memory as THREE bulk tensor stores (updated fp32 residual, bf16 cache, bf16 shadow):
//
//   * the fp32 tile is the transpose tile itself: the accumulator chunk is written row-per-thread into a 128B-swizzled
//     [32 x 128 B] tile (conflict-free both ways - the XOR swizzle replaces the pitch-36 padding), read back in the
//     coalesced (row group, 4-column) layout, combined with the prefetched residual registers, and written back IN PLACE
//     (each element is read and rewritten by the same lane);
//   * the bf16 tiles are [32 x 64 B] with the 64B swizzle; one lane issues cp.async.bulk.tensor stores with matching
//     tensor maps; the staging is reused by the next chunk once `cp.async.bulk.wait_group.read` says it has been read.
//     (Measured and rejected: separate transpose / staging tiles - 12 KB per warp, one pipeline stage less - so that
//     the wait for the previous chunk's stores sits behind the next chunk's TMEM load and transpose: 180 -> 183 us.
//     Decomposition with -DECADK_EPI_DBG builds: MMA + operands 98 us, + residual reads 125 us, + the three bulk
//     stores 152 us, all 180 us; each bulk store costs ~17 us per launch whatever its size.)
__device__ __forceinline__ void epilogue_tile_residual_tma(const GemmParams& p, const CUtensorMap* tm_x,
                                                           const CUtensorMap* tm_cache, const CUtensorMap* tm_xb,
                                                           const uint32_t t_row, uint8_t* stage, const int lane,
                                                           const int parity, const int row_base, const int n0,
                                                           const int n_chunks) {
  const int cg = lane & 7, rs = lane >> 3;
  const uint32_t tile_x = smem_u32(stage);           // fp32 [32 rows x 128 B], SWIZZLE_128B
  const uint32_t tile_c = tile_x + 4096;              // bf16 [32 rows x 64 B], SWIZZLE_64B: un-gated output (cache)
  const uint32_t tile_b = tile_x + 6144;              // bf16 shadow of the updated stream
  const int sample0 = row_base / p.tokens;            // tokens % 32 == 0: one sample per 32-row chunk
  struct Vecs {
    float4 b4, g4, t4;
  };
  auto load_vecs = [&](int c, Vecs& v) {
    const int col = n0 + c * 32 + cg * 4;
    v.b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    v.g4 = make_float4(1.f, 1.f, 1.f, 1.f);
    v.t4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias != nullptr) v.b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    if (p.gate_table != nullptr || p.gate_temb != nullptr) {
      v.g4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.gate_table != nullptr) v.g4 = __ldg(reinterpret_cast<const float4*>(p.gate_table + col));
      if (p.gate_temb != nullptr)
        v.t4 = __ldg(reinterpret_cast<const float4*>(p.gate_temb + static_cast<size_t>(sample0) * p.temb_stride + col));
    }
  };
  auto load_x = [&](int c, float4 (&dst)[8]) {
    const int col = n0 + c * 32 + cg * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row_base + i * 4 + rs;
      dst[i] = row < p.M ? *reinterpret_cast<const float4*>(p.x + static_cast<size_t>(row) * p.N + col)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto process = [&](const int c, const float4 (&xin)[8], const Vecs& vec) {
    uint32_t v[32];
    tmem_ld_32x32(t_row + c * 32, v);
    const float4 b4 = vec.b4;
    const float4 g4 = make_float4(vec.g4.x + vec.t4.x, vec.g4.y + vec.t4.y, vec.g4.z + vec.t4.z, vec.g4.w + vec.t4.w);
    // the previous chunk's bulk stores must have read the staging tiles before they are overwritten
    if (lane == 0) tma_store_wait_read0();
    __syncwarp();
    tmem_ld_wait();
    // row-per-thread -> 128B-swizzled tile: 16-byte piece j of row `lane` lives at piece (j ^ (lane & 7))
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts_u4(tile_x + lane * 128 + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + rs;
      const uint32_t ax = tile_x + r * 128 + ((cg ^ (r & 7)) << 4);
      float4 o = lds_f4(ax);
      o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
      // bf16 tiles: 8-byte half (cg & 1) of 16-byte piece (cg >> 1), 64B swizzle = piece ^ ((r >> 1) & 3)
      const uint32_t ob = r * 64 + ((((cg >> 1) ^ ((r >> 1) & 3)) << 4) | ((cg & 1) << 3));
      if (p.cache != nullptr) sts_u2(tile_c + ob, pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      float4 x = xin[i];
      x.x = fmaf(g4.x, o.x, x.x); x.y = fmaf(g4.y, o.y, x.y);
      x.z = fmaf(g4.z, o.z, x.z); x.w = fmaf(g4.w, o.w, x.w);
      sts_u4(ax, __float_as_uint(x.x), __float_as_uint(x.y), __float_as_uint(x.z), __float_as_uint(x.w));
      if (p.xb != nullptr) sts_u2(tile_b + ob, pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
    }
    fence_proxy_async_smem();  // generic-proxy tile writes -> visible to the bulk stores (async proxy)
    __syncwarp();
    if (lane == 0) {
      const int col0 = n0 + c * 32;
      tma_store_2d(tm_x, stage, col0, row_base);  // rows past M are clipped by the tensor map
      if (p.cache != nullptr) tma_store_2d(tm_cache, stage + 4096, col0, row_base);
      if (p.xb != nullptr) tma_store_2d(tm_xb, stage + 6144, col0, row_base);
      tma_store_commit();
    }
  };
  const int half_chunks = (n_chunks + 1) / 2;
  const int c_begin = parity * half_chunks;
  const int c_end = min(n_chunks, c_begin + half_chunks);
  float4 xa[8], xb_[8];
  Vecs va, vb;
  if (c_begin < c_end) {
    load_vecs(c_begin, va);
    load_x(c_begin, xa);
  }
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += 2) {
    const bool more = c + 1 < c_end;
    if (more) {
      load_vecs(c + 1, vb);
      load_x(c + 1, xb_);
    }
    process(c, xa, va);
    if (more) {
      if (c + 2 < c_end) {
        load_vecs(c + 2, va);
        load_x(c + 2, xa);
      }
      process(c + 1, xb_, vb);
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                 const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_x,
                 const __grid_constant__ CUtensorMap tmap_cache, const __grid_constant__ CUtensorMap tmap_xb,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN, EPI>;
  constexpr bool kTmaEpi = EPI == EPI_GATED_RESIDUAL && ECADK_EPI_TMA_STORE;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::kStage);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStage + Cfg::kEpiBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  // warp-uniform by construction; the shuffle lets ptxas know (see the MMA issuer)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
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
      mbar_init(&tmem_empty[i], kEpiWarps);  // one arrive per epilogue warp
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  // programmatic dependent launch: the prologue above may have overlapped the previous kernel's tail; nothing below
  // touches global memory before that kernel has completed.  This grid is persistent (<= one CTA per SM), so the next
  // kernel's CTAs are released right away to set themselves up on the idle SMs.
  griddep_launch_dependents();
  griddep_wait();

  // Both single-thread roles run as WHOLE warps with only the TMA / tcgen05 instructions predicated on an elected lane.
  // Inside an `if (lane == 0)` branch ptxas cannot prove their operands warp-uniform and moves every one into a uniform
  // register through an ELECT / R2UR.BROADCAST / BRA.U.ANY loop: 4-6 R2UR and ~100 clocks of the issuing thread per
  // tcgen05.mma - harmless next to the 128-clock instructions of the 256-wide pair tiles, but three times the 32 clocks
  // of the 64-wide MMAs of the batch-1 configuration (tools/micro/mma_issue_bench.cu).
  if (warp == 0) {
    {
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
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStage);
            if (kb < p.kb_split) {
              int a_col, a_row;
              a_tile_coords(p, kb, m0, a_col, a_row);
              tma_load_2d(sa, &tmap_a, &full_bar[stage], a_col, a_row);
            } else {
              tma_load_2d(sa, &tmap_a2, &full_bar[stage], (kb - p.kb_split) * kGemmBK, m0);
            }
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kGemmBK, n0);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {
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
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kGemmBK / 16; ++k) {
              // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
              umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
            umma_commit(&empty_bar[stage]);  // smem slot is free once these MMAs retire
            if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9 -> TMEM lane quarters 2,3,0,1,2,3,0,1) =====================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * kGemmBM;
      const int n0 = (tile % num_n) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row_base = m0 + quarter * 32;
      const uint32_t t_row = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(quarter * 32) << 16);
      if constexpr (kTmaEpi) {
        epilogue_tile_residual_tma(p, &tmap_x, &tmap_cache, &tmap_xb, t_row,
                                   reinterpret_cast<uint8_t*>(epi_stage) + (warp - 2) * kEpiTmaWarpBytes, lane,
                                   (warp - 2) >> 2, row_base, n0, BN / 32);
      } else {
        float* stage = epi_stage + (warp - 2) * 32 * kEpiPitch;
        epilogue_tile<EPI>(p, t_row, stage, lane, (warp - 2) >> 2, row_base, n0, BN / 32);
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
    if constexpr (kTmaEpi) {
      if (lane == 0) tma_store_wait_all0();  // this lane's bulk stores have landed before the CTA exits
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =====================================================================================================
// Split-K variant for SMALL problems (the batch-1 latency configuration, M = 512 rows): a thread-block cluster of
// S = 2, 3 or 4 CTAs shares one 128 x BN output tile, CTA r multiplying the r-th slice of the K range.
//
// Why: with fewer tiles than SMs a GEMM is one CTA's serial operand stream - 128 x K of A plus BN x K of W through a
// single SM's ~64 B/clk L2 port (FF2 at M = 512: 1.8 MB per CTA, 21 us of a kernel whose MMAs take 2 us; 20 % of a
// batch-1 generation, profiles/r2_launches_c1_batch1.md).  More, narrower tiles re-read A; more CTAs per tile divide
// BOTH operand streams.
//
// Reduction = reduce-scatter through an L2-resident workspace: the tile's 32-column chunks are dealt out in contiguous
// ranges, CTA r OWNING chunks [r * chunks / S, (r + 1) * chunks / S).  After its MMAs every CTA writes the accumulator chunks it does not own to the workspace
// (fragment layout, one coalesced 512-byte store per warp instruction), the cluster barrier (release / acquire)
// publishes them, and each CTA adds the S - 1 foreign partials of its own chunks into its TMEM accumulator
// (tcgen05.ld + add + tcgen05.st) - after which the UNCHANGED epilogue functions run on the owned column range, so the
// epilogue's latency chain (residual-stream read -> FMA -> stores) is divided by S as well.
//
// grid = tiles * S, cluster (S, 1, 1); one tile per cluster, no persistence (the launcher checks that all clusters
// are co-resident).  Workspace: [tile][source rank][chunk][lane quarter][8][32] float4.
// =====================================================================================================
constexpr int kSplitMax = 4;
template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_splitk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                   const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_x,
                   const __grid_constant__ CUtensorMap tmap_cache, const __grid_constant__ CUtensorMap tmap_xb,
                   const GemmParams p, float4* __restrict__ ws) {
  using Cfg = GemmCfg<BN, EPI>;
  constexpr bool kTmaEpi = EPI == EPI_GATED_RESIDUAL && ECADK_EPI_TMA_STORE;
  constexpr int STAGES = Cfg::kStages;
  constexpr int kChunks = BN / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::kStage);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStage + Cfg::kEpiBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 4);

  // warp-uniform by construction; the shuffle lets ptxas know (see the MMA issuer)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int S = static_cast<int>(cluster_nctarank());
  const int rank = static_cast<int>(cluster_ctarank());
  const int tile = blockIdx.x / S;
  const int num_n = p.N / BN;
  const int m0 = (tile / num_n) * kGemmBM;
  const int n0 = (tile % num_n) * BN;
  const int num_kb = p.K / kGemmBK;
  const int kb_begin = rank * num_kb / S, kb_end = (rank + 1) * num_kb / S;  // >= 1 k-block each (launcher)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  griddep_launch_dependents();  // programmatic dependent launch: see gemm_bf16_kernel
  griddep_wait();

  const int quarter = warp & 3;
  const int par = (warp - 2) >> 2;
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
  const int own0 = rank * kChunks / S, own1 = (rank + 1) * kChunks / S;  // column chunks owned by this CTA
  const int per = own1 - own0;
  float4* const ws_tile = ws + static_cast<size_t>(tile) * S * kChunks * 1024;
  // (whole-warp roles with elected-lane TMA / tcgen05 instructions: see gemm_bf16_kernel)
  if (warp == 0) {
    {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::kStage;
        uint8_t* sb = sa + Cfg::kStageA;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStage);
          if (kb < p.kb_split) {
            int a_col, a_row;
            a_tile_coords(p, kb, m0, a_col, a_row);
            tma_load_2d(sa, &tmap_a, &full_bar[stage], a_col, a_row);
          } else {
            tma_load_2d(sa, &tmap_a2, &full_bar[stage], (kb - p.kb_split) * kGemmBK, m0);
          }
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kGemmBK, n0);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_bf16(kGemmBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStage);
        const uint32_t sb = sa + Cfg::kStageA;
        const uint64_t da = make_smem_desc(sa, 0, 1024, kLayoutSW128);
        const uint64_t db = make_smem_desc(sb, 0, 1024, kLayoutSW128);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k)
            umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb != kb_begin) || (k != 0));
          umma_commit(&empty_bar[stage]);
          if (kb == kb_end - 1) umma_commit(&tmem_full[0]);  // this CTA's partial accumulator is complete
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ===================== partials out: the chunks owned by the other CTAs =====================
    mbar_wait(&tmem_full[0], 0);
    tc_fence_after();
    for (int c = par; c < kChunks; c += 2) {
      if (c >= own0 && c < own1) continue;
      uint32_t v[32];
      tmem_ld_32x32(t_row + c * 32, v);
      tmem_ld_wait();
      float4* dst = ws_tile + (static_cast<size_t>(rank * kChunks + c) * 4 + quarter) * 256 + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        __stcg(dst + j * 32, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
    }
  }
  // every thread of every CTA of the cluster: the partials above are released, the peers' acquired
  __syncwarp();
  cluster_sync();
  if (warp >= 2) {
    // ===================== reduce the owned chunks into TMEM, then the ordinary epilogue on them =====================
    const int half = (per + 1) / 2;  // the epilogue's own split of a column range between the two warps of a quarter
    const int c_begin = par * half, c_end = min(per, c_begin + half);
    for (int c = c_begin; c < c_end; ++c) {
      const int cc = own0 + c;
      uint32_t v[32];
      tmem_ld_32x32(t_row + cc * 32, v);
      float4 pr[kSplitMax - 1][8];
#pragma unroll
      for (int i = 0; i < kSplitMax - 1; ++i) {
        int src = rank + 1 + i;
        src = src >= S ? src - S : src;
        const float4* sp = ws_tile + (static_cast<size_t>(src * kChunks + cc) * 4 + quarter) * 256 + lane;
#pragma unroll
        for (int j = 0; j < 8; ++j) pr[i][j] = (i < S - 1) ? __ldcg(sp + j * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 a = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
#pragma unroll
        for (int i = 0; i < kSplitMax - 1; ++i) {
          a.x += pr[i][j].x; a.y += pr[i][j].y; a.z += pr[i][j].z; a.w += pr[i][j].w;
        }
        v[4 * j] = __float_as_uint(a.x);
        v[4 * j + 1] = __float_as_uint(a.y);
        v[4 * j + 2] = __float_as_uint(a.z);
        v[4 * j + 3] = __float_as_uint(a.w);
      }
      tmem_st_32x32(t_row + cc * 32, v);
    }
    tmem_st_wait();
    const int row_base = m0 + quarter * 32;
    if constexpr (kTmaEpi) {
      epilogue_tile_residual_tma(p, &tmap_x, &tmap_cache, &tmap_xb, t_row + own0 * 32,
                                 reinterpret_cast<uint8_t*>(epi_stage) + (warp - 2) * kEpiTmaWarpBytes, lane, par,
                                 row_base, n0 + own0 * 32, per);
      if (lane == 0) tma_store_wait_all0();
    } else {
      float* stage = epi_stage + (warp - 2) * 32 * kEpiPitch;
      epilogue_tile<EPI>(p, t_row + own0 * 32, stage, lane, par, row_base, n0 + own0 * 32, per);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =====================================================================================================
// 2-CTA variant (cta_group::2): a CTA pair on one TPC computes a 256 x BN tile.  Each CTA stages ITS 128 rows of A
// and ITS half (BN/2 rows) of W, so the L2->SM traffic per FLOP drops by 1/3 (BN = 256) versus the 1-CTA kernel,
// which is L2-bandwidth-bound on B200.  The leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 for the
// pair; tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals to both CTAs; every CTA's epilogue
// warps drain their own 128 TMEM lanes.
// =====================================================================================================
// TAP3 (3x3 convolutions, see below): a stage holds ONE A tile of 128 + 8 rows that starts one pixel to the left of
// the output rows, and the three W tiles of the taps (ky, 0..2) that read it.
// Each CTA computes TWO 128-row sub-tiles there (rows m0 .. m0+255), which share the stage's W tiles: with one sub-tile
// the narrow (C_out = 128) tile is bound by its W stream - 24 KB of W per 17 KB of A per 768 MMA clocks is the SM's L2
// port (ncu: tensor pipe 55-66 %).  The two accumulators sit side by side in one 256-column TMEM buffer.
#ifndef ECADK_GEMM2_UNIFORM
#define ECADK_GEMM2_UNIFORM 1
#endif
constexpr bool kUniformRoles = ECADK_GEMM2_UNIFORM != 0;
constexpr int kTap3Rows = kGemmBM + 8;
constexpr int kTap3Sub = 2;
template <int BN, int EPI = 0, bool TAP3 = false>
struct Gemm2Cfg {
  static constexpr int kSub = TAP3 ? kTap3Sub : 1;             // 128-row sub-tiles per CTA
  static constexpr int kBytesA = (TAP3 ? kTap3Rows : kGemmBM) * kGemmBK * 2;  // one A tile: 16 KB (17 KB)
  static constexpr int kTileA = (kBytesA + 1023) / 1024 * 1024;
  static constexpr int kStageA = kSub * kTileA;
  static constexpr int kTileB = (BN / 2) * kGemmBK * 2;        // this CTA's half of one W tile
  static constexpr int kStageB = (TAP3 ? 3 : 1) * kTileB;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kEpiBytes = epi_stage_bytes(EPI);
  static constexpr int kStages = fit_stages(TAP3 ? 3 : ((BN <= 128) ? 7 : (BN <= 192 ? 6 : 5)), kStage, EPI);
  static_assert(!TAP3 || kTap3Sub * BN <= 256, "the sub-tile accumulators share one 256-column TMEM buffer");
  static constexpr int kAccStride = 256;
  static constexpr int kTmemCols = 512;
  static constexpr int kSmemBytes = kStages * kStage + kEpiBytes + 1024 + 256;
  static_assert(kStages >= (TAP3 ? 3 : 4), "GEMM pipeline depth");
};

// TAP3 = true (3x3 convolution, taps of one kernel row share their A tile): the implicit GEMM above fetches every A block
// nine times, once per tap, and at C_out = 128 - the tile already spans all output channels - that alone is the SM's
// whole L2 port (2 B of A per 256 FLOP; ncu: tensor pipe 31.6 % at the 256 x 256 layers of the VAE decoder).  The three
// taps (ky, 0), (ky, 1), (ky, 2) read the SAME rows shifted by one pixel, so one TMA load of 128 + 8 rows starting at
// `m0 + (ky-1)*pitch - 1` serves all three: the MMA of tap kx addresses the tile at row offset kx through its
// shared-memory descriptor (start address + kx * 128 B; the 128B swizzle follows the absolute address, so the rows
// land where TMA put them).  A traffic drops 3x; k-blocks are ordered (ky, channel block) with three W tiles per stage.
template <int BN, int EPI, bool TAP3 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                  const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_b_tail,
                  const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_cache,
                  const __grid_constant__ CUtensorMap tmap_xb, const GemmParams p, const int tail) {
  // `tail` (0 or 128, only with BN = 256): N = k*256 + 128 is covered by k full-width tiles plus one 128-wide tile per
  // row block, so N = 1152 / 3456 run at the L2->SM traffic per FLOP of 256-wide tiles instead of 192-wide ones.
  using Cfg = Gemm2Cfg<BN, EPI, TAP3>;
  constexpr bool kTmaEpi = EPI == EPI_GATED_RESIDUAL && ECADK_EPI_TMA_STORE;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::kStage);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStage + Cfg::kEpiBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  // warp-uniform by construction; the shuffle lets ptxas know (see the MMA issuer)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  constexpr int kRowsCta = Cfg::kSub * kGemmBM;  // rows of a tile per CTA (the pair's tile has twice as many)
  const int num_m = (p.M + 2 * kRowsCta - 1) / (2 * kRowsCta);
  const int num_n_full = (p.N - tail) / BN;
  const int num_n = num_n_full + (tail ? 1 : 0);
  const int num_tiles = num_m * num_n;
  const int num_kb = TAP3 ? p.K / kGemmBK / 3 : p.K / kGemmBK;  // TAP3: (ky, channel block) steps of three taps each

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (tail) tma_prefetch_desc(&tmap_b_tail);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // used in the leader: its producer's arrive.expect_tx (bytes of both CTAs)
      mbar_init(&empty_bar[i], 1);  // per CTA: multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);   // per CTA: multicast tcgen05.commit
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);  // used in the leader: 8 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync();  // barriers of both CTAs initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  griddep_launch_dependents();  // see gemm_bf16_kernel
  griddep_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m0 = (tile / num_n) * (2 * kRowsCta) + cta * kRowsCta;
      const int n_idx = tile % num_n;
      const bool is_tail = n_idx >= num_n_full;
      const int bn_cur = is_tail ? tail : BN;
      // (measured: an L2 bulk prefetch of this tile's residual-stream block issued here does NOT help - out-projection
      // 197 -> 202 us, FF2 427 -> 488 us: the epilogue is not exposed-latency-bound and the prefetch competes with the
      // operand stream for L2)
      // (kUniformRoles: whole-warp roles with elected-lane TMA / tcgen05 instructions, see gemm_bf16_kernel)
      if (kUniformRoles || lane == 0) {
        const int n0 = n_idx * BN + cta * (bn_cur / 2);
        const CUtensorMap* tb = is_tail ? &tmap_b_tail : &tmap_b;
        const uint32_t stage_bytes =
            TAP3 ? 2 * (Cfg::kSub * Cfg::kBytesA + Cfg::kStageB) : 2 * (Cfg::kStageA + (bn_cur / 2) * kGemmBK * 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStage;
          uint8_t* sb = sa + Cfg::kStageA;
          if (!kUniformRoles || elect_one()) {
          if (cta == 0) mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          if constexpr (TAP3) {
            const int ky = kb / p.conv_cblocks, cb = kb - ky * p.conv_cblocks;
#pragma unroll
            for (int sub = 0; sub < Cfg::kSub; ++sub)
              tma_load_2d_2sm(sa + sub * Cfg::kTileA, &tmap_a, &full_bar[stage], cb * kGemmBK,
                              m0 + sub * kGemmBM + (ky - 1) * p.conv_pitch - 1);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              tma_load_2d_2sm(sb + kx * Cfg::kTileB, tb, &full_bar[stage],
                              ((ky * 3 + kx) * p.conv_cblocks + cb) * kGemmBK, n0);
          } else if (kb < p.kb_split) {
            int a_col, a_row;
            a_tile_coords(p, kb, m0, a_col, a_row);
            tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], a_col, a_row);
          } else {
            tma_load_2d_2sm(sa, &tmap_a2, &full_bar[stage], (kb - p.kb_split) * kGemmBK, m0);
          }
          if constexpr (!TAP3) tma_load_2d_2sm(sb, tb, &full_bar[stage], kb * kGemmBK, n0);
          }
          if (kUniformRoles) __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      __syncwarp();  // lanes 1..31 must not run ahead: the prefetch distance stays one tile
    }
  } else if (warp == 1) {
    if ((kUniformRoles || lane == 0) && cta == 0) {
      // ===================== MMA issuer (leader CTA only) =====================
      constexpr uint32_t idesc_full = make_idesc_bf16(2 * kGemmBM, BN);
      constexpr uint32_t idesc_tail = make_idesc_bf16(2 * kGemmBM, 128);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const uint32_t idesc = (tile % num_n) >= num_n_full ? idesc_tail : idesc_full;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStage);
          const uint32_t sb = sa + Cfg::kStageA;
          if (!kUniformRoles || elect_one()) {
          if constexpr (TAP3) {
#pragma unroll
            for (int sub = 0; sub < Cfg::kSub; ++sub) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              // rows kx .. kx + 127 of the 136-row tile: start address + kx rows.  The 128B swizzle is a function of
              // the absolute shared-memory address bits (the tile base stays 1024-byte aligned), so the descriptor's
              // matrix-base-offset field stays 0 (measured: setting it to kx reads the wrong 16-byte pieces)
              const uint64_t da = make_smem_desc(sa + sub * Cfg::kTileA + kx * 128, 16, 1024, kLayoutSW128);
              const uint64_t db = make_smem_desc(sb + kx * Cfg::kTileB, 16, 1024, kLayoutSW128);
#pragma unroll
              for (int k = 0; k < kGemmBK / 16; ++k)
                umma_bf16_ss_2sm(d_tmem + sub * BN, da + 2 * k, db + 2 * k, idesc, (kb | kx | k) != 0);
            }
            }
          } else {
            const uint64_t da = make_smem_desc(sa, 16, 1024, kLayoutSW128);
            const uint64_t db = make_smem_desc(sb, 16, 1024, kLayoutSW128);
#pragma unroll
            for (int k = 0; k < kGemmBK / 16; ++k) {
              umma_bf16_ss_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
          }
          umma_commit_2sm(&empty_bar[stage], 0b11);
          if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[acc], 0b11);
          }
          if (kUniformRoles) __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (both CTAs; this CTA's 128 rows) =====================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m0 = (tile / num_n) * (2 * kRowsCta) + cta * kRowsCta;
      const int n_idx = tile % num_n;
      const int n0 = n_idx * BN;
      const int n_chunks = (n_idx >= num_n_full ? tail : BN) / 32;
      // (measured and rejected, round 2: an L2 prefetch of the row pieces this warp reads in the NEXT tile, one 512-byte
      // cp.async.bulk.prefetch.L2 per lane issued here - out-projection 196 -> 200 us and DRAM reads 363 -> 500 MB under
      // ncu: the residual-stream reads are not what the epilogue waits for)
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row_base = m0 + quarter * 32;
      const uint32_t t_row = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(quarter * 32) << 16);
      if constexpr (kTmaEpi) {
        epilogue_tile_residual_tma(p, &tmap_x, &tmap_cache, &tmap_xb, t_row,
                                   reinterpret_cast<uint8_t*>(epi_stage) + (warp - 2) * kEpiTmaWarpBytes, lane,
                                   (warp - 2) >> 2, row_base, n0, n_chunks);
      } else {
        float* stage = epi_stage + (warp - 2) * 32 * kEpiPitch;
#pragma unroll 1
        for (int sub = 0; sub < Cfg::kSub; ++sub)  // TAP3: the second sub-tile's accumulator follows the first in TMEM
          epilogue_tile<EPI>(p, t_row + sub * BN, stage, lane, (warp - 2) >> 2, row_base + sub * kGemmBM, n0, n_chunks);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);  // the leader's barrier collects both CTAs
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if constexpr (kTmaEpi) {
      if (lane == 0) tma_store_wait_all0();
    }
  }

  tc_fence_before();
  cluster_sync();  // the peer may still be reading our smem / signalling our barriers until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace ecadk
