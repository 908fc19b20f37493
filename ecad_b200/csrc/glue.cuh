// Memory-bound glue of the PixArt hot path: vectorised, coalesced, warp-shuffle reductions (no tensor cores).
//
//   residual_ln_kernel   K2/K9/K12 of SURVEY.md section 2b in one kernel: optional cached-residual reuse
//                        (x += sum_j gate_j * cache_j), optional bf16 shadow of x, optional LayerNorm + adaLN
//                        modulate -> bf16.  One warp per token row, the row stays in registers.
//   patch_embed_kernel   K13: 2x2/s2 patch conv + bias + 2-D sincos position table -> fp32 residual stream.
//   (K16, the final layer, = residual_ln_kernel + the GEMM's EPI_UNPATCHIFY epilogue; see capi.cu)
//   small_linear_kernel  K14: fp32 GEMV-ish linears of the timestep / adaLN-single embedder.
//   timestep_sinusoid_kernel, cast_to_bf16_kernel, mask_bias_kernel, cfg_dpm_step_kernel (K17).
#pragma once
#include "ptx.cuh"

namespace ecadk {

constexpr int kMaxReuse = 12;

struct ReuseEntry {
  const __nv_bfloat16* cache;  // [M, D] cached un-gated sub-block output
  const float* gate_table;     // [D] gate row of the block's scale_shift_table, or null (attn2: no gate)
  const float* gate_temb;      // [samples, temb_stride] offset to the gate chunk
};

struct ResidualLnParams {
  float* x;            // [M, D] fp32 residual stream (read; written when n_reuse > 0)
  __nv_bfloat16* xb;   // optional bf16 copy of the (updated) stream
  __nv_bfloat16* h;    // optional LN+modulate output
  int M, tokens;       // rows; rows per sample
  int n_reuse;
  ReuseEntry reuse[kMaxReuse];
  const float* shift_table;  // [D]   (LN path)
  const float* scale_table;  // [D]
  const float* shift_temb;   // [samples, temb_stride] offset to the shift chunk
  const float* scale_temb;   // [samples, temb_stride] offset to the scale chunk
  int temb_stride;
  float eps;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// VPL = float4 vectors per lane: D = 128 * VPL  (1152 -> 9, 3072 -> 24)
//
// A block owns 32 consecutive rows of ONE sample (tokens % 32 == 0): the per-sample modulation vectors
// (1 + scale, shift, and every reuse gate = table + temb[sample]) are combined once into shared memory instead of
// being re-read from L2 by every row; each warp then streams 4 rows, two at a time so that two rows of x (and of each
// cache) are in flight per warp.
constexpr int kRlnRowsPerBlock = 32;

// cached rows of reuse entries [r0, r0 + RB) of one token row, all requested before any is used
template <int VPL, int RB>
__device__ __forceinline__ void rln_load_reuse(const ResidualLnParams& p, const int r0, const int row, const int lane,
                                               uint2 (&c)[RB][VPL]) {
  const size_t roff = static_cast<size_t>(row) * (128 * VPL);
#pragma unroll
  for (int u = 0; u < RB; ++u) {
    if (r0 + u < p.n_reuse) {
      const uint2* cr = reinterpret_cast<const uint2*>(p.reuse[r0 + u].cache + roff);
#pragma unroll
      for (int i = 0; i < VPL; ++i) c[u][i] = __ldg(cr + lane + 32 * i);
    }
  }
}

template <int VPL, int RB>
__device__ __forceinline__ void rln_add_reuse(const ResidualLnParams& p, const float4* __restrict__ sm, const int r0,
                                              const int lane, const uint2 (&c)[RB][VPL], float4 (&v)[VPL]) {
  constexpr int DV = 32 * VPL;
#pragma unroll
  for (int u = 0; u < RB; ++u) {
    if (r0 + u < p.n_reuse) {
      const float4* g = sm + (2 + r0 + u) * DV;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float2 c01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&c[u][i].x));
        const float2 c23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&c[u][i].y));
        const float4 gg = g[lane + 32 * i];
        v[i].x = fmaf(gg.x, c01.x, v[i].x);
        v[i].y = fmaf(gg.y, c01.y, v[i].y);
        v[i].z = fmaf(gg.z, c23.x, v[i].z);
        v[i].w = fmaf(gg.w, c23.y, v[i].w);
      }
    }
  }
}

// RB = reuse entries whose cached rows are in flight together (1 in the streaming configuration, where many rows per
// SM hide the latency; 3 in the small-batch one, where a row's reuse chain IS the kernel's latency: 6 serial round
// trips made the batch-1 flush kernels 17-20 us instead of 4.5).  PRE: the first group was requested by the caller.
template <int VPL, int RB = 1, bool PRE = false>
__device__ __forceinline__ void rln_row(const ResidualLnParams& p, const float4* __restrict__ sm, const int row,
                                        const int lane, float4 (&v)[VPL], uint2 (*pre)[VPL] = nullptr) {
  constexpr int D = 128 * VPL;
  constexpr int DV = D / 4;
  const size_t roff = static_cast<size_t>(row) * D;
  if (p.n_reuse > 0) {
    for (int r = 0; r < p.n_reuse; r += RB) {
      uint2 c[RB][VPL];
      if (PRE && r == 0) {
#pragma unroll
        for (int u = 0; u < RB; ++u)
#pragma unroll
          for (int i = 0; i < VPL; ++i) c[u][i] = pre[u][i];
      } else {
        rln_load_reuse<VPL, RB>(p, r, row, lane, c);
      }
      rln_add_reuse<VPL, RB>(p, sm, r, lane, c, v);
    }
    float4* xw = reinterpret_cast<float4*>(p.x + roff);
#pragma unroll
    for (int i = 0; i < VPL; ++i) xw[lane + 32 * i] = v[i];
  }
  if (p.xb != nullptr) {
    uint2* bw = reinterpret_cast<uint2*>(p.xb + roff);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      uint2 o;
      o.x = pack_bf16x2(v[i].x, v[i].y);
      o.y = pack_bf16x2(v[i].z, v[i].w);
      bw[lane + 32 * i] = o;
    }
  }
  if (p.h != nullptr) {
    // two-pass LayerNorm statistics in fp32 (no affine, eps inside the sqrt), like torch.nn.LayerNorm
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + p.eps);
    uint2* hw = reinterpret_cast<uint2*>(p.h + roff);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int k = lane + 32 * i;
      const float4 sc = sm[k], sh = sm[DV + k];  // (1 + scale), shift
      uint2 o;
      o.x = pack_bf16x2((v[i].x - mean) * rstd * sc.x + sh.x, (v[i].y - mean) * rstd * sc.y + sh.y);
      o.y = pack_bf16x2((v[i].z - mean) * rstd * sc.z + sh.z, (v[i].w - mean) * rstd * sc.w + sh.w);
      hw[k] = o;
    }
  }
}

template <int VPL, int GB = 1>
__device__ __forceinline__ void rln_stage_vectors(const ResidualLnParams& p, float4* rln_sm, const int sample) {
  constexpr int DV = 32 * VPL;
  for (int k = threadIdx.x; k < DV; k += blockDim.x) {
    if (p.h != nullptr) {
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);  // a null table = modulation from the per-sample vector only
      const float4 ca = p.scale_table ? __ldg(reinterpret_cast<const float4*>(p.scale_table) + k) : z4;
      const float4 cb = __ldg(reinterpret_cast<const float4*>(p.scale_temb + static_cast<size_t>(sample) * p.temb_stride) + k);
      const float4 sa = p.shift_table ? __ldg(reinterpret_cast<const float4*>(p.shift_table) + k) : z4;
      const float4 sb = __ldg(reinterpret_cast<const float4*>(p.shift_temb + static_cast<size_t>(sample) * p.temb_stride) + k);
      rln_sm[k] = make_float4(1.f + (ca.x + cb.x), 1.f + (ca.y + cb.y), 1.f + (ca.z + cb.z), 1.f + (ca.w + cb.w));
      rln_sm[DV + k] = make_float4(sa.x + sb.x, sa.y + sb.y, sa.z + sb.z, sa.w + sb.w);
    }
    // gate = table + per-sample vector (PixArt adaLN-single), per-sample vector alone (FLUX), or 1 (no gate); the
    // loads of GB reuse entries are in flight together
    for (int r0 = 0; r0 < p.n_reuse; r0 += GB) {
      float4 ga[GB], gb[GB];
#pragma unroll
      for (int u = 0; u < GB; ++u) {
        ga[u] = make_float4(1.f, 1.f, 1.f, 1.f);
        gb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + u < p.n_reuse) {
          const ReuseEntry& e = p.reuse[r0 + u];
          if (e.gate_table != nullptr || e.gate_temb != nullptr) {
            ga[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.gate_table != nullptr) ga[u] = __ldg(reinterpret_cast<const float4*>(e.gate_table) + k);
            if (e.gate_temb != nullptr)
              gb[u] = __ldg(reinterpret_cast<const float4*>(e.gate_temb + static_cast<size_t>(sample) * p.temb_stride) + k);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < GB; ++u) {
        if (r0 + u < p.n_reuse)
          rln_sm[(2 + r0 + u) * DV + k] = make_float4(ga[u].x + gb[u].x, ga[u].y + gb[u].y, ga[u].z + gb[u].z, ga[u].w + gb[u].w);
      }
    }
  }
}

template <int VPL>
__device__ __forceinline__ void rln_load_row(const ResidualLnParams& p, const int row, const int lane, float4 (&v)[VPL]) {
  const float4* xr = reinterpret_cast<const float4*>(p.x + static_cast<size_t>(row) * (128 * VPL));
#pragma unroll
  for (int i = 0; i < VPL; ++i) v[i] = xr[lane + 32 * i];
}

// RPB = rows per block: 32 for large problems (the modulation vectors are staged once per 32 rows); 8 when the whole
// problem has fewer than ~2 waves of 32-row blocks (batch-1 latency configuration: M = 512 rows would be 16 blocks of
// four serial row-pairs per warp - 14 us per launch, 25 % of a batch-1 generation; 64 blocks of one row per warp
// take the latency of a single HBM round trip).
template <int VPL, int RPB = kRlnRowsPerBlock>
__global__ void __launch_bounds__(256, (VPL <= 12 && RPB != 8) ? 2 : 1) residual_ln_kernel(const ResidualLnParams p) {
  extern __shared__ float4 rln_sm[];  // [(2 + n_reuse)][D/4]
  // programmatic dependent launch: a small grid releases the next kernel at once (its CTAs set up on idle SMs); a
  // multi-wave grid keeps the SM slots for its own blocks and lets the implicit trigger at exit do it
  if (gridDim.x <= 296) griddep_launch_dependents();
  griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * RPB;
  const int sample = row0 / p.tokens;
  if constexpr (VPL > 12 || RPB == 8) {  // one row per warp at a time (wide rows: the row fills the registers)
    float4 va[VPL];
    int ra = row0 + warp;
    if constexpr (RPB == 8) {
      // one row per warp, once: the row and its first three cached rows are requested before the vectors are staged
      constexpr int RB = VPL <= 12 ? 3 : 1;
      uint2 pre[RB][VPL];
      if (ra < p.M) {
        rln_load_row<VPL>(p, ra, lane, va);
        rln_load_reuse<VPL, RB>(p, 0, ra, lane, pre);
      }
      rln_stage_vectors<VPL, 4>(p, rln_sm, sample);
      __syncthreads();
      if (ra < p.M) rln_row<VPL, RB, true>(p, rln_sm, ra, lane, va, pre);
    } else {
      rln_stage_vectors<VPL>(p, rln_sm, sample);
      __syncthreads();
#pragma unroll 1
      for (int j = 0; j < RPB / 8; ++j) {
        ra = row0 + warp + 8 * j;
        if (ra >= p.M) break;
        rln_load_row<VPL>(p, ra, lane, va);
        rln_row<VPL>(p, rln_sm, ra, lane, va);
      }
    }
  } else {
    // rows row0 + warp + 8*j, two in flight per warp; the first pair is requested BEFORE the modulation vectors are
    // staged so the block's start-up latency overlaps with its first HBM reads
    float4 va[VPL], vb[VPL];
    int ra = row0 + warp, rb = ra + 8;
    if (ra < p.M) rln_load_row<VPL>(p, ra, lane, va);
    if (rb < p.M) rln_load_row<VPL>(p, rb, lane, vb);
    rln_stage_vectors<VPL>(p, rln_sm, sample);
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < RPB / 8; j += 2) {
      if (j > 0) {
        ra = row0 + warp + 8 * j;
        rb = ra + 8;
        if (ra < p.M) rln_load_row<VPL>(p, ra, lane, va);
        if (rb < p.M) rln_load_row<VPL>(p, rb, lane, vb);
      }
      if (ra < p.M) rln_row<VPL>(p, rln_sm, ra, lane, va);
      if (rb < p.M) rln_row<VPL>(p, rln_sm, rb, lane, vb);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// K13 patch embed: latents [S, C, Hl, Wl] fp32 -> x [S, N, D] fp32, N = (Hl/2)*(Wl/2).
// wt = conv weight transposed to [C*4, D] (k = c*4 + p*2 + q), pos = [N, D] sincos table.
struct PatchEmbedParams {
  const float* latents;
  const float* wt;
  const float* bias;
  const float* pos;
  float* x;
  int S, C, Hl, Wl, D;
  int tokens_pad;  // rows per sample of x (>= N; the rows [N, tokens_pad) are zeroed by the launcher)
};
constexpr int kPatchTokensPerBlock = 8;
__global__ void __launch_bounds__(288) patch_embed_kernel(const PatchEmbedParams p) {
  const int Wp = p.Wl >> 1, Hp = p.Hl >> 1;
  const int N = Wp * Hp;
  const int K = p.C * 4;
  const int token0 = blockIdx.x * kPatchTokensPerBlock;  // s*N + n
  const int total = p.S * N;
  __shared__ float in[kPatchTokensPerBlock][64];
  for (int idx = threadIdx.x; idx < kPatchTokensPerBlock * K; idx += blockDim.x) {
    const int tkn = idx / K, kk = idx - tkn * K;
    const int token = token0 + tkn;
    float val = 0.f;
    if (token < total) {
      const int s = token / N, n = token - s * N;
      const int i = n / Wp, j = n - i * Wp;
      const int c = kk >> 2, pq = kk & 3;
      val = p.latents[((static_cast<size_t>(s) * p.C + c) * p.Hl + (2 * i + (pq >> 1))) * p.Wl + 2 * j + (pq & 1)];
    }
    in[tkn][kk] = val;
  }
  __syncthreads();
  for (int d4 = threadIdx.x; d4 < p.D / 4; d4 += blockDim.x) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias) + d4);
    float4 acc[kPatchTokensPerBlock];
#pragma unroll
    for (int t = 0; t < kPatchTokensPerBlock; ++t) acc[t] = b;
    for (int k = 0; k < K; ++k) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(p.wt + static_cast<size_t>(k) * p.D) + d4);
#pragma unroll
      for (int t = 0; t < kPatchTokensPerBlock; ++t) {
        const float a = in[t][k];
        acc[t].x = fmaf(a, w.x, acc[t].x);
        acc[t].y = fmaf(a, w.y, acc[t].y);
        acc[t].z = fmaf(a, w.z, acc[t].z);
        acc[t].w = fmaf(a, w.w, acc[t].w);
      }
    }
#pragma unroll
    for (int t = 0; t < kPatchTokensPerBlock; ++t) {
      const int token = token0 + t;
      if (token < total) {
        const int n = token % N;
        const float4 pe = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(n) * p.D) + d4);
        float4 o = acc[t];
        o.x += pe.x; o.y += pe.y; o.z += pe.z; o.w += pe.w;
        const size_t xrow = static_cast<size_t>(token / N) * p.tokens_pad + n;
        reinterpret_cast<float4*>(p.x + xrow * p.D)[d4] = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// K14 helpers.  y[s, o] = b[o] + sum_i W[o, i] * act_in(x[s, i]);  act_in: 0 none, 1 SiLU.  fp32 throughout.
// One warp per output feature; the weight row is read once and reused across all samples.
struct SmallLinearParams {
  const float* x;  // [S, K] with row pitch ldx
  const float* w;  // [O, K]
  const float* b;  // [O]
  float* y;        // [S, ldy] written at column offset y_off
  int S, K, O, ldy, y_off, act_in, accumulate, ldx;
};
__global__ void __launch_bounds__(256) small_linear_kernel(const SmallLinearParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + warp;
  if (o >= p.O) return;
  const float* wr = p.w + static_cast<size_t>(o) * p.K;
  {
    const int s0 = blockIdx.y * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int k = lane; k < p.K; k += 32) {
      const float w = __ldg(wr + k);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (s0 + j < p.S) {
          float a = __ldg(p.x + static_cast<size_t>(s0 + j) * p.ldx + k);
          if (p.act_in == 1) a = a / (1.f + __expf(-a));
          acc[j] = fmaf(w, a, acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float r = warp_sum(acc[j]);
      if (lane == 0 && s0 + j < p.S) {
        float* dst = p.y + static_cast<size_t>(s0 + j) * p.ldy + p.y_off + o;
        const float val = r + p.b[o];
        *dst = p.accumulate ? (*dst + val) : val;
      }
    }
  }
}

// diffusers Timesteps(256, flip_sin_to_cos=True, shift 0): out[s] = [cos(t f_i) | sin(t f_i)], f_i = 1e4^(-i/128)
__global__ void timestep_sinusoid_kernel(const float* t, float* out, int S, int dim) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim >> 1;
  if (idx >= S * half) return;
  const int s = idx / half, i = idx - s * half;
  const float f = expf(-9.210340371976184f * static_cast<float>(i) / static_cast<float>(half));
  const float a = t[s] * f;
  out[static_cast<size_t>(s) * dim + i] = cosf(a);
  out[static_cast<size_t>(s) * dim + half + i] = sinf(a);
}

__global__ void cast_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n4) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = o;
  }
}

// out = bf16(SiLU(in)): the activation in front of every AdaLayerNormZero / ZeroSingle / Continuous linear of FLUX; the
// result is the A operand of ONE stacked modulation GEMM per step
__global__ void silu_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n4) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    uint2 o;
    o.x = pack_bf16x2(v.x / (1.0f + __expf(-v.x)), v.y / (1.0f + __expf(-v.y)));
    o.y = pack_bf16x2(v.z / (1.0f + __expf(-v.z)), v.w / (1.0f + __expf(-v.w)));
    reinterpret_cast<uint2*>(out)[i] = o;
  }
}

// TGATE: cache[0:n] = (cache[0:n] + cache[n:2n]) / 2 (bf16, in place) - the averaged cross-attention output that
// replaces the CFG pair from the gate step on (cached_transformer_block.py:443-449)
__global__ void average_halves_kernel(__nv_bfloat16* buf, size_t n8) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  uint4* lo = reinterpret_cast<uint4*>(buf);
  const uint4* hi = reinterpret_cast<const uint4*>(buf) + n8;
  for (; i < n8; i += stride) {
    const uint4 a = lo[i], b = hi[i];
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&av[k]));
      const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bv[k]));
      o[k] = pack_bf16x2((fa.x + fb.x) * 0.5f, (fa.y + fb.y) * 0.5f);
    }
    lo[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// FLUX: per-head RMSNorm (learned weight) + rotary embedding on head-major q and k, in place.
// q, k: bf16 [B, H, S, 128]; tokens s < split use the "added" (text-stream) norm weights; rope: fp32 [S, 64] cos / sin.
// One warp per (b, h, s) row of q and of k (lane = 4 consecutive elements = 2 rotation pairs).
struct QkNormRopeParams {
  __nv_bfloat16* q;
  __nv_bfloat16* k;
  const float* wq;       // [128] norm_q.weight
  const float* wk;       // [128] norm_k.weight
  const float* wq_add;   // [128] norm_added_q.weight (text tokens), or null
  const float* wk_add;   // [128]
  const float* cos_t;    // [S, 64], or [B, S, 64] with sample_stride = S * 64
  const float* sin_t;
  int rows;              // B * H * S
  int S, split;
  float eps;
  int rows_per_sample;   // H * S
  int sample_stride;     // floats between the rotation tables of two samples; 0 = one table for the batch
};
__global__ void __launch_bounds__(256) qk_norm_rope_kernel(const QkNormRopeParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= p.rows) return;
  const int s = row % p.S;
  const bool added = s < p.split;
  const size_t tab = static_cast<size_t>(row / p.rows_per_sample) * p.sample_stride + static_cast<size_t>(s) * 64;
  const float2 c2 = __ldg(reinterpret_cast<const float2*>(p.cos_t + tab) + lane);
  const float2 s2 = __ldg(reinterpret_cast<const float2*>(p.sin_t + tab) + lane);
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    __nv_bfloat16* base = (which == 0 ? p.q : p.k) + static_cast<size_t>(row) * 128;
    const float* w = which == 0 ? (added ? p.wq_add : p.wq) : (added ? p.wk_add : p.wk);
    const uint2 raw = *reinterpret_cast<const uint2*>(base + lane * 4);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    const float ss = warp_sum(a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y);
    const float r = rsqrtf(ss * (1.0f / 128.0f) + p.eps);
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + lane);
    const float x0 = a.x * r * wv.x, x1 = a.y * r * wv.y, x2 = b.x * r * wv.z, x3 = b.y * r * wv.w;
    uint2 o;
    o.x = pack_bf16x2(c2.x * x0 - s2.x * x1, s2.x * x0 + c2.x * x1);
    o.y = pack_bf16x2(c2.y * x2 - s2.y * x3, s2.y * x2 + c2.y * x3);
    *reinterpret_cast<uint2*>(base + lane * 4) = o;
  }
}

// dst[r, 0:cols] = op(src[r, 0:cols]) with independent row pitches; op: 0 copy, 1 GELU(tanh).  bf16, cols % 8 == 0.
__global__ void strided_unary_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                     int rows, int cols8, int ld_src, int ld_dst, int op) {
  const size_t total = static_cast<size_t>(rows) * cols8;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const size_t r = i / cols8;
    const int c = static_cast<int>(i - r * cols8) * 8;
    uint4 v = __ldg(reinterpret_cast<const uint4*>(src + r * ld_src + c));
    if (op == 1) {
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
        w[k] = pack_bf16x2(gelu_tanh(f.x), gelu_tanh(f.y));
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(dst + r * ld_dst + c) = v;
  }
}

// y += a * x (fp32): the flow-matching Euler update  latents += (sigma_next - sigma) * velocity
__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, size_t n) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) y[i] = fmaf(a, x[i], y[i]);
}

// (1 - mask) * -10000 for real text tokens (pixart_transformer_2d_edited.py:282-289), -inf for padding keys
__global__ void mask_bias_kernel(const float* mask, float* bias, int S, int T, int T_pad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * T_pad) return;
  const int s = idx / T_pad, t = idx - s * T_pad;
  bias[idx] = t < T ? (1.0f - mask[s * T + t]) * -10000.0f : -INFINITY;
}

// K17: classifier-free guidance + learned-sigma drop + one DPM-Solver++(2M) update, fused.
//   eps = uncond + g * (text - uncond) on channels [0, C);  x0 = (x - sigma_s * eps) / alpha_s
//   x_next = c_x * x + c_d0 * x0 + c_d1 * x0_prev      (host folds the 1st/2nd-order coefficients)
struct CfgDpmParams {
  const float* noise;  // [2B or B, 2C, H, W] transformer output
  float* latents;      // [B, C, H, W] in/out
  float* x0_prev;      // [B, C, H, W] in/out (previous x0 prediction)
  int B, C, HW, has_cfg;
  float guidance, sigma_s, alpha_s, c_x, c_d0, c_d1;
};
__global__ void cfg_dpm_step_kernel(const CfgDpmParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = p.C * p.HW;
  if (idx >= p.B * per) return;
  const int b = idx / per, r = idx - b * per;  // r = c*HW + hw, c < C
  const size_t nper = static_cast<size_t>(2 * p.C) * p.HW;
  float eps;
  if (p.has_cfg) {
    const float u = p.noise[static_cast<size_t>(b) * nper + r];
    const float t = p.noise[static_cast<size_t>(b + p.B) * nper + r];
    eps = u + p.guidance * (t - u);
  } else {
    eps = p.noise[static_cast<size_t>(b) * nper + r];
  }
  const float x = p.latents[idx];
  const float x0 = (x - p.sigma_s * eps) / p.alpha_s;
  const float prev = p.x0_prev[idx];
  p.latents[idx] = p.c_x * x + p.c_d0 * x0 + p.c_d1 * prev;
  p.x0_prev[idx] = x0;
}

}  // namespace ecadk
