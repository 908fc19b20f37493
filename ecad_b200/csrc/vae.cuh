// Memory-bound glue of the VAE decoder (`self.vae.decode(latents / scaling_factor)`,
// ecad/pipelines/pass_through.py:382-385 - diffusers AutoencoderKL, restated in oracle/vae_oracle.py).
//
// Activation layout: zero-bordered NHWC bf16, [batch, H+2, W+2, C].  The one-pixel border IS the zero padding of the
// 3x3 convolutions, which run as implicit GEMMs on the tensor cores (gemm.cuh, GemmParams::conv_taps): every kernel
// that writes such a tensor writes zeros on the border.
//
//   vae_prepare_latents_kernel  z / scaling_factor + shift_factor -> post_quant_conv (1x1, optional) -> NHWC, 4 or 16
//                               latent channels padded to 64
//   gn_stats_kernel             per (sample, block, group) sum and sum of squares, fixed-order reductions
//   gn_finalize_kernel          block partials summed in fp64 -> (mean, 1/sqrt(var + eps)) per (sample, group)
//   gn_apply_kernel             GroupNorm affine (+ SiLU) -> bordered NHWC, or -> plain [batch, H*W, C] tokens (attention)
//   upsample2x_kernel           nearest-neighbour 2x (Upsample2D before its convolution)
//   softmax_rows_kernel         fp32 scores -> bf16 probabilities (single-head 512-wide mid-block attention)
//   vae_add_tokens_kernel       bordered += tokens (residual connection of the attention block)
//   vae_finish_kernel           [M, 32] bf16 (3 real channels) -> fp32 NCHW image, optionally (x / 2 + 0.5).clamp(0, 1)
#pragma once
#include "ptx.cuh"

namespace ecadk {

struct VaePrepParams {
  const float* z;     // [B, CL, H, W] fp32 latents, CL = 4 (SD / SDXL VAE) or 16 (FLUX VAE)
  const float* pq_w;  // [CL, CL] post_quant_conv weight (out, in), or null (FLUX: use_post_quant_conv = False)
  const float* pq_b;  // [CL]
  __nv_bfloat16* out; // [B, H+2, W+2, 64]
  int B, H, W, CL;
  float inv_scaling, shift;  // z' = z * inv_scaling + shift
};
__global__ void __launch_bounds__(256) vae_prepare_latents_kernel(const VaePrepParams p) {
  // one thread per (pixel of the bordered image, 8-channel group): 8 groups of 8 channels = 64 channels
  const int plane = (p.H + 2) * (p.W + 2);
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(p.B) * plane * 8;
  if (idx >= total) return;
  const int cgp = static_cast<int>(idx & 7);
  const long long pix = idx >> 3;
  const int b = static_cast<int>(pix / plane);
  const int r = static_cast<int>(pix - static_cast<long long>(b) * plane);
  const int y = r / (p.W + 2), x = r - y * (p.W + 2);
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (cgp * 8 < p.CL && y >= 1 && y <= p.H && x >= 1 && x <= p.W) {
    float zin[16];
#pragma unroll
    for (int c = 0; c < 16; ++c)
      zin[c] = c < p.CL ? fmaf(p.z[((static_cast<size_t>(b) * p.CL + c) * p.H + (y - 1)) * p.W + (x - 1)], p.inv_scaling,
                               p.shift)
                        : 0.f;
    float zo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cgp * 8 + j;
      float a = 0.f;
      if (c < p.CL) {
        if (p.pq_w != nullptr) {
          a = p.pq_b[c];
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (k < p.CL) a = fmaf(p.pq_w[c * p.CL + k], zin[k], a);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (k == c) a = zin[k];
        }
      }
      zo[j] = a;
    }
    o.x = pack_bf16x2(zo[0], zo[1]);
    o.y = pack_bf16x2(zo[2], zo[3]);
    o.z = pack_bf16x2(zo[4], zo[5]);
    o.w = pack_bf16x2(zo[6], zo[7]);
  }
  reinterpret_cast<uint4*>(p.out)[idx] = o;
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm.  C channels, G groups of C/G in {4, 8, 16} consecutive channels; a thread always handles 8 consecutive
// channels (one uint4), i.e. two groups, one group or half a group.
struct GroupNormParams {
  const __nv_bfloat16* x;  // bordered NHWC
  __nv_bfloat16* out;      // bordered NHWC, or tokens [B, unpadded_out, C] when unpadded_out > 0
  const float* gamma;      // [C]
  const float* beta;       // [C]
  float2* partial;         // [B, blocks, G] per-block (sum, sum of squares)
  float2* mean_rstd;       // [B, G] written by gn_finalize_kernel
  int B, H, W, C, G;
  int blocks;              // gridDim.x of gn_stats_kernel
  float eps;
  int silu, unpadded_out;
};

constexpr int kGnPixelsPerBlock = 256;  // bordered pixels reduced by one block of gn_stats_kernel

// grid (ceil(plane / kGnPixelsPerBlock), B), block 256.  Border pixels are zero and add nothing.  Every reduction
// runs in a fixed order (no atomics): two decodes of the same latents are bit-identical.
__global__ void __launch_bounds__(256) gn_stats_kernel(const GroupNormParams p) {
  extern __shared__ float gn_sm[];  // [pixel sub-row][C / 4 sub-groups][2]
  const int plane = (p.H + 2) * (p.W + 2);
  const int b = blockIdx.y;
  const int slots = p.C >> 3;                   // uint4 slots per pixel
  const int pix_per_iter = blockDim.x / slots;  // 256 % slots == 0 (checked by the launcher)
  const int slot = threadIdx.x % slots;
  const int psub = threadIdx.x / slots;
  const int cpg = p.C / p.G;                    // channels per group
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;  // sub-group 0: channels 0..3 of the slot, 1: channels 4..7
  const int pix0 = blockIdx.x * kGnPixelsPerBlock;
  const int pix1 = min(plane, pix0 + kGnPixelsPerBlock);
  const uint4* xr = reinterpret_cast<const uint4*>(p.x + static_cast<size_t>(b) * plane * p.C);
  auto add = [&](const uint4 v) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
    const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
    const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.z));
    const float2 e = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.w));
    s0 += (a.x + a.y) + (c.x + c.y);
    q0 += (a.x * a.x + a.y * a.y) + (c.x * c.x + c.y * c.y);
    s1 += (d.x + d.y) + (e.x + e.y);
    q1 += (d.x * d.x + d.y * d.y) + (e.x * e.x + e.y * e.y);
  };
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int pix = pix0 + psub; pix < pix1; pix += 4 * pix_per_iter) {  // four rows in flight (fixed order of adds)
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = pix + u * pix_per_iter;
      v[u] = q < pix1 ? __ldg(xr + static_cast<size_t>(q) * slots + slot) : zero4;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) add(v[u]);
  }
  float* mine = gn_sm + (static_cast<size_t>(psub) * slots * 2 + slot * 2) * 2;
  mine[0] = s0; mine[1] = q0; mine[2] = s1; mine[3] = q1;
  __syncthreads();
  // sub-group (4 channels) -> group: one thread per group walks its sub-groups and the pixel sub-rows in order
  const int subs_per_group = cpg >> 2;
  for (int g = threadIdx.x; g < p.G; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int r = 0; r < pix_per_iter; ++r) {
      for (int j = 0; j < subs_per_group; ++j) {
        const float* v = gn_sm + (static_cast<size_t>(r) * slots * 2 + g * subs_per_group + j) * 2;
        s += v[0];
        q += v[1];
      }
    }
    p.partial[(static_cast<size_t>(b) * p.blocks + blockIdx.x) * p.G + g] = make_float2(s, q);
  }
}

// one thread per (sample, group): mean and 1/sqrt(var + eps) from the block partials, summed in fp64 in block order
// (biased variance, like torch)
__global__ void __launch_bounds__(256) gn_finalize_kernel(const GroupNormParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B * p.G) return;
  const int b = i / p.G, g = i - b * p.G;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < p.blocks; ++k) {
    const float2 v = p.partial[(static_cast<size_t>(b) * p.blocks + k) * p.G + g];
    s += static_cast<double>(v.x);
    q += static_cast<double>(v.y);
  }
  const double n = static_cast<double>(p.H) * p.W * (p.C / p.G);
  const double m = s / n;
  const double var = fmax(q / n - m * m, 0.0);
  p.mean_rstd[i] = make_float2(static_cast<float>(m), static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps))));
}

// One thread per (group of kGnApplyPix consecutive bordered pixels, 8-channel slot); grid (ceil(quads * slots / 256), B)
// keeps the index math in 32 bits.  The kernel is instruction-bound before it is HBM-bound (a first version with one
// pixel per thread, exp + IEEE division for SiLU, ran at 2.5 TB/s): the per-channel affine is folded into one FMA
// (a = rstd * gamma, b = beta - mean * a, shared by the thread's pixels) and SiLU is t * sigmoid(t) =
// 0.5 t (1 + tanh(t / 2)) with one MUFU op.
constexpr int kGnApplyPix = 4;
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__global__ void __launch_bounds__(256) gn_apply_kernel(const GroupNormParams p) {
  const int plane = (p.H + 2) * (p.W + 2);
  const int slots = p.C >> 3;
  const int quads = (plane + kGnApplyPix - 1) / kGnApplyPix;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= quads * slots) return;
  const int b = blockIdx.y;
  const int quad = i / slots;
  const int slot = i - quad * slots;
  const int c0 = slot * 8;
  const int cpg = p.C / p.G;
  float ca[8], cb[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + c0));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + c0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta + c0 + 4));
    const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 mr = __ldg(p.mean_rstd + static_cast<size_t>(b) * p.G + (c0 + 4 * h) / cpg);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ca[4 * h + j] = mr.y * ga[4 * h + j];
        cb[4 * h + j] = fmaf(-mr.x, ca[4 * h + j], be[4 * h + j]);
      }
    }
  }
  const int r0 = quad * kGnApplyPix;
  const uint4* xin = reinterpret_cast<const uint4*>(p.x) + static_cast<size_t>(b) * plane * slots;
  uint4 v[kGnApplyPix];
#pragma unroll
  for (int u = 0; u < kGnApplyPix; ++u)
    v[u] = (r0 + u < plane) ? __ldg(xin + static_cast<size_t>(r0 + u) * slots + slot) : make_uint4(0u, 0u, 0u, 0u);
  int y = r0 / (p.W + 2), x = r0 - y * (p.W + 2);
#pragma unroll
  for (int u = 0; u < kGnApplyPix; ++u) {
    const int r = r0 + u;
    if (r < plane) {
      const bool inside = y >= 1 && y <= p.H && x >= 1 && x <= p.W;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (inside) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        uint32_t ow[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
          float t0 = fmaf(f.x, ca[2 * k], cb[2 * k]);
          float t1 = fmaf(f.y, ca[2 * k + 1], cb[2 * k + 1]);
          if (p.silu) {
            const float h0 = 0.5f * t0, h1 = 0.5f * t1;
            t0 = fmaf(h0, tanh_fast(h0), h0);
            t1 = fmaf(h1, tanh_fast(h1), h1);
          }
          ow[k] = pack_bf16x2(t0, t1);
        }
        o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
      if (p.unpadded_out) {  // = tokens per sample in the output (>= H*W: rows past H*W are the caller's zero padding)
        if (inside) {
          const size_t tok = static_cast<size_t>(b) * p.unpadded_out + (y - 1) * p.W + (x - 1);
          reinterpret_cast<uint4*>(p.out)[tok * slots + slot] = o;
        }
      } else {
        reinterpret_cast<uint4*>(p.out)[(static_cast<size_t>(b) * plane + r) * slots + slot] = o;
      }
    }
    if (++x == p.W + 2) {
      x = 0;
      ++y;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// nearest 2x: out [B, 2H+2, 2W+2, C] <- in [B, H+2, W+2, C]; one thread per (output pixel, 8-channel slot)
struct UpsampleParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
  int B, H, W, C;  // input size
};
__global__ void __launch_bounds__(256) upsample2x_kernel(const UpsampleParams p) {
  const int Ho = 2 * p.H, Wo = 2 * p.W;
  const int plane_o = (Ho + 2) * (Wo + 2), plane_i = (p.H + 2) * (p.W + 2);
  const int slots = p.C >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // grid (ceil(plane_o * slots / 256), B)
  if (i >= plane_o * slots) return;
  const int b = blockIdx.y;
  const int r = i / slots;
  const int slot = i - r * slots;
  const size_t idx = static_cast<size_t>(b) * plane_o * slots + i;
  const int y = r / (Wo + 2), x = r - y * (Wo + 2);
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (y >= 1 && y <= Ho && x >= 1 && x <= Wo) {
    const int yi = ((y - 1) >> 1) + 1, xi = ((x - 1) >> 1) + 1;
    o = __ldg(reinterpret_cast<const uint4*>(p.x) +
              (static_cast<size_t>(b) * plane_i + static_cast<size_t>(yi) * (p.W + 2) + xi) * slots + slot);
  }
  reinterpret_cast<uint4*>(p.out)[idx] = o;
}

// zero the one-pixel border of a bordered NHWC tensor (the upsampling convolution writes interior pixels only);
// one thread per (border pixel, 8-channel slot), grid (ceil(border * slots / 256), B)
struct ZeroBorderParams {
  __nv_bfloat16* out;
  int B, H, W, C;
};
__global__ void __launch_bounds__(256) zero_border_kernel(const ZeroBorderParams p) {
  const int slots = p.C >> 3;
  const int border = 2 * (p.W + 2) + 2 * p.H;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= border * slots) return;
  const int k = i / slots, slot = i - k * slots;
  int y, x;
  if (k < p.W + 2) {
    y = 0; x = k;
  } else if (k < 2 * (p.W + 2)) {
    y = p.H + 1; x = k - (p.W + 2);
  } else {
    const int j = k - 2 * (p.W + 2);
    y = 1 + (j >> 1);
    x = (j & 1) ? p.W + 1 : 0;
  }
  const size_t pix = (static_cast<size_t>(blockIdx.y) * (p.H + 2) + y) * (p.W + 2) + x;
  reinterpret_cast<uint4*>(p.out)[pix * slots + slot] = make_uint4(0u, 0u, 0u, 0u);
}

// ---------------------------------------------------------------------------------------------------
// softmax over rows of fp32 scores (cols % 128 == 0, cols <= 8192): p = softmax(scale * s) -> bf16; one warp per row
// columns >= valid (padding keys; valid % 4 == 0) get probability 0
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ out,
                                                           const int rows, const int cols, const int valid,
                                                           const float scale_log2e) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* sr = reinterpret_cast<const float4*>(s + static_cast<size_t>(row) * cols);
  const int nv = valid >> 2;
  float m = -INFINITY;
  for (int i = lane; i < nv; i += 32) {
    const float4 v = sr[i];
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float mb = m * scale_log2e;
  float sum = 0.f;
  for (int i = lane; i < nv; i += 32) {
    const float4 v = sr[i];
    sum += (exp2f(fmaf(v.x, scale_log2e, -mb)) + exp2f(fmaf(v.y, scale_log2e, -mb))) +
           (exp2f(fmaf(v.z, scale_log2e, -mb)) + exp2f(fmaf(v.w, scale_log2e, -mb)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  uint2* orow = reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * cols);
  for (int i = lane; i < nv; i += 32) {
    const float4 v = sr[i];
    uint2 w;
    w.x = pack_bf16x2(exp2f(fmaf(v.x, scale_log2e, -mb)) * inv, exp2f(fmaf(v.y, scale_log2e, -mb)) * inv);
    w.y = pack_bf16x2(exp2f(fmaf(v.z, scale_log2e, -mb)) * inv, exp2f(fmaf(v.w, scale_log2e, -mb)) * inv);
    orow[i] = w;
  }
  for (int i = nv + lane; i < (cols >> 2); i += 32) orow[i] = make_uint2(0u, 0u);
}

// ---------------------------------------------------------------------------------------------------
// out(bordered) = x(bordered) + tokens([B, tokens_per_sample, C]) on interior pixels, 0 on the border
struct AddTokensParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* tokens;
  __nv_bfloat16* out;
  int B, H, W, C;
  int tokens_per_sample;  // row stride of `tokens` per sample (>= H*W)
};
__global__ void __launch_bounds__(256) vae_add_tokens_kernel(const AddTokensParams p) {
  const int plane = (p.H + 2) * (p.W + 2);
  const int slots = p.C >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // grid (ceil(plane * slots / 256), B)
  if (i >= plane * slots) return;
  const int b = blockIdx.y;
  const int r = i / slots;
  const int slot = i - r * slots;
  const size_t idx = static_cast<size_t>(b) * plane * slots + i;
  const int y = r / (p.W + 2), x = r - y * (p.W + 2);
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (y >= 1 && y <= p.H && x >= 1 && x <= p.W) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.x) + idx);
    const size_t tok = static_cast<size_t>(b) * p.tokens_per_sample + (y - 1) * p.W + (x - 1);
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p.tokens) + tok * slots + slot);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, tv[4] = {t.x, t.y, t.z, t.w};
    uint32_t ov[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&av[i]));
      const float2 ft = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&tv[i]));
      ov[i] = pack_bf16x2(fa.x + ft.x, fa.y + ft.y);
    }
    o = make_uint4(ov[0], ov[1], ov[2], ov[3]);
  }
  reinterpret_cast<uint4*>(p.out)[idx] = o;
}

// ---------------------------------------------------------------------------------------------------
// y [B*(H+2)*(W+2), 32] bf16 (channels 0..2 real) -> image fp32 [B, 3, H, W]
struct VaeFinishParams {
  const __nv_bfloat16* y;
  float* image;
  int B, H, W;
  int denormalize;  // (x / 2 + 0.5).clamp(0, 1) - VaeImageProcessor.postprocess
};
__global__ void __launch_bounds__(256) vae_finish_kernel(const VaeFinishParams p) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(p.B) * p.H * p.W;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % p.W);
  const long long t = idx / p.W;
  const int yy = static_cast<int>(t % p.H);
  const int b = static_cast<int>(t / p.H);
  const size_t row = (static_cast<size_t>(b) * (p.H + 2) + (yy + 1)) * (p.W + 2) + (x + 1);
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p.y + row * 32));
  const float2 c01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
  const float2 c23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
  float c[3] = {c01.x, c01.y, c23.x};
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float o = c[ch];
    if (p.denormalize) o = fminf(fmaxf(o * 0.5f + 0.5f, 0.f), 1.f);
    p.image[((static_cast<size_t>(b) * 3 + ch) * p.H + yy) * p.W + x] = o;
  }
}

}  // namespace ecadk
