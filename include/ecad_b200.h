/*
 * ecad_b200.h - C ABI of libecad_b200.so: the B200 (sm_100a) kernels behind ECAD's cached diffusion-transformer
 * forward pass.
 *
 * The reference (AniAggarwal/ecad) is 100% Python and has NO FFI; every entry point below replaces work the
 * reference reaches through torch/diffusers modules.  Each declaration cites the reference code it stands in for
 * (paths relative to the reference repo root).  INTEGRATION.md shows the ctypes binding and the reference-side
 * plug-in (ImageGeneratorRegistry / ComputeAttnRegistry) a maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no torch/C++ types cross the boundary.
 *   - All pointers are DEVICE pointers unless a comment says "host".  The caller (PyTorch) owns every buffer; the
 *     library never allocates persistent device memory.  No tensor lifetime crosses a call.
 *   - Every entry point is an asynchronous enqueue on `stream` (a cudaStream_t passed as void*).
 *   - Return 0 on success, a negative ECADK_E* code on failure; ecadk_last_error() gives the thread-local message.
 *     No C++ exception crosses the boundary.  There is no CPU fallback: on a non-sm_100 device the calls fail.
 *   - bf16 tensors are passed as void* (uint16 storage); fp32 as float*.
 */
#ifndef ECAD_B200_H_
#define ECAD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECADK_ABI_VERSION 2

#define ECADK_OK 0
#define ECADK_EINVAL (-1)  /* bad shape / alignment / null pointer */
#define ECADK_ECUDA (-2)   /* CUDA runtime or launch error */
#define ECADK_EARCH (-3)   /* device is not sm_100 */
#define ECADK_EDRIVER (-4) /* driver entry point (cuTensorMapEncodeTiled) unavailable */

#define ECADK_MAX_REUSE 12
#define ECADK_HEAD_DIM 72 /* PixArt attention_head_dim (pixart_transformer_2d_edited.py:27) */
#define ECADK_HEAD_PAD 80 /* head_dim zero-padded to a multiple of the UMMA K step (16) */

typedef void* ecadk_stream_t;
typedef struct EcadkHandle_* ecadk_handle_t;

int ecadk_abi_version(void);
const char* ecadk_last_error(void);
/* 0 if `device` is an sm_100 GPU, ECADK_EARCH otherwise. */
int ecadk_device_check(int device);

/* ------------------------------------------------------------------------------------------------------------
 * Glue kernels (HBM-bound)
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct {
  const void* cache;       /* bf16 [rows, dim] cached un-gated sub-block output */
  const float* gate_table; /* fp32 [dim] gate row of the block's scale_shift_table; NULL = no gate (attn2) */
  const float* gate_temb;  /* fp32 [samples, temb_stride], already offset to the gate chunk */
} EcadkReuse;

typedef struct {
  float* x;      /* fp32 [rows, dim] residual stream; rewritten when n_reuse > 0 */
  void* xb;      /* optional bf16 [rows, dim] copy of the (updated) stream, or NULL */
  void* h;       /* optional bf16 [rows, dim] LayerNorm+modulate output, or NULL */
  int rows;      /* samples * tokens */
  int tokens;    /* rows per sample */
  int dim;       /* 1152 (PixArt) or 3072 */
  int n_reuse;   /* 0..ECADK_MAX_REUSE */
  EcadkReuse reuse[ECADK_MAX_REUSE];
  const float* shift_table; /* fp32 [dim]  (needed when h != NULL) */
  const float* scale_table; /* fp32 [dim] */
  const float* shift_temb;  /* fp32 [samples, temb_stride], offset to the shift chunk */
  const float* scale_temb;  /* fp32 [samples, temb_stride], offset to the scale chunk */
  int temb_stride;          /* 6*dim for adaLN-single */
  float eps;                /* 1e-6 */
} EcadkResidualLnArgs;

/* Fused cached-residual reuse + LayerNorm + adaLN-single modulate.
 * Replaces: norm1/norm2 + `norm*(1+scale)+shift` (ecad/transformer_blocks/cached_transformer_block.py:208-215,
 * 306-310) and the reuse branch `gate * cached + hidden_states` (:244-246, :289, :318-320 fed by :355, :386). */
int ecadk_residual_ln(const EcadkResidualLnArgs* args, ecadk_stream_t stream);

/* Patch embedding: Conv2d(C, dim, k=2, s=2) + bias + 2-D sincos position table -> fp32 stream.
 * Replaces self.pos_embed(hidden_states) (ecad/transformer_2d_models/pixart_transformer_2d_edited.py:306).
 * latents fp32 [S,C,Hl,Wl]; wt fp32 [C*4, dim] (conv weight transposed); pos fp32 [(Hl/2)*(Wl/2), dim]. */
int ecadk_patch_embed(const float* latents, const float* wt, const float* bias, const float* pos, float* x,
                      int samples, int channels, int hl, int wl, int dim, ecadk_stream_t stream);
/* The same into a stream with `tokens_pad` >= (Hl/2)*(Wl/2) rows per sample; the padding rows are zeroed. */
int ecadk_patch_embed_padded(const float* latents, const float* wt, const float* bias, const float* pos, float* x,
                             int samples, int channels, int hl, int wl, int dim, int tokens_pad, ecadk_stream_t stream);

/* Timestep sinusoid (diffusers Timesteps(256, flip_sin_to_cos=True)): out fp32 [S, dim] = [cos | sin].
 * Replaces the first stage of self.adaln_single (pixart_transformer_2d_edited.py:308-313). */
int ecadk_timestep_sinusoid(const float* t, float* out, int samples, int dim, ecadk_stream_t stream);

/* y[s, y_off + o] (+)= b[o] + sum_i W[o,i] * act(x[s*ldx + i]); act_in: 0 none, 1 SiLU.  fp32.
 * Replaces the TimestepEmbedding MLPs (timestep, and the 1024-MS resolution / aspect-ratio embedders) and
 * AdaLayerNormSingle.linear (pixart_transformer_2d_edited.py:308-313). */
int ecadk_small_linear(const float* x, int ldx, const float* w, const float* b, float* y, int samples, int k, int o,
                       int ldy, int y_off, int act_in, int accumulate, ecadk_stream_t stream);

int ecadk_cast_f32_bf16(const float* in, void* out, size_t n, ecadk_stream_t stream);

/* bias[s,t] = (1-mask[s,t]) * -10000 for t < T, -inf for padding keys t in [T, T_pad).
 * Replaces _create_attention_mask (pixart_transformer_2d_edited.py:255-291). */
int ecadk_mask_bias(const float* mask, float* bias, int samples, int t, int t_pad, ecadk_stream_t stream);

/* Final layer: LayerNorm + (scale_shift_table[2,dim] + embedded_timestep) modulate + Linear(dim -> p*p*C) +
 * unpatchify "nhwpqc->nchpwq".  Replaces _create_output (pixart_transformer_2d_edited.py:332-376).
 * Two kernels: the fused LayerNorm+modulate (bf16 into h_scratch) and a tcgen05 GEMM whose epilogue scatters the
 * p*p*C real columns straight into the NCHW output.
 * emb fp32 [S, dim] with row pitch emb_stride (0 = one embedded timestep shared by every sample);
 * w_pad bf16 [128, dim] = proj_out.weight zero-padded from p*p*C to 128 rows; bias fp32 [128] (zero-padded);
 * h_scratch bf16 [S*hp*wp, dim]; out fp32 [S, C, 2*hp, 2*wp]. */
int ecadk_final_layer(const float* x, const float* table, const float* emb, int emb_stride, const void* w_pad,
                      const float* bias, void* h_scratch, float* out, int samples, int hp, int wp, int dim,
                      int out_channels, float eps, ecadk_stream_t stream);
/* The same with `tokens_pad` >= hp*wp rows per sample in x / h_scratch (padded token counts, see
 * EcadkBlocksArgs.self_bias): the padding rows are normalised like the rest and dropped by the unpatchify epilogue. */
int ecadk_final_layer_padded(const float* x, const float* table, const float* emb, int emb_stride, const void* w_pad,
                             const float* bias, void* h_scratch, float* out, int samples, int hp, int wp,
                             int tokens_pad, int dim, int out_channels, float eps, ecadk_stream_t stream);

/* TGATE cache averaging: buf[0:half] = (buf[0:half] + buf[half:2*half]) / 2 over bf16 elements, in place.
 * Replaces `to_cache = (hidden_uncond + hidden_pred_text) / 2` (cached_transformer_block.py:443-449). */
int ecadk_average_halves(void* buf, size_t half_elems, ecadk_stream_t stream);

/* Fused classifier-free guidance + learned-sigma drop + one DPM-Solver++(2M) update on fp32 latents.
 * Replaces the tail of the denoising loop (ecad/pipelines/pass_through.py:341-370).
 * x_next = c_x*x + c_d0*x0 + c_d1*x0_prev with x0 = (x - sigma_s*eps)/alpha_s; x0_prev is updated to x0. */
int ecadk_cfg_dpm_step(const float* noise, float* latents, float* x0_prev, int batch, int channels, int hw,
                       int has_cfg, float guidance, float sigma_s, float alpha_s, float c_x, float c_d0, float c_d1,
                       ecadk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * VAE decoder (AutoencoderKL.decode of the reference's pipelines: `image = self.vae.decode(latents /
 * self.vae.config.scaling_factor)`, ecad/pipelines/pass_through.py:382-385; the module itself is diffusers').
 * Activation layout: ZERO-BORDERED NHWC bf16 [batch, h+2, w+2, c] - the border is the padding of the 3x3
 * convolutions; every entry point below that writes such a tensor writes zeros on the border.
 * ---------------------------------------------------------------------------------------------------------- */

/* Convolution as an implicit GEMM on the tensor cores: out[pixel, :] = bias + sum_taps x[pixel + tap] * W_tap^T
 * (+ residual[pixel, :]) on interior pixels, 0 on the border.  taps = 9: 3x3, stride 1, padding 1; taps = 1: 1x1.
 * x bf16 [batch, h+2, w+2, c_in] (c_in % 64 == 0); w bf16 [c_out, taps*c_in], K index = (ky*3 + kx)*c_in + c
 * (c_out % 128 == 0: pad the rows); bias fp32 [c_out] or NULL; residual bf16, same layout as out (must not alias out), or NULL;
 * out bf16 with row pitch out_ld; only 32-column chunks that start below out_cols are written.
 * Replaces torch.nn.Conv2d inside diffusers ResnetBlock2D / Upsample2D / Decoder.conv_in / conv_out. */
int ecadk_conv_nhwc(const void* x, const void* w, const float* bias, const void* residual, void* out, int batch, int h,
                    int w_, int c_in, int c_out, int out_ld, int out_cols, int taps, ecadk_stream_t stream);

/* Upsample2D = nearest 2x + 3x3 convolution, computed WITHOUT materialising the upsampled image: an output pixel of
 * parity (a, b) sees only a 2x2 neighbourhood of the original image (kernel rows / columns that fall on the same source
 * pixel are pre-summed), so the layer is four 4-tap implicit GEMMs on x - 16 instead of 36 tap-products per input
 * pixel.  x bf16 [batch, h+2, w+2, c_in]; out bf16 [batch, 2h+2, 2w+2, c_out] (border zeroed here);
 * w4 bf16 [4 (a*2+b)][c_out][4 (ry*2+rx) * c_in + c]: for parity a the source rows are {y-1, y} (a = 0) or {y, y+1}
 * (a = 1), with W'[ry] = sum of the kernel rows ky whose upsampled row (2y + a + ky - 1) / 2 is that source row; same
 * for columns (ecad_b200/vae.py::pack_upsample_conv builds it).
 * Replaces Upsample2D.forward (F.interpolate(scale_factor=2, mode="nearest") + conv) in UpDecoderBlock2D. */
int ecadk_conv_up2x_nhwc(const void* x, const void* w4, const float* bias, void* out, int batch, int h, int w_, int c_in,
                         int c_out, ecadk_stream_t stream);

/* GroupNorm (biased variance, eps inside the sqrt) + affine (+ SiLU when silu != 0) over a bordered NHWC tensor.
 * channels per group in {4, 8, 16}.  unpadded_out > 0 writes plain tokens [batch, unpadded_out, c] instead (the input
 * of the mid-block attention; unpadded_out >= h*w is the token count per sample incl. the caller's zero padding).  scratch: >= ecadk_groupnorm_scratch_bytes(...) bytes, 16-byte aligned (per-block partial
 * sums: every reduction runs in a fixed order, so the result is bit-reproducible).
 * Replaces torch.nn.GroupNorm + SiLU (ResnetBlock2D.norm1/norm2, Decoder.conv_norm_out, Attention.group_norm). */
size_t ecadk_groupnorm_scratch_bytes(int batch, int h, int w_, int groups);
int ecadk_groupnorm_nhwc(const void* x, const float* gamma, const float* beta, void* out, void* scratch, int batch, int h,
                         int w_, int c, int groups, float eps, int silu, int unpadded_out, ecadk_stream_t stream);

/* Nearest-neighbour 2x upsampling: x [batch, h+2, w+2, c] -> out [batch, 2h+2, 2w+2, c].
 * Replaces F.interpolate(scale_factor=2, mode="nearest") in Upsample2D. */
int ecadk_upsample2x_nhwc(const void* x, void* out, int batch, int h, int w_, int c, ecadk_stream_t stream);

/* probs bf16 [rows, cols] = softmax(scale * scores fp32 [rows, cols]) over the first valid_cols columns of each row,
 * 0 in the others (padding keys); cols % 4 == 0, valid_cols % 4 == 0.
 * Replaces the softmax inside F.scaled_dot_product_attention of the single-head mid-block attention. */
int ecadk_softmax_rows(const float* scores, void* probs, int rows, int cols, int valid_cols, float scale,
                       ecadk_stream_t stream);

/* latents fp32 [batch, latent_channels, h, w] (latent_channels 4: SD / SDXL VAE; 16: FLUX VAE) ->
 * post_quant_conv(z * inv_scaling + shift) as bordered NHWC bf16 [batch, h+2, w+2, 64] (remaining channels zero: the
 * operand of conv_in).  pq_w fp32 [latent_channels, latent_channels] (out, in) and pq_b fp32 [latent_channels], or both
 * NULL when the VAE has no post_quant_conv (FLUX).
 * Replaces `latents / scaling_factor (+ shift_factor)` + AutoencoderKL.post_quant_conv. */
int ecadk_vae_prepare_latents(const float* z, const float* pq_w, const float* pq_b, float inv_scaling, float shift,
                              void* out, int batch, int latent_channels, int h, int w_, ecadk_stream_t stream);

/* out (bordered) = x (bordered) + tokens [batch, tokens_per_sample, c] (token y*w + x of a sample; tokens_per_sample
 * >= h*w): the residual connection of the mid-block attention. */
int ecadk_vae_add_tokens(const void* x, const void* tokens, int tokens_per_sample, void* out, int batch, int h, int w_,
                         int c, ecadk_stream_t stream);

/* y bf16 [batch*(h+2)*(w+2), 32] (conv_out: channels 0..2 real) -> image fp32 [batch, 3, h, w];
 * denormalize != 0 applies (x / 2 + 0.5).clamp(0, 1) (VaeImageProcessor.postprocess). */
int ecadk_vae_finish(const void* y, float* image, int batch, int h, int w_, int denormalize, ecadk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Dense contractions (tcgen05 / TMEM / TMA)
 * All GEMMs: C[M,N] = A[M,K] * W[N,K]^T + bias;  A, W bf16 row-major (K contiguous, pitch = K), bias fp32 [N].
 * Requirements: K % 64 == 0, N % 128 == 0, pointers 16-byte aligned.
 * ---------------------------------------------------------------------------------------------------------- */

/* Small problems (fewer than half as many 128-wide output tiles as SMs: the batch-1 latency configuration, BASELINE
 * config 1) run a split-K form: a thread-block cluster of 2 or 4 CTAs shares one output tile and reduces its partial
 * accumulators through an L2-resident workspace before the ordinary epilogue.  The executor entry points
 * (ecadk_pixart_blocks*, ecadk_pixart_text_kv) use a workspace owned by their handle; this call installs one for the
 * stand-alone GEMM calls of the CALLING THREAD (bytes >= tiles * 4 * 128 * BN * 4; 16 MiB covers every PixArt shape),
 * NULL / 0 removes it.  Without a workspace the GEMMs do not split.  Same results up to fp32 summation order.
 * No reference counterpart (torch.nn.Linear -> cuBLAS picks its own split-K). */
int ecadk_set_splitk_workspace(void* workspace, size_t bytes);
/* Measurements and tests: force the split-K plan to tiles of `bn` (64 | 128) columns shared by `split` (2..4) CTAs
 * wherever that plan is feasible (all clusters resident, workspace large enough; otherwise no split); 0, 0 restores
 * the planner.  Process-wide. */
int ecadk_set_splitk_force(int bn, int split);
/* Process-wide number of split-K GEMM launches so far (tests and the bench use it to show which kernel ran). */
long long ecadk_splitk_launches(void);

/* out bf16 [M, ldo].  gelu != 0 applies GELU(tanh) (diffusers GELU(approximate="tanh")).
 * Replaces FeedForward.net[0] (cached_transformer_block.py:382) and PixArtAlphaTextProjection
 * (pixart_transformer_2d_edited.py:315-321). */
int ecadk_gemm_bias(const void* a, const void* w, const float* bias, void* out, int m, int n, int k, int ldo,
                    int gelu, ecadk_stream_t stream);

/* o = A W^T + bias;  cache = bf16(o);  x += gate * o  (gate = gate_table + gate_temb[sample], or 1 if NULL);
 * optional xb = bf16(x).  x fp32 [M,N] in place.  cache may be NULL (store skipped).
 * Replaces attn.to_out[0] / ff.net[2] + "update the cache" + gated residual
 * (cached_transformer_block.py:357-358,388-389 and :244-246, :289, :318-320). */
int ecadk_gemm_bias_gated_residual_cache(const void* a, const void* w, const float* bias, float* x, void* xb,
                                         void* cache, const float* gate_table, const float* gate_temb,
                                         int temb_stride, int tokens, int m, int n, int k, ecadk_stream_t stream);

/* Q/K/V projection with head-major scatter: column c of the GEMM output belongs to part c/(heads*72)
 * (0..n_parts-1), head (c%(heads*72))/72; row r to sample r/tokens, token r%tokens.  Written to
 * out[part][sample][head][token (pitch tokens_pad)][ECADK_HEAD_PAD] bf16; padding is left untouched (keep it zero).
 * Replaces attn.to_q/to_k/to_v + the (B,L,H,d)->(B,H,L,d) reshape of AttnProcessor2_0
 * (cached_transformer_block.py:348-353). */
int ecadk_gemm_bias_headmajor(const void* a, const void* w, const float* bias, void* out0, void* out1, void* out2,
                              int n_parts, int heads, int tokens, int tokens_pad, int m, int k,
                              ecadk_stream_t stream);

/* softmax(Q K^T / sqrt(72) + bias) V.  n_keys <= 256 with q_tokens == 256: single-pass kernel (whole score row in
 * TMEM); otherwise keys are streamed in blocks of 128 with an online softmax (q_tokens % 256 == 0 required).
 * q bf16 [samples, heads, q_tokens, 80]; k, v bf16 [samples, heads, n_keys, 80] (n_keys % 128 == 0);
 * bias fp32 [samples, n_keys] or NULL; out bf16 [samples, q_tokens, heads*72].
 * Replaces F.scaled_dot_product_attention inside AttnProcessor2_0 (cached_transformer_block.py:348-353). */
int ecadk_attention(const void* q, const void* k, const void* v, const float* bias, void* out, int samples,
                    int heads, int q_tokens, int n_keys, ecadk_stream_t stream);

/* Same with the operand layout stated per operand: q_ld / kv_ld == 0 -> head-major [samples, heads, tokens, 80] as
 * above; > 0 -> ROW-major [samples*tokens, ld] bf16 - the plain output of the projection GEMM - where the pointer
 * addresses column 0 of head 0 of that operand (for a fused [M, 3*dim] q|k|v buffer: k = base + dim, v = base + 2*dim)
 * and head h occupies columns [72 h, 72 h + 72).  Row-major operands need q_tokens % 256 == 0 (the 256-query and the
 * streaming kernels gather them through 3-D tensor maps whose out-of-bounds zero fill supplies the 72 -> 80 padding). */
int ecadk_attention_ex(const void* q, int q_ld, const void* k, const void* v, int kv_ld, const float* bias, void* out,
                       int samples, int heads, int q_tokens, int n_keys, ecadk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * FLUX building blocks (kernel level; the FLUX step executor is built on them)
 * ---------------------------------------------------------------------------------------------------------- */

/* FLUX joint attention: softmax(Q K^T / sqrt(128)) V with head_dim 128 (no padding, no key bias), keys streamed in
 * blocks of 128.  q bf16 [samples, heads, q_tokens, 128] (q_tokens % 256 == 0); k, v bf16 [samples, heads, n_keys, 128]
 * (n_keys % 128 == 0).  Head h is written at columns [h*128, h*128+128) of rows with pitch out_ld.  split_tokens == 0:
 * out bf16 [samples, q_tokens, out_ld]; split_tokens > 0 (double-stream blocks, text tokens first): query rows
 * q < split_tokens go to out_lo [samples, split_tokens, out_ld], the rest to out [samples, q_tokens-split_tokens, out_ld].
 * Replaces F.scaled_dot_product_attention + the text/image split inside FluxAttnProcessor2_0
 * (ecad/transformer_blocks/cached_flux_transformer_block.py:60-66,188-192). */
int ecadk_attention_d128(const void* q, const void* k, const void* v, void* out, int out_ld, void* out_lo,
                         int split_tokens, int samples, int heads, int q_tokens, int n_keys, ecadk_stream_t stream);

/* Per-head RMSNorm (learned weight [128]) + rotary embedding, in place on head-major q and k
 * (bf16 [samples, heads, seq, 128]).  Tokens s < split use the added-stream weights (norm_added_q / norm_added_k);
 * rope_cos / rope_sin fp32 [seq, 64].  Replaces attn.norm_q/norm_k/norm_added_q/norm_added_k + apply_rope of
 * FluxAttnProcessor2_0 (diffusers 0.30.3). */
int ecadk_qk_norm_rope(void* q, void* k, const float* wq, const float* wk, const float* wq_add, const float* wk_add,
                       const float* rope_cos, const float* rope_sin, int samples, int heads, int seq, int split,
                       float eps, ecadk_stream_t stream);

/* Same with one rotation table per sample: rope_cos / rope_sin fp32 [samples, seq, 64] with `rope_sample_stride`
 * floats between samples (0 = one [seq, 64] table shared by the batch).  diffusers' EmbedND takes batched ids
 * [B, S, 3] (flux_transformer_2d_edited.py:290-291), so per-sample position ids are part of the interface. */
int ecadk_qk_norm_rope_batched(void* q, void* k, const float* wq, const float* wk, const float* wq_add,
                               const float* wk_add, const float* rope_cos, const float* rope_sin, int rope_sample_stride,
                               int samples, int heads, int seq, int split, float eps, ecadk_stream_t stream);

/* dst[r, 0:cols] = op(src[r, 0:cols]) on bf16 rows with independent pitches; op 0 = copy, 1 = GELU(tanh).
 * Replaces act_mlp(...) on the (possibly cached, pre-activation) proj_mlp output and the torch.cat of the
 * single-stream block (cached_flux_transformer_block.py:107-117). */
int ecadk_strided_unary(const void* src, void* dst, int rows, int cols, int ld_src, int ld_dst, int op,
                        ecadk_stream_t stream);

/* y += a * x (fp32): FlowMatchEulerDiscreteScheduler.step, latents += (sigma_next - sigma) * model_output. */
int ecadk_axpy_f32(float* y, const float* x, float a, size_t n, ecadk_stream_t stream);

/* out = bf16(SiLU(in)).  Replaces the nn.SiLU in front of AdaLayerNormZero / AdaLayerNormZeroSingle /
 * AdaLayerNormContinuous .linear (diffusers 0.30.3, called from cached_flux_transformer_block.py:101,234-243); the
 * result feeds ONE stacked modulation GEMM per step (every block's norm*.linear at once). */
int ecadk_silu_f32_bf16(const float* in, void* out, size_t n, ecadk_stream_t stream);

/* fp32 out[row, 0:out_cols] = A W^T + bias (row pitch ldo); W may be zero-padded to N % 128 == 0 rows.
 * Replaces x_embedder / context_embedder / proj_out of FluxTransformer2DModel
 * (ecad/transformer_2d_models/flux_transformer_2d_edited.py:275,288,317). */
int ecadk_gemm_bias_f32(const void* a, const void* w, const float* bias, float* out, int m, int n, int k, int ldo,
                        int out_cols, ecadk_stream_t stream);

/* ecadk_gemm_bias_gated_residual_cache (per-sample gate vector only) whose A operand is split along K between two
 * matrices: columns [0, k1) from a [M, k1], columns [k1, k) from a2 [M, k - k1] (k1 % 64 == 0).  Replaces
 * torch.cat([attn_output, mlp_hidden_states], dim=2) + proj_out + gate + residual of the FLUX single-stream block
 * (cached_flux_transformer_block.py:113-124) without materialising the concatenation. */
int ecadk_gemm2src_gated_residual_cache(const void* a, int k1, const void* a2, const void* w, const float* bias,
                                        float* x, void* cache, const float* gate_temb, int temb_stride, int tokens,
                                        int m, int n, int k, ecadk_stream_t stream);

/* o = A W^T + bias; out_pre = bf16(o) (may be NULL); out_gelu = bf16(GELU_tanh(o)).  Replaces proj_mlp + act_mlp of the
 * FLUX single-stream block, whose cache holds the PRE-activation (cached_flux_transformer_block.py:107-110). */
int ecadk_gemm_bias_dual(const void* a, const void* w, const float* bias, void* out_pre, void* out_gelu, int m, int n,
                         int k, int ldo_pre, int ldo_gelu, ecadk_stream_t stream);

/* ecadk_gemm_bias_headmajor with explicit head geometry and a token offset inside the head-major sequence:
 * row r -> sample r / tokens, token (r % tokens) + tok_offset of a sequence of pitch tokens_pad. */
int ecadk_gemm_bias_headmajor_ex(const void* a, const void* w, const float* bias, void* out0, void* out1, void* out2,
                                 int n_parts, int heads, int head_dim, int head_pad, int tokens, int tokens_pad,
                                 int tok_offset, int m, int k, ecadk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Step-level executor: all transformer blocks of one forward pass under one decision row.
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct {
  int num_layers;  /* 28 */
  int dim;         /* 1152 */
  int heads;       /* 16 */
  int ff_dim;      /* 4608 */
  float norm_eps;  /* 1e-6 */
} EcadkModelDesc;

typedef struct {
  const void* w_qkv1; const float* b_qkv1; /* attn1 to_q|to_k|to_v stacked: bf16 [3*dim, dim] */
  const void* w_out1; const float* b_out1; /* attn1.to_out[0]: [dim, dim] */
  const void* w_q2;   const float* b_q2;   /* attn2.to_q */
  const void* w_kv2;  const float* b_kv2;  /* attn2 to_k|to_v stacked: [2*dim, dim] */
  const void* w_out2; const float* b_out2;
  const void* w_ff1;  const float* b_ff1;  /* ff.net[0].proj: [ff_dim, dim] */
  const void* w_ff2;  const float* b_ff2;  /* ff.net[2]: [dim, ff_dim] */
  const float* scale_shift_table;          /* fp32 [6, dim]: shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp */
} EcadkBlockWeights;

/* Creates a host-side handle (weight pointer table + TMA descriptor cache); allocates no device memory.
 * `blocks` is a host array of num_layers entries (copied). */
int ecadk_create(int device, const EcadkModelDesc* desc, const EcadkBlockWeights* blocks, ecadk_handle_t* out);
int ecadk_destroy(ecadk_handle_t h);

typedef struct {
  int samples;      /* 2B with CFG */
  int tokens;       /* N image tokens per sample */
  int text_pad;     /* padded key count of the cross-attention: text tokens rounded up to a multiple of 128 */
  float* x;         /* fp32 [samples*tokens, dim] residual stream (in/out) */
  void* xb;         /* bf16 scratch [samples*tokens, dim] */
  void* h;          /* bf16 scratch [samples*tokens, dim] */
  void* q;          /* bf16 scratch [samples, heads, tokens, 80]; padding must be zero */
  void* k;
  void* v;
  void* attn_o;     /* bf16 scratch [samples*tokens, dim] */
  void* ffh;        /* bf16 scratch [samples*tokens, ff_dim] */
  const float* temb6;    /* fp32 [samples, 6*dim] adaLN-single output for this timestep */
  int temb_stride;       /* row pitch of temb6 in floats: 6*dim, or 0 when every sample shares one timestep */
  const float* text_bias;/* fp32 [samples, text_pad] */
  void* const* k2;       /* host array [num_layers]: bf16 [samples, heads, text_pad, 80] projected caption keys */
  void* const* v2;       /* host array [num_layers] */
  void* const* cache;    /* host array [num_layers*3]: bf16 [samples*tokens, dim] cached attn1/attn2/ff outputs */
  const uint8_t* cache_dead; /* host array [num_layers*3] or NULL.  cache_dead[b*3+c] != 0: the caller knows (from the
                          * schedule) that this slot is overwritten before any later step reads it - the next step
                          * recomputes the sub-block, or the generation ends - so an EXECUTED sub-block skips the cache
                          * store (the reference's `self.cached_* = out`, cached_transformer_block.py:357-358,388-389,
                          * is a dead store then).  Ignored for reused sub-blocks. */
  void* qkv;             /* bf16 [samples*tokens, 3*dim] or NULL.  When given, the q/k/v projections are written as plain
                          * row-major GEMM outputs into it and the attention kernels gather their (sample, head) tiles
                          * through 3-D tensor maps (ecadk_attention_ex); q / k / v above are then unused.  NULL keeps
                          * the head-major scatter path. */
  const float* self_bias;/* fp32 [samples, tokens] or NULL (ABI v2): additive key bias of the SELF-attention.  Token
                          * counts that are not a multiple of 256 run padded - `tokens` is the padded count, the stream
                          * and every scratch / cache buffer hold `tokens` rows per sample, the padding rows start at zero
                          * (ecadk_patch_embed_padded) and carry -10000 here (the reference's mask constant: the probability underflows
                          * to exactly 0) so that no real query attends to them. */
} EcadkBlocksArgs;

/* Runs blocks 0..num_layers-1.  executed[b*3 + c] != 0 -> compute sub-block c in {attn1, attn2, ff} of block b and
 * refresh its cache; 0 -> reuse the cached tensor (the caller has already applied the reference's
 * "flag or cache is None" rule, cached_transformer_block.py:340-347,367-373).  `executed` is a HOST array.
 * Replaces DiTScheduler.forward over the default sequential graph + CachedTransformerBlock.forward
 * (ecad/schedulers/dit_scheduler/dit_scheduler.py:50-59, cached_transformer_block.py:167-324).
 * n_launches (host, optional) receives the number of kernels enqueued. */
int ecadk_pixart_blocks(ecadk_handle_t h, const EcadkBlocksArgs* args, const uint8_t* executed, int* n_launches,
                        ecadk_stream_t stream);

/* Same for blocks [block_begin, block_end) only; `executed` / `args->cache_dead` are still indexed by absolute block
 * number.  Pending cached-residual reuses are flushed before returning, so `args->x` is complete when the call ends.
 * This is what lets a caller run a user-registered tensor-level compute function
 * (custom_compute_attn / custom_compute_ff of one block, cached_transformer_block.py:125-165) between two ranges. */
int ecadk_pixart_blocks_range(ecadk_handle_t h, const EcadkBlocksArgs* args, const uint8_t* executed, int block_begin,
                              int block_end, int* n_launches, ecadk_stream_t stream);

/* Projects caption embeddings once per generation: enc bf16 [samples*text_tokens, dim] -> k2[b], v2[b] for every
 * block (head-major, zero padding preserved).  Hoists attn2.to_k/to_v, which the reference recomputes every step
 * (cached_transformer_block.py:348-353 with encoder_hidden_states). */
int ecadk_pixart_text_kv(ecadk_handle_t h, const void* enc, int samples, int text_tokens, int text_pad,
                         void* const* k2, void* const* v2, int* n_launches, ecadk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * FLUX step-level executor: 19 double-stream + 38 single-stream blocks of one forward under one decision row.
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct {
  int num_layers;        /* 19 double-stream blocks */
  int num_single_layers; /* 38 single-stream blocks */
  int dim;               /* 3072 = heads * 128 */
  int heads;             /* 24 */
  float eps;             /* 1e-6 (LayerNorm and q/k RMSNorm) */
} EcadkFluxDesc;

typedef struct { /* FluxTransformerBlock; weights bf16 [out, in], biases / norm weights fp32 */
  const void* w_qkv;     const float* b_qkv;     /* attn.to_q|to_k|to_v stacked [3*dim, dim] (image stream) */
  const void* w_qkv_ctx; const float* b_qkv_ctx; /* attn.add_q_proj|add_k_proj|add_v_proj (text stream) */
  const void* w_out;     const float* b_out;     /* attn.to_out.0 */
  const void* w_out_ctx; const float* b_out_ctx; /* attn.to_add_out */
  const void* w_ff1;     const float* b_ff1;     /* ff.net.0.proj [4*dim, dim] */
  const void* w_ff2;     const float* b_ff2;     /* ff.net.2 [dim, 4*dim] */
  const void* w_ff1_ctx; const float* b_ff1_ctx; /* ff_context */
  const void* w_ff2_ctx; const float* b_ff2_ctx;
  const float* norm_q; const float* norm_k; const float* norm_added_q; const float* norm_added_k; /* [128] each */
} EcadkFluxDoubleWeights;

typedef struct { /* FluxSingleTransformerBlock */
  const void* w_qkv; const float* b_qkv; /* attn.to_q|to_k|to_v stacked [3*dim, dim] */
  const void* w_mlp; const float* b_mlp; /* proj_mlp [4*dim, dim] */
  const void* w_out; const float* b_out; /* proj_out [dim, 5*dim] */
  const float* norm_q; const float* norm_k;
} EcadkFluxSingleWeights;

typedef struct EcadkFluxHandle_* ecadk_flux_handle_t;
int ecadk_flux_create(int device, const EcadkFluxDesc* desc, const EcadkFluxDoubleWeights* dbl,
                      const EcadkFluxSingleWeights* sgl, ecadk_flux_handle_t* out);
int ecadk_flux_destroy(ecadk_flux_handle_t h);

typedef struct {
  int samples;     /* B (FLUX runs without classifier-free guidance) */
  int img_tokens;  /* N = (H/16)*(W/16), multiple of 32 */
  int txt_tokens;  /* T (512), multiple of 32; T + N must be a multiple of 256 */
  float* x_img;    /* fp32 [B*N, dim]  image residual stream (in: embedded latents; out: image rows after all blocks) */
  float* x_txt;    /* fp32 [B*T, dim]  text residual stream (in: context_embedder output) */
  float* x_cat;    /* fp32 scratch [B*(T+N), dim]: the concatenated stream of the single-stream phase */
  void* h_img;     /* bf16 scratch [B*N, dim] */
  void* h_txt;     /* bf16 scratch [B*T, dim] */
  void* h_cat;     /* bf16 scratch [B*(T+N), dim] */
  void* q; void* k; void* v; /* bf16 scratch [B, heads, T+N, 128] */
  void* attn_img;  /* bf16 scratch [B*N, dim] */
  void* attn_txt;  /* bf16 scratch [B*T, dim] */
  void* ffh;       /* bf16 scratch [B*max(N,T), 4*dim] */
  void* cat;       /* bf16 scratch [B*(T+N), 4*dim]: GELU(proj_mlp) of the single-stream block being executed */
  const float* mod;  /* fp32 [B, mod_stride]: all modulation vectors of this step, one row per sample.  Column layout:
                      * double block b: [b*12*dim, +6*dim) image stream (norm1.linear), then +6*dim text stream
                      * (norm1_context.linear), each = shift_msa|scale_msa|gate_msa|shift_mlp|scale_mlp|gate_mlp;
                      * single block b: [num_layers*12*dim + b*3*dim, +3*dim) = shift|scale|gate (norm.linear). */
  int mod_stride;
  const float* rope_cos;   /* fp32 [T+N, 64] (or [B, T+N, 64], see rope_sample_stride) */
  const float* rope_sin;
  void* const* cache_double; /* host array [num_layers*4]: attn [B*N,dim], context_attn [B*T,dim], ff [B*N,dim], ff_context [B*T,dim] */
  void* const* cache_single; /* host array [num_single_layers*3]: attn [B*(T+N),dim], proj_mlp (pre-GELU) [B*(T+N),4*dim], proj_out [B*(T+N),dim] */
  const uint8_t* cache_dead; /* host array, same indexing as `executed`, or NULL: slots that are overwritten before any
                              * later read.  Honoured for the epilogue side-stores (double blocks: attn pair, ff,
                              * ff_context; single blocks: proj_out); single_attn / single_proj_mlp are produced straight
                              * into their cache slots and are always written. */
  int rope_sample_stride;    /* floats between the rotation tables of two samples ([B, T+N, 64] tables), 0 = one
                              * [T+N, 64] table for the whole batch */
} EcadkFluxArgs;

/* executed[(b)*3 + c]: rows 0..num_layers-1 = double blocks with c in {full_attn, full_ff, full_ff_context}, rows
 * num_layers.. = single blocks with c in {single_attn, single_proj_mlp, single_proj_out} (FluxCacheSchedule order;
 * the caller has applied "flag or cache is None").  HOST array.
 * Replaces _dit_scheduler_forward_short_circuit + CachedFluxTransformerBlock / CachedFluxSingleTransformerBlock.forward
 * (ecad/transformer_2d_models/flux_transformer_2d_edited.py:191-218,
 *  ecad/transformer_blocks/cached_flux_transformer_block.py:99-130,228-291). */
int ecadk_flux_blocks(ecadk_flux_handle_t h, const EcadkFluxArgs* args, const uint8_t* executed, int* n_launches,
                      ecadk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * In-situ kernel timing (measurement support for bench.py; no reference counterpart - the reference only times
 * whole pipeline calls, ecad/image_generators/pixart_image_generator.py:405-439).
 * Between ecadk_profile_start() and ecadk_profile_stop() every kernel the library enqueues is bracketed by a CUDA
 * event pair on its launch stream.  stop() synchronises the device and fills one record per kernel class.
 * ---------------------------------------------------------------------------------------------------------- */
#define ECADK_PROF_GEMM 0      /* tcgen05 GEMMs, all epilogues */
#define ECADK_PROF_ATTENTION 1 /* attention kernels */
#define ECADK_PROF_GLUE 2      /* residual_ln (LayerNorm/modulate/cached-residual reuse) */
#define ECADK_PROF_OTHER 3     /* embedders, final layer, casts, solver step */
#define ECADK_PROF_CLASSES 4

typedef struct {
  long long launches;
  double total_ms;   /* sum of per-launch CUDA-event durations */
  double flops;      /* algorithmic FLOPs of those launches (GEMM: 2*M*N*K; attention: 4*Nq*Nk*72 per head) */
  double bytes;      /* algorithmic HBM bytes (glue kernels), 0 where not tracked */
} EcadkProfileRecord;

int ecadk_profile_start(void);
int ecadk_profile_stop(EcadkProfileRecord* out /* [ECADK_PROF_CLASSES] */);

/* NVTX ranges (no reference counterpart: the reference has no tracing).  While enabled, every
 * executor call (ecadk_pixart_blocks[_range], ecadk_flux_blocks) pushes one range and every EXECUTED sub-block a nested
 * one named after the reference's component - "b07.attn1" / "b07.attn2" / "b07.ff"
 * (ecad/transformer_blocks/cached_transformer_block.py:208-320), "d03.attn" / "d03.ff" / "d03.ff_context",
 * "s11.attn" / "s11.proj_mlp" / "s11.proj_out" (cached_flux_transformer_block.py:99-130,228-291) - so that a timeline,
 * or `ncu --nvtx --nvtx-include "b07.ff]"`, can be cut at that granularity.  A reused sub-block launches nothing of
 * its own (its cached residual is folded into the next kernel that reads the stream) and has no range.
 * Off by default; the environment variable ECADK_NVTX=1 turns it on at load time.  Ranges cost nothing on the device
 * and are no-ops on the host unless a tool has injected itself (NVTX v3, header-only). */
int ecadk_set_nvtx(int on);
long long ecadk_nvtx_ranges(void); /* ranges pushed so far (test / diagnostics) */

#ifdef __cplusplus
}
#endif
#endif /* ECAD_B200_H_ */
