"""CPU fp32 ORACLE of ECAD's FLUX.1 hot path.  TEST INFRASTRUCTURE - NOT PRODUCT CODE.

Restates /root/reference/ecad/transformer_blocks/cached_flux_transformer_block.py (CachedFluxSingleTransformerBlock
:12-130, CachedFluxTransformerBlock :133-291) and the model wrapper
ecad/transformer_2d_models/flux_transformer_2d_edited.py:183-326, plus the diffusers 0.30.3 module semantics they
call into (AdaLayerNormZero / ZeroSingle / Continuous, FluxAttnProcessor2_0 with per-head
RMSNorm on q,k and EmbedND RoPE, CombinedTimestepGuidanceTextProjEmbeddings, FlowMatchEulerDiscreteScheduler).

PARITY STATUS: decisions pinned (tests/test_flux_oracle.py reproduces the reference's per-step MACs of its shipped
FLUX schedules from this oracle's execution trace).  Numerics: diffusers is not installable here and the reference
ships no tensors, but the ARCHITECTURE has a second, independent implementation in this image - Black Forest Labs'
own model as vendored by `torchtitan.experiments.flux` - and tests/test_third_party_anchors.py pins this oracle to it:
the same random weights, re-keyed with diffusers' published conversion rules, give the same dense forward (rel. max
error < 2e-5) - and the shifted flow-match schedule equals BFL's `get_schedule`.  What stays recalled: the guidance
embedder (absent from that model; same MLP form as the timestep embedder) and the caching wrapper itself, which is
the reference's own code and is restated line by line.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from oracle.pixart_oracle import Trace, linear, timestep_mlp, timestep_sinusoid


@dataclass
class FluxOracleConfig:
    """Defaults = FLUX.1-dev."""

    num_attention_heads: int = 24
    attention_head_dim: int = 128
    num_layers: int = 19
    num_single_layers: int = 38
    in_channels: int = 64
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    axes_dims_rope: tuple[int, int, int] = (16, 56, 56)
    guidance_embeds: bool = True
    eps: float = 1e-6

    @property
    def inner_dim(self):
        return self.num_attention_heads * self.attention_head_dim


class FluxOracleSchedule:
    """dict schedule with the reference's step-counter semantics (cache_schedule.py:58-73); block keys "0".."18" and
    "single_0".."single_37" (flux_cache_schedule.py)."""

    FULL = ["full_attn", "full_ff", "full_ff_context"]
    SINGLE = ["single_attn", "single_proj_mlp", "single_proj_out"]

    def __init__(self, schedule: dict, num_inference_steps: int, num_blocks: int, num_single_blocks: int):
        self.schedule = {int(k): v for k, v in schedule.items()}
        self.num_inference_steps, self.num_blocks, self.num_single_blocks = (
            num_inference_steps, num_blocks, num_single_blocks)
        self._last_step = -1

    @classmethod
    def from_flags(cls, flags, num_blocks: int, num_single_blocks: int) -> "FluxOracleSchedule":
        flags = np.asarray(flags, dtype=bool)
        S = flags.shape[0]
        sched = {}
        for s in range(S):
            blocks = {}
            for b in range(num_blocks):
                blocks[str(b)] = {c: bool(flags[s, b, i]) for i, c in enumerate(cls.FULL)}
            for b in range(num_single_blocks):
                blocks[f"single_{b}"] = {c: bool(flags[s, num_blocks + b, i]) for i, c in enumerate(cls.SINGLE)}
            sched[s] = blocks
        return cls(sched, S, num_blocks, num_single_blocks)

    def reset_step(self):
        self._last_step = -1

    @property
    def curr_step(self):
        return self._last_step + 1

    def per_step_callback(self, step, timestep=None, **kw):
        self._last_step = step

    def get_recompute(self, block_num: str, component: str) -> bool:
        if component not in self.FULL + self.SINGLE:
            raise ValueError(f"Invalid component {component}.")
        return self.schedule[self.curr_step][block_num][component]


# ---- diffusers 0.30.3 pieces -------------------------------------------------------------------------------------
def rope_axis(pos: torch.Tensor, dim: int, theta: float = 10000.0) -> torch.Tensor:
    """EmbedND's `rope`: [..., n] positions -> [..., n, dim/2, 2, 2] rotation matrices (float64 angles)."""
    scale = torch.arange(0, dim, 2, dtype=torch.float64) / dim
    omega = 1.0 / (theta**scale)
    out = torch.einsum("...n,d->...nd", pos.double(), omega)
    out = torch.stack([torch.cos(out), -torch.sin(out), torch.sin(out), torch.cos(out)], dim=-1)
    return out.reshape(*out.shape[:-1], 2, 2).float()


def embed_nd(ids: torch.Tensor, axes_dim) -> torch.Tensor:
    emb = torch.cat([rope_axis(ids[..., i], axes_dim[i]) for i in range(ids.shape[-1])], dim=-3)
    return emb.unsqueeze(1)  # [B, 1, S, d/2, 2, 2]


def apply_rope(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    x_ = x.float().reshape(*x.shape[:-1], -1, 1, 2)
    out = freqs[..., 0] * x_[..., 0] + freqs[..., 1] * x_[..., 1]
    return out.reshape(*x.shape)


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    var = x.float().pow(2).mean(-1, keepdim=True)
    return x * torch.rsqrt(var + eps) * weight


def layer_norm(x, eps):
    return F.layer_norm(x, (x.shape[-1],), eps=eps)


class FluxOracle:
    def __init__(self, state_dict: dict[str, torch.Tensor], cfg: FluxOracleConfig, cache_schedule: FluxOracleSchedule):
        self.sd = {k: v.detach().float() for k, v in state_dict.items()}
        self.cfg = cfg
        self.cache_schedule = cache_schedule
        self.double_cache = [dict(attn=None, context_attn=None, ff=None, ff_context=None)
                             for _ in range(cfg.num_layers)]
        self.single_cache = [dict(attn=None, proj_mlp=None, proj_out=None) for _ in range(cfg.num_single_layers)]
        self.trace = Trace()
        self.warnings: list[str] = []

    # flux_transformer_2d_edited.py:183-189
    def reset_cache(self):
        for c in self.double_cache + self.single_cache:
            for k in c:
                c[k] = None

    def _decide(self, row: int, comp: int, block_key: str, component: str, no_cache: bool, what: str) -> bool:
        recompute = self.cache_schedule.get_recompute(block_key, component)
        if not recompute and no_cache:
            self.warnings.append(f"WARNING: No cached {what} found. Recomputing.")
        run = recompute or no_cache
        self.trace.mark(self.cache_schedule.curr_step, row, comp, run)
        return run

    # ---- attention (FluxAttnProcessor2_0 / FluxSingleAttnProcessor2_0) --------------------------------------------
    def _heads(self, t):
        B, L, _ = t.shape
        return t.view(B, L, self.cfg.num_attention_heads, self.cfg.attention_head_dim).transpose(1, 2)

    def joint_attention(self, pre: str, x, enc, rope):
        sd, eps = self.sd, self.cfg.eps
        q = rms_norm(self._heads(linear(sd, pre + ".to_q", x)), sd[pre + ".norm_q.weight"], eps)
        k = rms_norm(self._heads(linear(sd, pre + ".to_k", x)), sd[pre + ".norm_k.weight"], eps)
        v = self._heads(linear(sd, pre + ".to_v", x))
        T = 0
        if enc is not None:
            T = enc.shape[1]
            eq = rms_norm(self._heads(linear(sd, pre + ".add_q_proj", enc)), sd[pre + ".norm_added_q.weight"], eps)
            ek = rms_norm(self._heads(linear(sd, pre + ".add_k_proj", enc)), sd[pre + ".norm_added_k.weight"], eps)
            ev = self._heads(linear(sd, pre + ".add_v_proj", enc))
            q, k, v = torch.cat([eq, q], 2), torch.cat([ek, k], 2), torch.cat([ev, v], 2)
        q, k = apply_rope(q, rope), apply_rope(k, rope)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(x.shape[0], -1, self.cfg.inner_dim)
        if enc is not None:
            enc_o, o = o[:, :T], o[:, T:]
            return linear(sd, pre + ".to_out.0", o), linear(sd, pre + ".to_add_out", enc_o)
        return o

    # ---- cached_flux_transformer_block.py:228-291 ------------------------------------------------------------------
    def double_block(self, b: int, x, enc, temb, rope):
        sd, eps = self.sd, self.cfg.eps
        pre = f"transformer_blocks.{b}"
        cache = self.double_cache[b]

        def ada_zero(name, h):  # AdaLayerNormZero
            e = linear(sd, f"{pre}.{name}.linear", F.silu(temb))
            shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = e.chunk(6, dim=1)
            return layer_norm(h, eps) * (1 + scale_msa[:, None]) + shift_msa[:, None], gate_msa, shift_mlp, scale_mlp, gate_mlp

        nx, gate_msa, shift_mlp, scale_mlp, gate_mlp = ada_zero("norm1", x)
        nenc, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = ada_zero("norm1_context", enc)
        # compute_attn_cached (:170-201): the pair (attn, context_attn) is cached AFTER the output projections
        no_cache = cache["attn"] is None or cache["context_attn"] is None
        if self._decide(b, 0, str(b), "full_attn", no_cache, "attn"):
            attn_o, ctx_o = self.joint_attention(pre + ".attn", nx, nenc, rope)
        else:
            attn_o, ctx_o = cache["attn"], cache["context_attn"]
        cache["attn"], cache["context_attn"] = attn_o, ctx_o
        x = x + gate_msa.unsqueeze(1) * attn_o
        nx = layer_norm(x, eps) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
        if self._decide(b, 1, str(b), "full_ff", cache["ff"] is None, "ff"):
            ff_o = linear(sd, pre + ".ff.net.2", F.gelu(linear(sd, pre + ".ff.net.0.proj", nx), approximate="tanh"))
        else:
            ff_o = cache["ff"]
        cache["ff"] = ff_o
        x = x + gate_mlp.unsqueeze(1) * ff_o
        enc = enc + c_gate_msa.unsqueeze(1) * ctx_o
        nenc = layer_norm(enc, eps) * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
        if self._decide(b, 2, str(b), "full_ff_context", cache["ff_context"] is None, "ff_context"):
            cff = linear(sd, pre + ".ff_context.net.2",
                         F.gelu(linear(sd, pre + ".ff_context.net.0.proj", nenc), approximate="tanh"))
        else:
            cff = cache["ff_context"]
        cache["ff_context"] = cff
        enc = enc + c_gate_mlp.unsqueeze(1) * cff
        return enc, x

    # ---- cached_flux_transformer_block.py:99-130 -------------------------------------------------------------------
    def single_block(self, b: int, x, temb, rope):
        sd, eps = self.sd, self.cfg.eps
        pre = f"single_transformer_blocks.{b}"
        cache = self.single_cache[b]
        row = self.cfg.num_layers + b
        key = f"single_{b}"
        residual = x
        e = linear(sd, pre + ".norm.linear", F.silu(temb))  # AdaLayerNormZeroSingle
        shift, scale, gate = e.chunk(3, dim=1)
        nx = layer_norm(x, eps) * (1 + scale[:, None]) + shift[:, None]
        # proj_mlp is cached PRE-GELU (:107-110); component order of the schedule: attn, proj_mlp, proj_out
        if self._decide(row, 1, key, "single_proj_mlp", cache["proj_mlp"] is None, "proj_mlp"):
            mlp = linear(sd, pre + ".proj_mlp", nx)
        else:
            mlp = cache["proj_mlp"]
        cache["proj_mlp"] = mlp
        mlp_h = F.gelu(mlp, approximate="tanh")
        if self._decide(row, 0, key, "single_attn", cache["attn"] is None, "attn"):
            attn_o = self.joint_attention(pre + ".attn", nx, None, rope)
        else:
            attn_o = cache["attn"]
        cache["attn"] = attn_o
        cat = torch.cat([attn_o, mlp_h], dim=2)
        if self._decide(row, 2, key, "single_proj_out", cache["proj_out"] is None, "proj_out"):
            out = linear(sd, pre + ".proj_out", cat)
        else:
            out = cache["proj_out"]
        cache["proj_out"] = out
        return residual + gate.unsqueeze(1) * out

    # ---- flux_transformer_2d_edited.py:220-326 ---------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids,
                guidance=None):
        sd, cfg = self.sd, self.cfg
        x = linear(sd, "x_embedder", hidden_states.float())
        t = timestep.float() * 1000
        temb = timestep_mlp(sd, "time_text_embed.timestep_embedder", timestep_sinusoid(t))
        if guidance is not None:
            temb = temb + timestep_mlp(sd, "time_text_embed.guidance_embedder", timestep_sinusoid(guidance.float() * 1000))
        pooled = linear(sd, "time_text_embed.text_embedder.linear_2",
                        F.silu(linear(sd, "time_text_embed.text_embedder.linear_1", pooled_projections.float())))
        temb = temb + pooled
        enc = linear(sd, "context_embedder", encoder_hidden_states.float())
        rope = embed_nd(torch.cat((txt_ids, img_ids), dim=1), cfg.axes_dims_rope)
        for b in range(cfg.num_layers):
            enc, x = self.double_block(b, x, enc, temb, rope)
        x = torch.cat([enc, x], dim=1)
        for b in range(cfg.num_single_layers):
            x = self.single_block(b, x, temb, rope)
        x = x[:, enc.shape[1]:, ...]
        e = linear(sd, "norm_out.linear", F.silu(temb))  # AdaLayerNormContinuous: scale first, then shift
        scale, shift = e.chunk(2, dim=1)
        x = layer_norm(x, cfg.eps) * (1 + scale)[:, None, :] + shift[:, None, :]
        return linear(sd, "proj_out", x)


# ---- FlowMatchEulerDiscreteScheduler with FLUX's dynamic shifting + the denoising loop ------------------------------
def flux_sigmas(num_inference_steps: int, image_seq_len: int, base_shift: float = 0.5, max_shift: float = 1.15,
                base_image_seq_len: int = 256, max_image_seq_len: int = 4096) -> np.ndarray:
    """diffusers 0.30.3 FluxPipeline.__call__: mu = calculate_shift(seq_len, scheduler.config.base_image_seq_len,
    .max_image_seq_len, .base_shift, .max_shift) with FLUX.1-dev's scheduler_config.json (0.5 / 1.15 / 256 / 4096 -
    the defaults of FlowMatchEulerDiscreteScheduler; recalled, the reference ships no scheduler config)."""
    sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
    m = (max_shift - base_shift) / (max_image_seq_len - base_image_seq_len)
    mu = image_seq_len * m + (base_shift - m * base_image_seq_len)
    sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1))
    return np.concatenate([sigmas, [0.0]]).astype(np.float32)


@torch.no_grad()
def generate_flux_latents(model: FluxOracle, prompt_embeds, pooled, latents, img_ids, txt_ids,
                          num_inference_steps: int, guidance_scale: float = 5.0):
    """latents: packed [B, N, 64] initial noise.  Callback order as in the reference (image_generator.py:153-213)."""
    sched = model.cache_schedule
    sig = flux_sigmas(num_inference_steps, latents.shape[1])
    B = latents.shape[0]
    guidance = torch.full((B,), guidance_scale) if model.cfg.guidance_embeds else None
    for i in range(num_inference_steps):
        t = torch.full((B,), float(sig[i]))  # the pipeline passes timestep / 1000 = sigma
        v = model.forward(latents, prompt_embeds, pooled, t, img_ids, txt_ids, guidance)
        latents = latents + (float(sig[i + 1]) - float(sig[i])) * v
        sched.per_step_callback(i, float(sig[i]) * 1000)
        if i >= num_inference_steps - 1:
            sched.reset_step()
            model.reset_cache()
    return latents
