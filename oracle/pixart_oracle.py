"""CPU fp32 ORACLE of ECAD's PixArt hot path.  TEST INFRASTRUCTURE - NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this module, and only as the checker / CPU baseline.  Nothing under ``ecad_b200/`` imports it.

PARITY STATUS
  * compute/reuse DECISIONS: pinned - the decision trace this oracle emits reproduces the per-step MAC counts
    the reference recorded in 1387 shipped PixArt schedule JSONs (tests/test_schedule_golden.py).
  * NUMERICAL outputs (latents): **parity unpinned** - the reference ships no tensors/hashes, and the arithmetic
    lives in the un-vendored third-party dependency ``diffusers==0.30.3`` (/root/reference/requirements.txt:8),
    which is not installed here and cannot be (no network).  This file restates (a) the reference's own
    forward code, which re-states BasicTransformerBlock.forward inline, and (b) the published diffusers 0.30.3
    module semantics listed in SURVEY.md Appendix A.  Each function cites what it follows.
    What CAN be anchored without diffusers is (tests/test_third_party_anchors.py): the 2-D sincos position table
    against MAE's `get_2d_sincos_pos_embed` (transformers.models.vit_mae), the timestep sinusoid against BFL's
    `timestep_embedding`, and DPM-Solver++(2M) against two analytic properties of the published algorithm (exact for
    a constant data prediction; second-order convergence to the closed-form Gaussian probability-flow solution).
    The block wiring (adaLN-single chunk order, mask -> bias, caption projection, final layer) stays recalled.

Everything is plain fp32 PyTorch on CPU; weights come in as a ``state_dict`` keyed like
``diffusers.PixArtTransformer2DModel.state_dict()`` so a real checkpoint would load unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Callable

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# schedule bookkeeping (restates ecad/schedulers/cache_scheduler/cache_schedule.py:18-73 and
# pixart_cache_schedule.py:29-37)
# ------------------------------------------------------------------------------------------------
class OracleSchedule:
    """dict-of-dict-of-dict schedule with the reference's step counter semantics."""

    COMPONENTS = ["attn1", "attn2", "ff"]

    def __init__(self, schedule: dict, num_inference_steps: int, num_blocks: int):
        self.schedule = {int(k): v for k, v in schedule.items()}  # cache_schedule.py:38-41
        self.num_inference_steps = num_inference_steps
        self.num_blocks = num_blocks
        self._last_step = -1

    @classmethod
    def from_flags(cls, flags, custom_compute_attn: dict | None = None) -> "OracleSchedule":
        flags = np.asarray(flags, dtype=bool)
        S, NB, _ = flags.shape
        sched = {}
        for s in range(S):
            sched[s] = {}
            for b in range(NB):
                e = {"attn1": bool(flags[s, b, 0]), "attn2": bool(flags[s, b, 1]), "ff": bool(flags[s, b, 2])}
                if custom_compute_attn:
                    e["custom_compute_attn"] = custom_compute_attn
                sched[s][str(b)] = e
        return cls(sched, S, NB)

    def reset_step(self):  # cache_schedule.py:58-59
        self._last_step = -1

    @property
    def curr_step(self):  # cache_schedule.py:61-63
        return self._last_step + 1

    def per_step_callback(self, step, timestep=None, **kw):  # cache_schedule.py:65-66
        self._last_step = step

    def get_recompute(self, block_num: str, component: str) -> bool:  # cache_schedule.py:68-73
        if component not in self.COMPONENTS:
            raise ValueError(f"Invalid component {component}.")
        return self.schedule[self.curr_step][block_num][component]

    def get_custom_compute_attn(self, block_num: str) -> dict:  # pixart_cache_schedule.py:29-32
        return self.schedule[self.curr_step][block_num].get("custom_compute_attn", {})

    def get_custom_compute_ff(self, block_num: str) -> dict:  # pixart_cache_schedule.py:34-37
        return self.schedule[self.curr_step][block_num].get("custom_compute_ff", {})


# ------------------------------------------------------------------------------------------------
# model config
# ------------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """Defaults = PixArt-alpha XL/2 (pixart_transformer_2d_edited.py:25-45)."""

    num_attention_heads: int = 16
    attention_head_dim: int = 72
    in_channels: int = 4
    out_channels: int = 8
    num_layers: int = 28
    cross_attention_dim: int = 1152
    sample_size: int = 32  # 32/64/128 for 256/512/1024 px
    patch_size: int = 2
    norm_eps: float = 1e-6
    caption_channels: int = 4096
    interpolation_scale: int | None = None
    use_additional_conditions: bool | None = None

    @property
    def inner_dim(self):
        return self.num_attention_heads * self.attention_head_dim

    def resolved_interpolation_scale(self):
        # diffusers PixArtTransformer2DModel.__init__: max(sample_size // 64, 1)
        return self.interpolation_scale if self.interpolation_scale is not None else max(self.sample_size // 64, 1)

    def resolved_additional_conditions(self):
        # diffusers: use_additional_conditions defaults to (sample_size == 128)
        if self.use_additional_conditions is None:
            return self.sample_size == 128
        return self.use_additional_conditions


# ------------------------------------------------------------------------------------------------
# diffusers 0.30.3 module semantics - small pure functions
# ------------------------------------------------------------------------------------------------
def sincos_1d(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    """diffusers get_1d_sincos_pos_embed_from_grid: concat(sin(p*w), cos(p*w)), w = 10000^(-i/(d/2))."""
    omega = np.arange(embed_dim // 2, dtype=np.float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000**omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(embed_dim: int, grid_hw: tuple[int, int], base_size: int, interpolation_scale: float) -> np.ndarray:
    """diffusers get_2d_sincos_pos_embed: meshgrid with w first; first half of channels encodes the column."""
    gh, gw = grid_hw
    grid_h = np.arange(gh, dtype=np.float32) / (gh / base_size) / interpolation_scale
    grid_w = np.arange(gw, dtype=np.float32) / (gw / base_size) / interpolation_scale
    grid = np.stack(np.meshgrid(grid_w, grid_h), axis=0).reshape([2, 1, gw, gh])
    emb_a = sincos_1d(embed_dim // 2, grid[0])
    emb_b = sincos_1d(embed_dim // 2, grid[1])
    return np.concatenate([emb_a, emb_b], axis=1)  # (gh*gw, D)


def timestep_sinusoid(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): concat(cos, sin)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32) / half
    ang = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def linear(sd: dict, prefix: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def timestep_mlp(sd: dict, prefix: str, proj: torch.Tensor) -> torch.Tensor:
    """diffusers TimestepEmbedding: linear_1 -> SiLU -> linear_2."""
    return linear(sd, prefix + ".linear_2", F.silu(linear(sd, prefix + ".linear_1", proj)))


def _ident(t: torch.Tensor) -> torch.Tensor:
    return t


def attention(sd: dict, prefix: str, heads: int, x: torch.Tensor, enc: torch.Tensor | None,
              bias: torch.Tensor | None, rnd: Callable = _ident) -> torch.Tensor:
    """diffusers Attention + AttnProcessor2_0: q/k/v Linear, SDPA scale 1/sqrt(d), to_out[0] Linear.
    ``rnd`` (identity in the oracle proper) marks where a reduced-precision implementation stores a tensor."""
    B, L, _ = x.shape
    src = x if enc is None else enc
    q = rnd(linear(sd, prefix + ".to_q", x))
    k = rnd(linear(sd, prefix + ".to_k", src))
    v = rnd(linear(sd, prefix + ".to_v", src))
    d = q.shape[-1] // heads
    q = q.view(B, L, heads, d).transpose(1, 2)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    mask = None
    if bias is not None:
        # prepare_attention_mask: (B,1,T) -> repeat over heads -> (B,H,1,T)
        mask = bias.repeat_interleave(heads, dim=0).view(B, heads, -1, bias.shape[-1])
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, is_causal=False)
    o = rnd(o.transpose(1, 2).reshape(B, L, heads * d))
    return linear(sd, prefix + ".to_out.0", o)


def feed_forward(sd: dict, prefix: str, x: torch.Tensor, rnd: Callable = _ident) -> torch.Tensor:
    """diffusers FeedForward('gelu-approximate'): Linear -> GELU(tanh) -> Linear."""
    h = rnd(F.gelu(linear(sd, prefix + ".net.0.proj", x), approximate="tanh"))
    return linear(sd, prefix + ".net.2", h)


# ------------------------------------------------------------------------------------------------
# the cached block (restates ecad/transformer_blocks/cached_transformer_block.py)
# ------------------------------------------------------------------------------------------------
@dataclass
class BlockCache:
    attn1: torch.Tensor | None = None
    attn2: torch.Tensor | None = None
    ff: torch.Tensor | None = None


class Trace:
    """executed[step][block][comp] as observed while running (1 = the sub-block module was called)."""

    def __init__(self):
        self._data: dict[int, dict[tuple[int, int], int]] = {}

    def mark(self, step: int, block: int, comp: int, executed: bool):
        self._data.setdefault(step, {})[(block, comp)] = int(executed)

    def clear(self):
        self._data.clear()

    def to_numpy(self, num_steps: int, num_blocks: int) -> np.ndarray:
        out = np.zeros((num_steps, num_blocks, 3), np.uint8)
        for s, d in self._data.items():
            for (b, c), v in d.items():
                out[s, b, c] = v
        return out


class OracleBlockProxy:
    """The ``block`` a registered custom compute function sees on the oracle side - the attribute surface of the
    reference's CachedTransformerBlock that such functions use (cached_transformer_block.py:116-123,141-149,161-165):
    ``block_num``, ``cache_schedule``, the modules ``attn1`` / ``attn2`` / ``ff`` and the ``cached_*_output`` slots."""

    def __init__(self, oracle: "PixArtOracle", b: int):
        self._o, self._b = oracle, b
        self.block_num = str(b)
        self.cache_schedule = oracle.cache_schedule

    def _mod(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        o = self._o
        o.trace.mark(o.cache_schedule.curr_step, self._b, 0 if attn == "attn1" else 1, True)
        return attention(o.sd, f"transformer_blocks.{self._b}.{attn}", o.cfg.num_attention_heads, hidden_states,
                         encoder_hidden_states, attention_mask, o.qa)

    def attn1(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self._mod("attn1", hidden_states, encoder_hidden_states, attention_mask)

    def attn2(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self._mod("attn2", hidden_states, encoder_hidden_states, attention_mask)

    def ff(self, hidden_states, **kw):
        o = self._o
        o.trace.mark(o.cache_schedule.curr_step, self._b, 2, True)
        return feed_forward(o.sd, f"transformer_blocks.{self._b}.ff", hidden_states, o.qa)

    cached_attn1_output = property(lambda self: self._o.caches[self._b].attn1,
                                   lambda self, v: setattr(self._o.caches[self._b], "attn1", v))
    cached_attn2_output = property(lambda self: self._o.caches[self._b].attn2,
                                   lambda self, v: setattr(self._o.caches[self._b], "attn2", v))
    cached_ff_output = property(lambda self: self._o.caches[self._b].ff,
                                lambda self, v: setattr(self._o.caches[self._b], "ff", v))


class PixArtOracle:
    """fp32 CPU restatement of PixArtTransformer2DEdited + CachedTransformerBlock.

    ``custom_attn_fns`` / ``custom_ff_fns`` play the role of ComputeAttnRegistry / ComputeFFRegistry for user-registered
    functions (custom_attn_ff.py:10-35): lower-cased name -> ``f(block, attn, hidden_states, encoder_hidden_states,
    attention_mask, **kwargs)`` resp. ``f(block, norm_hidden_states, **kwargs)`` with ``block`` an OracleBlockProxy."""

    custom_attn_fns: dict[str, Callable] = {}
    custom_ff_fns: dict[str, Callable] = {}

    def __init__(self, state_dict: dict[str, torch.Tensor], cfg: OracleConfig, cache_schedule: OracleSchedule,
                 round_act: Callable[[torch.Tensor], torch.Tensor] | None = None,
                 round_res: Callable[[torch.Tensor], torch.Tensor] | None = None):
        self.sd = {k: v.detach().to(torch.float32) for k, v in state_dict.items()}
        self.cfg = cfg
        self.cache_schedule = cache_schedule
        self.caches = [BlockCache() for _ in range(cfg.num_layers)]
        self.trace = Trace()
        self.warnings: list[str] = []
        # optional rounding hooks used ONLY by the precision study (tests/test_precision_policy.py): where a
        # reduced-precision implementation stores activations (round_act) / the residual stream (round_res).
        # Both are the identity in the oracle proper.
        self.qa = round_act if round_act is not None else _ident
        self.qr = round_res if round_res is not None else _ident
        base = cfg.sample_size // cfg.patch_size
        pe = sincos_2d(cfg.inner_dim, (base, base), base, cfg.resolved_interpolation_scale())
        self.pos_embed = torch.from_numpy(pe).float().unsqueeze(0)
        self.pos_embed_base = base

    # pixart_transformer_2d_edited.py:155-158 + cached_transformer_block.py:120-123
    def reset_cache(self):
        for c in self.caches:
            c.attn1 = c.attn2 = c.ff = None

    # ---- cached_transformer_block.py:326-360 --------------------------------------------------------
    def compute_attn_cached(self, b: int, attn: str, hidden, enc, bias):
        if attn not in ("attn1", "attn2"):
            raise ValueError(f"Invalid attention type: {attn}. Must be attn1 or attn2")
        recompute = self.cache_schedule.get_recompute(str(b), attn)
        cache = self.caches[b]
        no_cache = getattr(cache, attn) is None
        if not recompute and no_cache:
            self.warnings.append(f"WARNING: No cached {attn} found. Recomputing.")
        run = recompute or no_cache
        self.trace.mark(self.cache_schedule.curr_step, b, 0 if attn == "attn1" else 1, run)
        if run:
            out = attention(self.sd, f"transformer_blocks.{b}.{attn}", self.cfg.num_attention_heads, hidden, enc, bias,
                            self.qa)
        else:
            out = getattr(cache, attn)
        setattr(cache, attn, out)  # "update the cache" - rewritten every step
        return out

    # ---- cached_transformer_block.py:362-391 --------------------------------------------------------
    def compute_ff_cached(self, b: int, norm_hidden):
        recompute = self.cache_schedule.get_recompute(str(b), "ff")
        cache = self.caches[b]
        no_cache = cache.ff is None
        if not recompute and no_cache:
            self.warnings.append("WARNING: No cached ff found. Recomputing.")
        run = recompute or no_cache
        self.trace.mark(self.cache_schedule.curr_step, b, 2, run)
        out = feed_forward(self.sd, f"transformer_blocks.{b}.ff", norm_hidden, self.qa) if run else cache.ff
        cache.ff = out
        return out

    # ---- cached_transformer_block.py:393-454 --------------------------------------------------------
    def compute_attn_tgate(self, b: int, attn: str, hidden, enc, bias, gate_step: int | None = None):
        if attn not in ("attn1", "attn2"):
            raise ValueError(f"Invalid attention type: {attn}. Must be attn1 or attn2")
        if gate_step is None:
            raise ValueError("gate_step must be provided as a kwarg to commpute_attn_tgate.")
        if attn == "attn1":
            return self.compute_attn_cached(b, attn, hidden, enc, bias)
        cache = self.caches[b]
        step = self.cache_schedule.curr_step
        if step <= gate_step - 1:
            hidden = self.compute_attn_cached(b, attn, hidden, enc, bias)
        else:
            assert cache.attn2 is not None, "Cross-Attention must be cached at gate step for TGATE."
            self.trace.mark(step, b, 1, False)
            hidden = cache.attn2
        if step == gate_step - 1:
            uncond, text = hidden.chunk(2)
            to_cache = (uncond + text) / 2
        else:
            to_cache = hidden
        cache.attn2 = to_cache
        return hidden

    # cached_transformer_block.py:125-149 (registry dispatch, custom_attn_ff.py:22-35)
    def compute_attn(self, b: int, attn: str, hidden, enc, bias):
        cfg = self.cache_schedule.get_custom_compute_attn(str(b))
        name = (cfg.get("name") or "compute_attn_cached").lower()
        kwargs = cfg.get("kwargs", {})
        if name == "compute_attn_tgate":
            return self.compute_attn_tgate(b, attn, hidden, enc, bias, **kwargs)
        if name in self.custom_attn_fns:
            self.trace.mark(self.cache_schedule.curr_step, b, 0 if attn == "attn1" else 1, False)  # until a module runs
            return self.custom_attn_fns[name](OracleBlockProxy(self, b), attn, hidden, enc, bias, **kwargs)
        return self.compute_attn_cached(b, attn, hidden, enc, bias)

    # cached_transformer_block.py:151-165
    def compute_ff(self, b: int, norm_hidden):
        cfg = self.cache_schedule.get_custom_compute_ff(str(b))
        name = (cfg.get("name") or "compute_ff_cached").lower()
        if name in self.custom_ff_fns:
            self.trace.mark(self.cache_schedule.curr_step, b, 2, False)
            return self.custom_ff_fns[name](OracleBlockProxy(self, b), norm_hidden, **cfg.get("kwargs", {}))
        return self.compute_ff_cached(b, norm_hidden)

    # ---- cached_transformer_block.py:167-324, ada_norm_single branch --------------------------------
    def block_forward(self, b: int, hidden, enc, enc_bias, timestep6):
        qa, qr = self.qa, self.qr
        B = hidden.shape[0]
        table = self.sd[f"transformer_blocks.{b}.scale_shift_table"]
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = (
            table[None] + timestep6.reshape(B, 6, -1)
        ).chunk(6, dim=1)  # :208-211
        D = hidden.shape[-1]
        norm = F.layer_norm(hidden, (D,), eps=self.cfg.norm_eps)  # norm1, no affine
        norm = qa(norm * (1 + scale_msa) + shift_msa)  # :212-215
        attn_out = qa(self.compute_attn(b, "attn1", norm, None, None))  # :231-240
        attn_out = gate_msa * attn_out  # :244
        hidden = qr(attn_out + hidden)  # :246
        # cross-attention consumes the UN-normalised stream for ada_norm_single (:264-267)
        attn_out = qa(self.compute_attn(b, "attn2", qa(hidden), enc, enc_bias))  # :282-288
        hidden = qr(attn_out + hidden)  # :289
        norm = F.layer_norm(hidden, (D,), eps=self.cfg.norm_eps)  # norm2 (:306-307)
        norm = qa(norm * (1 + scale_mlp) + shift_mlp)  # :308-310
        ff_out = qa(self.compute_ff(b, norm))  # :313
        ff_out = gate_mlp * ff_out  # :318
        hidden = qr(ff_out + hidden)  # :320
        return hidden

    # ---- pixart_transformer_2d_edited.py:255-291 ----------------------------------------------------
    @staticmethod
    def mask_to_bias(mask: torch.Tensor | None) -> torch.Tensor | None:
        if mask is not None and mask.ndim == 2:
            mask = (1 - mask.to(torch.float32)) * -10000.0
            mask = mask.unsqueeze(1)
        return mask

    # ---- pixart_transformer_2d_edited.py:293-330 ----------------------------------------------------
    def process_input(self, latents, enc, timestep, added_cond_kwargs):
        cfg, sd = self.cfg, self.sd
        B = latents.shape[0]
        p = cfg.patch_size
        h, w = latents.shape[-2] // p, latents.shape[-1] // p
        x = F.conv2d(latents.float(), sd["pos_embed.proj.weight"], sd["pos_embed.proj.bias"], stride=p)
        x = x.flatten(2).transpose(1, 2)
        if (h, w) == (self.pos_embed_base, self.pos_embed_base):
            pe = self.pos_embed
        else:  # PatchEmbed.forward recomputes the table for a different grid
            pe = torch.from_numpy(
                sincos_2d(cfg.inner_dim, (h, w), self.pos_embed_base, cfg.resolved_interpolation_scale())
            ).float().unsqueeze(0)
        x = x + pe
        # AdaLayerNormSingle (PixArtAlphaCombinedTimestepSizeEmbeddings)
        emb = timestep_mlp(sd, "adaln_single.emb.timestep_embedder", timestep_sinusoid(timestep))
        if cfg.resolved_additional_conditions():
            res = added_cond_kwargs["resolution"].float()
            ar = added_cond_kwargs["aspect_ratio"].float()
            res_emb = timestep_mlp(sd, "adaln_single.emb.resolution_embedder", timestep_sinusoid(res.flatten()))
            ar_emb = timestep_mlp(sd, "adaln_single.emb.aspect_ratio_embedder", timestep_sinusoid(ar.flatten()))
            emb = emb + torch.cat([res_emb.reshape(B, -1), ar_emb.reshape(B, -1)], dim=1)
        timestep6 = linear(sd, "adaln_single.linear", F.silu(emb))
        # caption projection: Linear -> GELU(tanh) -> Linear (:315-321)
        enc = linear(sd, "caption_projection.linear_2",
                     F.gelu(linear(sd, "caption_projection.linear_1", enc.float()), approximate="tanh"))
        enc = enc.view(B, -1, x.shape[-1])
        return h, w, x, enc, timestep6, emb

    # ---- pixart_transformer_2d_edited.py:332-376 ----------------------------------------------------
    def create_output(self, hidden, embedded_timestep, h, w):
        cfg, sd = self.cfg, self.sd
        shift, scale = (sd["scale_shift_table"][None] + embedded_timestep[:, None]).chunk(2, dim=1)
        hidden = F.layer_norm(hidden, (hidden.shape[-1],), eps=cfg.norm_eps)
        hidden = hidden * (1 + scale) + shift
        hidden = linear(sd, "proj_out", hidden)
        p, c = cfg.patch_size, cfg.out_channels
        hidden = hidden.reshape(-1, h, w, p, p, c)
        hidden = torch.einsum("nhwpqc->nchpwq", hidden)
        return hidden.reshape(-1, c, h * p, w * p)

    # ---- pixart_transformer_2d_edited.py:160-253 ----------------------------------------------------
    @torch.no_grad()
    def forward(self, hidden_states, encoder_hidden_states, timestep, added_cond_kwargs=None,
                encoder_attention_mask=None):
        if self.cfg.resolved_additional_conditions() and added_cond_kwargs is None:
            raise ValueError("`added_cond_kwargs` cannot be None when using additional conditions for `adaln_single`.")
        enc_bias = self.mask_to_bias(encoder_attention_mask)
        h, w, x, enc, timestep6, emb = self.process_input(
            hidden_states, encoder_hidden_states, timestep, added_cond_kwargs)
        x, enc = self.qr(x), self.qa(enc)
        # default DiT schedule == the 28 blocks in order (dit_scheduler.py:50-59, pixart_builder.py:96-124)
        for b in range(self.cfg.num_layers):
            x = self.block_forward(b, x, enc, enc_bias, timestep6)
        return self.create_output(x, emb, h, w)


# ------------------------------------------------------------------------------------------------
# DPM-Solver++(2M) - restates diffusers 0.30.3 DPMSolverMultistepScheduler with PixArt's scheduler_config
# (dpmsolver++, order 2, midpoint, epsilon prediction, linear betas 1e-4..0.02 x1000, linspace spacing,
#  lower_order_final, final sigma zero).  SURVEY.md Appendix A "Scheduler".
# ------------------------------------------------------------------------------------------------
class OracleDPMSolver:
    def __init__(self, num_inference_steps: int, num_train_timesteps: int = 1000,
                 beta_start: float = 1e-4, beta_end: float = 0.02):
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        sig = (((1 - alphas_cumprod) / alphas_cumprod) ** 0.5).numpy()
        ts = np.linspace(0, num_train_timesteps - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.step_index = 0
        self.lower_order_nums = 0
        self.outputs: list[torch.Tensor | None] = [None, None]

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma**2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def step(self, model_output: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        i = self.step_index
        n = len(self.timesteps)
        lower_order_final = i == n - 1  # final_sigmas_type == "zero"
        lower_order_second = (i == n - 2) and n < 15
        # convert_model_output: epsilon -> x0
        alpha_s0, sigma_s0 = self._alpha_sigma(self.sigmas[i])
        x0 = (sample - sigma_s0 * model_output) / alpha_s0
        self.outputs = [self.outputs[1], x0]
        sample = sample.float()
        sigma_t = self.sigmas[i + 1]
        alpha_t, sig_t = self._alpha_sigma(sigma_t)
        lambda_t = torch.log(alpha_t) - torch.log(sig_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        if self.lower_order_nums < 1 or lower_order_final:
            prev = (sig_t / sigma_s0) * sample - (alpha_t * (torch.exp(-h) - 1.0)) * x0
        else:
            alpha_s1, sigma_s1 = self._alpha_sigma(self.sigmas[i - 1])
            lambda_s1 = torch.log(alpha_s1) - torch.log(sigma_s1)
            m0, m1 = self.outputs[1], self.outputs[0]
            h_0 = lambda_s0 - lambda_s1
            r0 = h_0 / h
            D0, D1 = m0, (1.0 / r0) * (m0 - m1)
            prev = ((sig_t / sigma_s0) * sample - (alpha_t * (torch.exp(-h) - 1.0)) * D0
                    - 0.5 * (alpha_t * (torch.exp(-h) - 1.0)) * D1)
        _ = lower_order_second  # order-2 solver: the "second" rule coincides with the 2M update
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev


# ------------------------------------------------------------------------------------------------
# the denoising loop + callbacks (restates ecad/pipelines/pass_through.py:238-380 and
# ecad/image_generators/image_generator.py:153-213, pixart_image_generator.py:349-383)
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def generate_latents(model: PixArtOracle, prompt_embeds, prompt_mask, negative_embeds, negative_mask,
                     latents: torch.Tensor, num_inference_steps: int, guidance_scale: float = 4.5,
                     tgate_gate_step: int | None = None, record_steps: bool = False) -> dict[str, Any]:
    """One generation.  ``latents`` is the caller-drawn initial noise (B,4,h,w) (init_noise_sigma = 1)."""
    sched = model.cache_schedule
    B = latents.shape[0]
    cfg = model.cfg
    do_cfg = guidance_scale > 1.0
    embeds = torch.cat([negative_embeds, prompt_embeds], dim=0) if do_cfg else prompt_embeds
    mask = torch.cat([negative_mask, prompt_mask], dim=0) if do_cfg else prompt_mask
    solver = OracleDPMSolver(num_inference_steps)
    added = {"resolution": None, "aspect_ratio": None}
    if cfg.sample_size == 128:
        hh, ww = latents.shape[-2] * 8, latents.shape[-1] * 8
        res = torch.tensor([hh, ww], dtype=torch.float32).repeat(B, 1)
        ar = torch.tensor([float(hh / ww)]).repeat(B, 1)
        if do_cfg:
            res, ar = torch.cat([res, res]), torch.cat([ar, ar])
        added = {"resolution": res, "aspect_ratio": ar}
    per_step = []
    for i, t in enumerate(solver.timesteps):
        gated = tgate_gate_step is not None and i >= tgate_gate_step  # ecad/pipelines/tgate.py:329-341
        if gated:
            x_in, e_in, m_in = latents, negative_embeds, negative_mask
            a_in = {k: (v[:B] if v is not None else None) for k, v in added.items()}
        else:
            x_in = torch.cat([latents] * 2) if do_cfg else latents
            e_in, m_in, a_in = embeds, mask, added
        ts = t[None].expand(x_in.shape[0])
        noise = model.forward(x_in, e_in, ts, a_in, m_in)
        if do_cfg and not gated:
            uncond, text = noise.chunk(2)
            noise = uncond + guidance_scale * (text - uncond)
        if cfg.out_channels // 2 == cfg.in_channels:  # learned sigma dropped
            noise = noise.chunk(2, dim=1)[0]
        latents = solver.step(noise, latents)
        if record_steps:
            per_step.append(latents.clone())
        # callbacks in the reference's order: step counters, then reset on the last step
        sched.per_step_callback(i, int(t))
        if i >= num_inference_steps - 1:
            sched.reset_step()
            model.reset_cache()
    return {"latents": latents, "per_step": per_step}
