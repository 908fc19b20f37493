"""B200PixArtTransformer2D - drop-in for the reference's ``PixArtTransformer2DEdited`` on the hot path.

Same call surface as /root/reference/ecad/transformer_2d_models/pixart_transformer_2d_edited.py:
``forward(hidden_states, encoder_hidden_states, timestep, added_cond_kwargs, cross_attention_kwargs,
attention_mask, encoder_attention_mask, return_dict)`` (:160-170), ``reset_cache()`` (:155-158), attributes
``cache_schedule`` / ``dit_scheduler`` / ``config`` / ``dtype``; constructed with ``(dit_scheduler, cache_schedule)``
like ``from_pretrained(..., dit_scheduler=, cache_schedule=)`` (:104-117).  It reads ``cache_schedule.curr_step`` and
never advances it (the pipeline callback does, ecad/image_generators/image_generator.py:153-159).

Everything numerical runs in libecad_b200.so (hand-written sm_100a kernels); PyTorch only owns the buffers.  There is
no CPU path: constructing the module without a CUDA device or without the built library raises.

HBM layout (S = samples = 2B with CFG, N image tokens, D = 1152, H = 16, T text tokens padded to 128):
  x        fp32 [S*N, D]      residual stream (fp32: see DESIGN.md "precision policy")
  xb, h    bf16 [S*N, D]      bf16 shadow of x for the attn2 query projection; LN+modulate output
  q, k, v  bf16 [S,H,N,80]    head-major, head_dim 72 zero-padded to 80 (one TMA box per (sample, head))
  attn_o   bf16 [S*N, D]      attention output, A operand of the out projection
  ffh      bf16 [S*N, 4D]     GELU hidden
  cache    bf16 [L*3, S*N, D] the reference's cached_attn1/attn2/ff_output of all 28 blocks
  k2, v2   bf16 [L][S,H,Tp,80] caption keys/values (Tp = text tokens padded to 128 / 384), once per generation
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Any, Optional

import numpy as np
import torch

from . import _lib
from .registry import ComputeAttnRegistry, ComputeFFRegistry, DecisionContext
from .schedule import PixArtCacheSchedule, pixart_dead_store_mask
from .weights import PixArtConfig, random_init_state_dict



@dataclass
class Transformer2DModelOutput:
    """Stand-in for diffusers.models.modeling_outputs.Transformer2DModelOutput."""

    sample: torch.Tensor


class SequentialDiTScheduler:
    """The default DiT schedule: the blocks in order, every step.

    Mirrors the step-counter surface of /root/reference/ecad/schedulers/dit_scheduler/dit_scheduler.py:11-59.  Every
    shipped schedule uses the default sequential graph; non-default block graphs are out of
    scope and rejected by the image generator.
    """

    def __init__(self, num_inference_steps: int = 20, name: str = "default"):
        self.num_inference_steps = num_inference_steps
        self.name = name
        self._last_step = -1

    @property
    def curr_step(self) -> int:
        return self._last_step + 1

    def per_step_callback(self, step: int, timestep: Any = None, **kwargs: Any) -> None:
        self._last_step = step

    def reset_step(self) -> None:
        self._last_step = -1


def _sincos_pos_embed(dim: int, grid_hw: tuple[int, int], base_size: int, interpolation_scale: float) -> torch.Tensor:
    """2-D sin-cos table of diffusers' PatchEmbed (w-meshgrid first; each axis = concat(sin, cos))."""
    gh, gw = grid_hw
    ys = np.arange(gh, dtype=np.float32) / (gh / base_size) / interpolation_scale
    xs = np.arange(gw, dtype=np.float32) / (gw / base_size) / interpolation_scale
    col = np.tile(xs[None, :], (gh, 1)).reshape(-1)  # varies along w
    row = np.tile(ys[:, None], (1, gw)).reshape(-1)  # varies along h
    quarter = dim // 4
    omega = 1.0 / 10000 ** (np.arange(quarter, dtype=np.float64) / quarter)

    def axis(pos):
        ang = pos.astype(np.float64)[:, None] * omega[None, :]
        return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)

    return torch.from_numpy(np.concatenate([axis(col), axis(row)], axis=1)).float()


class B200BlockProxy:
    """What a tensor-level custom compute function receives as ``block`` (the reference passes the
    CachedTransformerBlock itself, cached_transformer_block.py:141-149,161-165).  Same attribute names:

      ``block_num`` (str), ``cache_schedule``, ``attn1(hidden_states, encoder_hidden_states=None, attention_mask=None)``,
      ``attn2(hidden_states, encoder_hidden_states=..., attention_mask=...)``, ``ff(hidden_states)`` - the block's own
      modules, executed by the sm_100a kernels through the C ABI (QKV projection + attention + output projection /
      GEMM + GELU + GEMM) and returning a fresh bf16 tensor ``[S, N, D]``;
      ``cached_attn1_output`` / ``cached_attn2_output`` / ``cached_ff_output`` - ``None`` while the slot is empty, else
      a bf16 view of the HBM cache slot; assigning a tensor copies it into the slot, assigning ``None`` empties it.

    ``attn2`` uses the caption keys / values the transformer projected for this generation: ``encoder_hidden_states`` and
    ``attention_mask`` are accepted for signature compatibility and must be the ones the forward was called with."""

    _SLOT = {"attn1": 0, "attn2": 1, "ff": 2}

    def __init__(self, owner: "B200PixArtTransformer2D", b: int, ws: dict, S: int, N: int, NP: int | None = None):
        # N = image tokens the functions see; NP = rows per sample of the HBM buffers (N rounded up to 256, see
        # B200PixArtTransformer2D.forward): the proxy pads operands on the way in and slices results on the way out
        self._o, self._b, self._ws, self._S, self._N = owner, b, ws, S, N
        self._NP = N if NP is None else NP
        self.ran = np.zeros(3, dtype=np.uint8)  # which of the block's own modules a function actually executed
        self.block_num = str(b)
        self.cache_schedule = owner.cache_schedule

    # ---- cache slots --------------------------------------------------------------------------------
    def _get(self, comp: str):
        c = self._SLOT[comp]
        if not self._o._has_cache[self._b, c]:
            return None
        return self._slot(c)

    def _slot(self, c: int) -> torch.Tensor:
        D = self._o.cfg.inner_dim
        return self._ws["cache"][self._b * 3 + c][: self._S * self._NP].view(self._S, self._NP, D)[:, : self._N]

    def _set(self, comp: str, value) -> None:
        c = self._SLOT[comp]
        o = self._o
        if value is None:
            o._has_cache[self._b, c] = False
            o._cache_written[self._b, c] = False
            return
        D = o.cfg.inner_dim
        slot = self._slot(c)
        if value.data_ptr() != slot.data_ptr() or value.stride() != slot.stride():
            slot.copy_(value.reshape(self._S, self._N, D))
        if self._NP != self._N:  # the padding rows of a slot are read by the reuse kernels: keep them finite
            self._ws["cache"][self._b * 3 + c][: self._S * self._NP].view(self._S, self._NP, D)[:, self._N:].zero_()
        o._has_cache[self._b, c] = True
        o._cache_written[self._b, c] = True

    cached_attn1_output = property(lambda self: self._get("attn1"), lambda self, v: self._set("attn1", v))
    cached_attn2_output = property(lambda self: self._get("attn2"), lambda self, v: self._set("attn2", v))
    cached_ff_output = property(lambda self: self._get("ff"), lambda self, v: self._set("ff", v))

    # ---- the block's modules ------------------------------------------------------------------------
    def _as_operand(self, hidden_states: torch.Tensor) -> torch.Tensor:
        D = self._o.cfg.inner_dim
        if hidden_states.shape[-1] != D or hidden_states.numel() != self._S * self._N * D:
            raise ValueError(f"expected hidden_states [{self._S}, {self._N}, {D}], got {tuple(hidden_states.shape)}")
        a = hidden_states.reshape(self._S, self._N, D).to(torch.bfloat16)
        if self._NP != self._N:
            padded = torch.zeros(self._S, self._NP, D, device=a.device, dtype=torch.bfloat16)
            padded[:, : self._N] = a
            a = padded
        return a.reshape(self._S * self._NP, D).contiguous()

    def _result(self, out: torch.Tensor) -> torch.Tensor:
        return out.view(self._S, self._NP, -1)[:, : self._N]

    def attn1(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs) -> torch.Tensor:
        if encoder_hidden_states is not None or attention_mask is not None:
            raise NotImplementedError("attn1 is plain self-attention on the PixArt path")
        o, ws, w = self._o, self._ws, self._o.blocks_w[self._b]
        S, N, H = self._S, self._NP, o.cfg.num_attention_heads
        a = self._as_operand(hidden_states)
        _lib.gemm_headmajor(a, w["w_qkv1"], w["b_qkv1"], (ws["q"], ws["k"], ws["v"]), H, N, N)
        _lib.attention(ws["q"], ws["k"], ws["v"], ws.get("self_bias"), ws["attn_o"], S, H, N, N)
        out = torch.empty(S * N, o.cfg.inner_dim, device=o.device, dtype=torch.bfloat16)
        _lib.gemm_bias(ws["attn_o"][: S * N], w["w_out1"], w["b_out1"], out)
        o.launches += 3
        self.ran[0] = 1
        return self._result(out)

    def attn2(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs) -> torch.Tensor:
        o, ws, w = self._o, self._ws, self._o.blocks_w[self._b]
        S, N, H = self._S, self._NP, o.cfg.num_attention_heads
        a = self._as_operand(hidden_states)
        _lib.gemm_headmajor(a, w["w_q2"], w["b_q2"], (ws["q"],), H, N, N)
        _lib.attention(ws["q"], ws["k2"][self._b], ws["v2"][self._b], ws["text_bias"], ws["attn_o"], S, H, N,
                       ws["text_pad"])
        out = torch.empty(S * N, o.cfg.inner_dim, device=o.device, dtype=torch.bfloat16)
        _lib.gemm_bias(ws["attn_o"][: S * N], w["w_out2"], w["b_out2"], out)
        o.launches += 3
        self.ran[1] = 1
        return self._result(out)

    def ff(self, hidden_states, **kwargs) -> torch.Tensor:
        o, ws, w = self._o, self._ws, self._o.blocks_w[self._b]
        S, N = self._S, self._NP
        a = self._as_operand(hidden_states)
        _lib.gemm_bias(a, w["w_ff1"], w["b_ff1"], ws["ffh"][: S * N], gelu=True)
        out = torch.empty(S * N, o.cfg.inner_dim, device=o.device, dtype=torch.bfloat16)
        _lib.gemm_bias(ws["ffh"][: S * N], w["w_ff2"], w["b_ff2"], out)
        o.launches += 2
        self.ran[2] = 1
        return self._result(out)


class B200PixArtTransformer2D(torch.nn.Module):
    """PixArt-alpha/sigma transformer with layer-wise feature caching, executed by libecad_b200.so.

    An ``nn.Module`` shell like the reference's model (a diffusers pipeline can hold it: ``.config``, ``.dtype``,
    ``.device``, ``.eval()``, ``.buffers()`` / ``.state_dict()`` of the packed bf16 / fp32 weights); there are no
    trainable parameters.  The C handle keeps raw device pointers to the packed weights, so ``.to()`` / ``.cuda()`` /
    ``.half()`` to another device or dtype raise instead of silently moving them."""

    def __init__(
        self,
        state_dict: dict[str, torch.Tensor],
        config: PixArtConfig = PixArtConfig(),
        dit_scheduler: SequentialDiTScheduler | None = None,
        cache_schedule: PixArtCacheSchedule | None = None,
        device: str | torch.device = "cuda:0",
    ):
        if dit_scheduler is None:
            # pixart_transformer_2d_edited.py:86-87
            raise ValueError("A DiTScheduler object must be provided.")
        if not torch.cuda.is_available():
            raise RuntimeError("B200PixArtTransformer2D needs a CUDA device; there is no CPU path")
        super().__init__()
        self.device = torch.device(device)
        self._lib = _lib.load()
        _lib.check(self._lib.ecadk_device_check(self.device.index or 0), "device_check")
        self.cfg = config
        self.config = SimpleNamespace(
            sample_size=config.sample_size, in_channels=config.in_channels, out_channels=config.out_channels,
            patch_size=config.patch_size, num_attention_heads=config.num_attention_heads,
            attention_head_dim=config.attention_head_dim, num_layers=config.num_layers,
            cross_attention_dim=config.cross_attention_dim, caption_channels=config.caption_channels,
            norm_eps=config.norm_eps,
        )
        self.dtype = torch.bfloat16
        self.dit_scheduler = dit_scheduler
        self.cache_schedule = cache_schedule if cache_schedule is not None else PixArtCacheSchedule.default(
            dit_scheduler.num_inference_steps, config.num_layers)
        if config.attention_head_dim != _lib.HEAD_DIM:
            raise ValueError("libecad_b200 is specialised for attention_head_dim = 72")
        self._pack_weights(state_dict)
        self._ws: dict[str, Any] = {}
        self._ws_key: tuple | None = None
        self._has_cache = np.zeros((config.num_layers, 3), dtype=np.bool_)
        self._text_key: tuple | None = None
        self._temb_cache: dict[float, tuple[torch.Tensor, torch.Tensor]] = {}  # timestep -> (emb, adaLN table)
        self.last_executed: np.ndarray | None = None
        # Dead-store elimination: an executed sub-block does not store its cache slot when the schedule shows that the
        # slot is overwritten (next step recomputes it) or dropped (generation ends) before anything reads it.
        # `_cache_written` tracks which slots really hold data; a reuse of an unwritten slot raises.
        self.skip_dead_cache_stores = True
        self._cache_written = np.zeros((config.num_layers, 3), dtype=np.bool_)
        self.last_dead: np.ndarray | None = None
        self.launches = 0  # kernels of libecad_b200 enqueued so far (bench.py's gpu_launches)
        # bumped whenever device buffers that recorded CUDA graphs point into are re-allocated or dropped (workspace,
        # per-timestep tables): the pipelines drop their graphs when it changes
        self.buffer_epoch = 0
        self._timestep_host_hint: float | None = None
        # tensor-signature custom compute functions (registry.TensorComputeAttnRegistry / TensorComputeFFRegistry)
        self.eval()

    def _apply(self, fn, recurse=True):  # .to() / .cuda() / .half() / .float() all funnel through here
        probe = fn(torch.empty(1, device=self.device, dtype=torch.bfloat16))
        if probe.device != self.device or probe.dtype != torch.bfloat16:
            raise RuntimeError("B200PixArtTransformer2D is bound to its device and precision policy (the C handle "
                               "holds raw pointers to the packed weights); construct it on the target device instead")
        return self

    # ------------------------------------------------------------------------------------------------
    @classmethod
    def from_random_init(cls, dit_scheduler, cache_schedule=None, config: PixArtConfig = PixArtConfig(),
                         seed: int = 0, device="cuda:0") -> "B200PixArtTransformer2D":
        return cls(random_init_state_dict(config, seed), config, dit_scheduler, cache_schedule, device)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, dit_scheduler=None, cache_schedule=None, **kwargs):
        """Same keyword surface as the reference's from_pretrained (:104-117); loads a diffusers-format state dict
        (``diffusion_pytorch_model*.safetensors`` incl. sharded checkpoints, or ``.bin``) from a local directory and
        the architecture from its ``config.json`` (``<dir>/config.json`` or ``<dir>/transformer/config.json``), so a
        512 px / 1024-MS / sigma checkpoint gets its own sample_size, interpolation scale and micro-condition
        embedders.  An explicit ``config=PixArtConfig(...)`` wins; unsupported architecture fields raise."""
        from .weights import load_diffusers_state_dict, pixart_config_from_pretrained

        sd = load_diffusers_state_dict(pretrained_model_name_or_path)
        config = kwargs.get("config")
        if config is None:
            config = pixart_config_from_pretrained(pretrained_model_name_or_path)
        return cls(sd, config, dit_scheduler, cache_schedule, kwargs.get("device", "cuda:0"))

    def _pack_weights(self, sd: dict[str, torch.Tensor]) -> None:
        dev, cfg = self.device, self.cfg
        D = cfg.inner_dim

        def f32(t):
            return t.detach().to(device=dev, dtype=torch.float32).contiguous()

        def bf16(t):
            return t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()

        w = {}
        p = cfg.patch_size
        w["patch_wt"] = f32(sd["pos_embed.proj.weight"].reshape(D, cfg.in_channels * p * p).t())
        w["patch_b"] = f32(sd["pos_embed.proj.bias"])
        for i, nm in enumerate(("linear_1", "linear_2")):
            w[f"t_w{i}"] = f32(sd[f"adaln_single.emb.timestep_embedder.{nm}.weight"])
            w[f"t_b{i}"] = f32(sd[f"adaln_single.emb.timestep_embedder.{nm}.bias"])
        if cfg.resolved_additional_conditions:  # PixArt-alpha 1024-MS: resolution + aspect-ratio embedders
            for nm in ("resolution_embedder", "aspect_ratio_embedder"):
                for i, lin in enumerate(("linear_1", "linear_2")):
                    w[f"{nm}_w{i}"] = f32(sd[f"adaln_single.emb.{nm}.{lin}.weight"])
                    w[f"{nm}_b{i}"] = f32(sd[f"adaln_single.emb.{nm}.{lin}.bias"])
        w["ada_w"] = f32(sd["adaln_single.linear.weight"])
        w["ada_b"] = f32(sd["adaln_single.linear.bias"])
        w["cap_w1"] = bf16(sd["caption_projection.linear_1.weight"])
        w["cap_b1"] = f32(sd["caption_projection.linear_1.bias"])
        w["cap_w2"] = bf16(sd["caption_projection.linear_2.weight"])
        w["cap_b2"] = f32(sd["caption_projection.linear_2.bias"])
        w["final_table"] = f32(sd["scale_shift_table"])
        # proj_out zero-padded to 128 output rows: the final layer runs on the tensor cores (ecadk_final_layer)
        n_out = sd["proj_out.weight"].shape[0]
        fw = torch.zeros(128, D, dtype=torch.float32)
        fw[:n_out] = sd["proj_out.weight"].detach().float()
        fb = torch.zeros(128, dtype=torch.float32)
        fb[:n_out] = sd["proj_out.bias"].detach().float()
        w["final_w"] = bf16(fw)
        w["final_b"] = f32(fb)
        self.w = w
        for name, t in w.items():
            self.register_buffer(f"w_{name}", t, persistent=True)
        self.blocks_w: list[dict[str, torch.Tensor]] = []
        arr = (_lib.EcadkBlockWeights * cfg.num_layers)()
        for b in range(cfg.num_layers):
            pre = f"transformer_blocks.{b}"
            bw = {
                "w_qkv1": bf16(torch.cat([sd[f"{pre}.attn1.to_{n}.weight"] for n in "qkv"], 0)),
                "b_qkv1": f32(torch.cat([sd[f"{pre}.attn1.to_{n}.bias"] for n in "qkv"], 0)),
                "w_out1": bf16(sd[f"{pre}.attn1.to_out.0.weight"]),
                "b_out1": f32(sd[f"{pre}.attn1.to_out.0.bias"]),
                "w_q2": bf16(sd[f"{pre}.attn2.to_q.weight"]),
                "b_q2": f32(sd[f"{pre}.attn2.to_q.bias"]),
                "w_kv2": bf16(torch.cat([sd[f"{pre}.attn2.to_{n}.weight"] for n in "kv"], 0)),
                "b_kv2": f32(torch.cat([sd[f"{pre}.attn2.to_{n}.bias"] for n in "kv"], 0)),
                "w_out2": bf16(sd[f"{pre}.attn2.to_out.0.weight"]),
                "b_out2": f32(sd[f"{pre}.attn2.to_out.0.bias"]),
                "w_ff1": bf16(sd[f"{pre}.ff.net.0.proj.weight"]),
                "b_ff1": f32(sd[f"{pre}.ff.net.0.proj.bias"]),
                "w_ff2": bf16(sd[f"{pre}.ff.net.2.weight"]),
                "b_ff2": f32(sd[f"{pre}.ff.net.2.bias"]),
                "scale_shift_table": f32(sd[f"{pre}.scale_shift_table"]),
            }
            self.blocks_w.append(bw)
            for name, t in bw.items():
                setattr(arr[b], name, t.data_ptr())
                self.register_buffer(f"block{b}_{name}", t, persistent=True)
        desc = _lib.EcadkModelDesc(cfg.num_layers, D, cfg.num_attention_heads, 4 * D, cfg.norm_eps)
        handle = C.c_void_p()
        _lib.check(self._lib.ecadk_create(self.device.index or 0, C.byref(desc), arr, C.byref(handle)), "create")
        self._handle = handle
        self._pos_cache: dict[tuple[int, int], torch.Tensor] = {}

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                self._lib.ecadk_destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------------------------------------
    def _pos_table(self, hp: int, wp: int) -> torch.Tensor:
        key = (hp, wp)
        if key not in self._pos_cache:
            base = self.cfg.sample_size // self.cfg.patch_size
            self._pos_cache[key] = _sincos_pos_embed(
                self.cfg.inner_dim, (hp, wp), base, self.cfg.resolved_interpolation_scale).to(self.device)
        return self._pos_cache[key]

    def _workspace(self, S: int, N: int, T: int, hl: int, wl: int) -> dict[str, Any]:
        """``N`` = rows per sample of every token buffer: the image token count rounded up to a multiple of 256 (the
        attention kernels work on 256-query items); the real count is (hl / p) * (wl / p)."""
        # Buffers are sample-major, so a forward with FEWER samples (TGATE drops the CFG pair from the gate step on,
        # ecad/pipelines/tgate.py:329-341) runs in the prefix of the same workspace and keeps every cache slot.
        if self._ws_key is not None:
            S0, N0, T0, hl0, wl0 = self._ws_key
            if (N0, T0, hl0, wl0) == (N, T, hl, wl) and S <= S0:
                self._ws["args"].samples = S
                return self._ws
        key = (S, N, T, hl, wl)
        cfg, dev = self.cfg, self.device
        D, H, L = cfg.inner_dim, cfg.num_attention_heads, cfg.num_layers
        bf, f32 = torch.bfloat16, torch.float32
        self._ws = {}  # drop the old workspace before allocating the new one
        self.buffer_epoch += 1
        ws: dict[str, Any] = {}
        M = S * N
        TEXT_PAD = ((T + 127) // 128) * 128  # 120 -> 128 (alpha), 300 -> 384 (sigma)
        ws["text_pad"] = TEXT_PAD
        ws["x"] = torch.empty(M, D, device=dev, dtype=f32)
        for n in ("xb", "h", "attn_o"):
            ws[n] = torch.empty(M, D, device=dev, dtype=bf)
        for n in ("q", "k", "v"):
            ws[n] = torch.zeros(S, H, N, _lib.HEAD_PAD, device=dev, dtype=bf)  # padding columns stay zero
        # plain [M, 3D] output of the fused q|k|v projection (and, in its first D columns, of attn2's query projection):
        # the attention kernels gather their (sample, head) tiles from it through 3-D tensor maps
        ws["qkv"] = torch.empty(M, 3 * D, device=dev, dtype=bf)
        ws["ffh"] = torch.empty(M, 4 * D, device=dev, dtype=bf)
        ws["cache"] = torch.empty(L * 3, M, D, device=dev, dtype=bf)
        ws["k2"] = torch.zeros(L, S, H, TEXT_PAD, _lib.HEAD_PAD, device=dev, dtype=bf)
        ws["v2"] = torch.zeros(L, S, H, TEXT_PAD, _lib.HEAD_PAD, device=dev, dtype=bf)
        ws["text_bias"] = torch.empty(S, TEXT_PAD, device=dev, dtype=f32)
        n_real = (hl // cfg.patch_size) * (wl // cfg.patch_size)
        if n_real != N:  # padded token count: the padding rows are masked as self-attention keys
            # -10000 (the reference's own mask constant, pixart_transformer_2d_edited.py:255-291) rather than -inf: the
            # probabilities underflow to exactly 0 either way, and a 128-key block made of padding only stays finite
            ws["self_bias"] = torch.zeros(S, N, device=dev, dtype=f32)
            ws["self_bias"][:, n_real:] = -10000.0
        ws["enc_bf"] = torch.empty(S * T, cfg.caption_channels, device=dev, dtype=bf)
        ws["enc_h"] = torch.empty(S * T, D, device=dev, dtype=bf)
        ws["enc_p"] = torch.empty(S * T, D, device=dev, dtype=bf)
        ws["t_proj"] = torch.empty(S, 256, device=dev, dtype=f32)
        ws["t_e1"] = torch.empty(S, D, device=dev, dtype=f32)
        ws["t_emb"] = torch.empty(S, D, device=dev, dtype=f32)
        ws["temb6"] = torch.empty(S, 6 * D, device=dev, dtype=f32)
        ws["out"] = torch.empty(S, cfg.out_channels, hl, wl, device=dev, dtype=f32)
        ptrs = C.c_void_p * L
        ws["k2_ptrs"] = ptrs(*[ws["k2"][b].data_ptr() for b in range(L)])
        ws["v2_ptrs"] = ptrs(*[ws["v2"][b].data_ptr() for b in range(L)])
        ws["cache_ptrs"] = (C.c_void_p * (L * 3))(*[ws["cache"][i].data_ptr() for i in range(L * 3)])
        args = _lib.EcadkBlocksArgs()
        args.samples, args.tokens, args.text_pad = S, N, TEXT_PAD
        for n in ("x", "xb", "h", "q", "k", "v", "attn_o", "ffh", "temb6", "text_bias", "qkv"):
            setattr(args, n, ws[n].data_ptr())
        args.k2 = C.cast(ws["k2_ptrs"], C.POINTER(C.c_void_p))
        args.v2 = C.cast(ws["v2_ptrs"], C.POINTER(C.c_void_p))
        args.cache = C.cast(ws["cache_ptrs"], C.POINTER(C.c_void_p))
        args.self_bias = ws["self_bias"].data_ptr() if "self_bias" in ws else None
        ws["args"] = args
        self._ws, self._ws_key = ws, key
        # a new workspace means new (empty) cache tensors
        self._has_cache[:] = False
        self._cache_written[:] = False
        self._text_key = None
        return ws

    # ------------------------------------------------------------------------------------------------
    def reset_cache(self) -> None:
        """pixart_transformer_2d_edited.py:155-158 + cached_transformer_block.py:120-123: drop every cached tensor.
        The HBM slots are kept (no allocator churn); only their validity is cleared."""
        self._has_cache[:] = False
        self._cache_written[:] = False
        self._text_key = None

    def _dead_stores(self, executed: np.ndarray) -> np.ndarray:
        """Cache stores of this step that nothing will read (ecad_b200.schedule.pixart_dead_store_mask)."""
        if not self.skip_dead_cache_stores:
            return np.zeros((self.cfg.num_layers, 3), dtype=np.uint8)
        return pixart_dead_store_mask(self.cache_schedule, self.cache_schedule.curr_step, executed, self._tgate_average)

    def _decide(self) -> np.ndarray:
        """Decision row of the current step: the reference's per-sub-block rule, evaluated through the same
        registries (cached_transformer_block.py:125-165 dispatch, :340-347 / :367-373 rule)."""
        sched = self.cache_schedule
        step = sched.curr_step
        L = self.cfg.num_layers
        executed = np.zeros((L, 3), dtype=np.uint8)
        row = sched.schedule[step]
        self._tgate_average = []  # blocks whose attn2 cache is averaged after this forward (TGATE, gate_step - 1)
        self._tensor_blocks = {}  # block -> (attn tensor fn | None, attn kwargs, ff tensor fn | None, ff kwargs)
        for b in range(L):
            entry = row[str(b)]
            attn_cfg = entry.get("custom_compute_attn", {}) or {}
            ff_cfg = entry.get("custom_compute_ff", {}) or {}
            t_attn = ComputeAttnRegistry.get_tensor(attn_cfg.get("name"))
            t_ff = ComputeFFRegistry.get_tensor(ff_cfg.get("name"))
            if t_attn is not None or t_ff is not None:
                # a user-registered tensor-level function (cached_transformer_block.py:125-165): this block is run
                # sub-block by sub-block from Python at this step; its executed flags are observed, not decided
                self._tensor_blocks[b] = (t_attn, dict(attn_cfg.get("kwargs", {}) or {}), t_ff,
                                          dict(ff_cfg.get("kwargs", {}) or {}))
                continue
            attn_fn = ComputeAttnRegistry.get(attn_cfg.get("name"), False)
            ff_fn = ComputeFFRegistry.get(ff_cfg.get("name"), False)
            if attn_fn.__name__ == "compute_attn_tgate":
                g = (attn_cfg.get("kwargs") or {}).get("gate_step")
                if g is not None and step == g - 1:
                    self._tgate_average.append(b)
            for c, comp in enumerate(("attn1", "attn2")):
                ctx = DecisionContext(b, comp, bool(sched.get_recompute(str(b), comp)), not self._has_cache[b, c],
                                      step, dict(attn_cfg.get("kwargs", {}) or {}))
                executed[b, c] = attn_fn(ctx)
            ctx = DecisionContext(b, "ff", bool(sched.get_recompute(str(b), "ff")), not self._has_cache[b, 2], step,
                                  dict(ff_cfg.get("kwargs", {}) or {}))
            executed[b, 2] = ff_fn(ctx)
        return executed

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(
        self,
        hidden_states: torch.Tensor,
        encoder_hidden_states: Optional[torch.Tensor] = None,
        timestep: Optional[torch.Tensor] = None,
        added_cond_kwargs: dict[str, torch.Tensor] | None = None,
        cross_attention_kwargs: dict[str, Any] | None = None,
        attention_mask: Optional[torch.Tensor] = None,
        encoder_attention_mask: Optional[torch.Tensor] = None,
        return_dict: bool = True,
    ):
        """Same signature as the reference (pixart_transformer_2d_edited.py:160-170).  Side channel (not an argument):
        ``hint_timestep(value)`` before the call tells the module the host value of a batch-shared timestep; the
        adaLN-single tables depend on nothing but the timestep and the weights, so with the hint they are computed once
        per distinct timestep and reused by every later step / generation without a device read-back.  The returned
        sample is a fresh tensor (the reference returns fresh tensors; the workspace buffer is reused every call)."""
        cfg, lib, w = self.cfg, self._lib, self.w
        timestep_host, self._timestep_host_hint = self._timestep_host_hint, None
        if attention_mask is not None:
            raise NotImplementedError("self-attention masks are never passed on the PixArt path")
        if encoder_hidden_states is None or timestep is None:
            raise ValueError("encoder_hidden_states and timestep are required")
        dev = self.device
        S, Cin, hl, wl = hidden_states.shape
        p = cfg.patch_size
        hp, wp = hl // p, wl // p
        n_real = hp * wp
        # The kernels work on 256-token items.  Any other token count (384 px -> 576 tokens, the non-square 1024-MS
        # buckets, ...) runs PADDED: N rows per sample in every buffer, the padding rows zero on entry, masked as
        # self-attention keys, and dropped by the unpatchify epilogue - the real rows see exactly the reference's math.
        N = -(-n_real // 256) * 256
        T = encoder_hidden_states.shape[1]
        D = cfg.inner_dim
        ws = self._workspace(S, N, T, hl, wl)
        TEXT_PAD = ws["text_pad"]
        st = _lib.stream_ptr()
        launches = 0

        # 1. input: patch embed + position table (pixart_transformer_2d_edited.py:306)
        lat = hidden_states.to(device=dev, dtype=torch.float32).contiguous()
        pos = self._pos_table(hp, wp)
        _lib.check(lib.ecadk_patch_embed_padded(lat.data_ptr(), w["patch_wt"].data_ptr(), w["patch_b"].data_ptr(),
                                                pos.data_ptr(), ws["x"].data_ptr(), S, Cin, hl, wl, D, N, st),
                   "patch_embed")
        # adaLN-single: sinusoid -> MLP -> SiLU -> Linear(D, 6D)  (:308-313)
        # The pipeline broadcasts ONE timestep over the batch (`t[None].expand(batch)`, pass_through.py:326-329):
        # a stride-0 / single-element timestep is embedded once and every kernel reads it with row pitch 0.
        addc = cfg.resolved_additional_conditions
        if addc and (added_cond_kwargs is None or added_cond_kwargs.get("resolution") is None):
            # pixart_transformer_2d_edited.py:206-209
            raise ValueError("`added_cond_kwargs` cannot be None when using additional conditions for `adaln_single`.")
        # (per-sample micro-conditions make the embedding per-sample, so the shared-timestep shortcut is off then)
        shared_t = (not addc) and (timestep.numel() == 1 or (timestep.ndim == 1 and timestep.stride(0) == 0))
        St = 1 if shared_t else S
        t32 = timestep.reshape(-1)[:1] if shared_t else timestep.reshape(-1)
        if not shared_t and t32.numel() != S:
            raise ValueError(f"timestep has {t32.numel()} entries for a batch of {S}")
        temb_stride = 0 if shared_t else 6 * D
        emb_stride = 0 if shared_t else D
        ws["args"].temb_stride = temb_stride
        cached = None
        if shared_t and timestep_host is not None:
            cached = self._temb_cache.get(float(timestep_host))
        if cached is not None:
            t_emb_buf, temb6_buf = cached
        else:
            if shared_t and timestep_host is not None:  # compute into tensors that stay alive in the cache
                t_emb_buf = torch.empty(1, D, device=dev, dtype=torch.float32)
                temb6_buf = torch.empty(1, 6 * D, device=dev, dtype=torch.float32)
                if len(self._temb_cache) >= 256:
                    self._temb_cache.clear()
                    self.buffer_epoch += 1
                self._temb_cache[float(timestep_host)] = (t_emb_buf, temb6_buf)
            else:
                t_emb_buf, temb6_buf = ws["t_emb"], ws["temb6"]
            t32 = t32.to(device=dev, dtype=torch.float32).contiguous()
            _lib.check(lib.ecadk_timestep_sinusoid(t32.data_ptr(), ws["t_proj"].data_ptr(), St, 256, st), "sinusoid")
            _lib.check(lib.ecadk_small_linear(ws["t_proj"].data_ptr(), 256, w["t_w0"].data_ptr(), w["t_b0"].data_ptr(),
                                              ws["t_e1"].data_ptr(), St, 256, D, D, 0, 0, 0, st), "t_mlp1")
            _lib.check(lib.ecadk_small_linear(ws["t_e1"].data_ptr(), D, w["t_w1"].data_ptr(), w["t_b1"].data_ptr(),
                                              t_emb_buf.data_ptr(), St, D, D, D, 0, 1, 0, st), "t_mlp2")
            if addc:
                # PixArtAlphaCombinedTimestepSizeEmbeddings: emb += cat([res_emb(h), res_emb(w), ar_emb]) with each
                # micro-condition through its own sinusoid -> Linear(256,384) -> SiLU -> Linear(384,384)
                E = D // 3
                res = added_cond_kwargs["resolution"].to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
                ar = added_cond_kwargs["aspect_ratio"].to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
                if res.numel() != 2 * S or ar.numel() != S:
                    raise ValueError("resolution must be (batch, 2) and aspect_ratio (batch, 1)")
                c_proj = torch.empty(3 * S, 256, device=dev, dtype=torch.float32)
                c_e1 = torch.empty(3 * S, E, device=dev, dtype=torch.float32)
                _lib.check(lib.ecadk_timestep_sinusoid(res.data_ptr(), c_proj.data_ptr(), 2 * S, 256, st), "res_sin")
                _lib.check(lib.ecadk_timestep_sinusoid(ar.data_ptr(), c_proj[2 * S:].data_ptr(), S, 256, st), "ar_sin")
                _lib.check(lib.ecadk_small_linear(c_proj.data_ptr(), 256, w["resolution_embedder_w0"].data_ptr(),
                                                  w["resolution_embedder_b0"].data_ptr(), c_e1.data_ptr(), 2 * S, 256, E,
                                                  E, 0, 0, 0, st), "res_mlp1")
                _lib.check(lib.ecadk_small_linear(c_proj[2 * S:].data_ptr(), 256, w["aspect_ratio_embedder_w0"].data_ptr(),
                                                  w["aspect_ratio_embedder_b0"].data_ptr(), c_e1[2 * S:].data_ptr(), S,
                                                  256, E, E, 0, 0, 0, st), "ar_mlp1")
                # second linears accumulate straight into the three thirds of the timestep embedding
                for part in range(2):  # (height, width) rows of c_e1 are interleaved per sample: row pitch 2*E
                    _lib.check(lib.ecadk_small_linear(c_e1[part:].data_ptr(), 2 * E, w["resolution_embedder_w1"].data_ptr(),
                                                      w["resolution_embedder_b1"].data_ptr(), t_emb_buf.data_ptr(), S,
                                                      E, E, D, part * E, 1, 1, st), "res_mlp2")
                _lib.check(lib.ecadk_small_linear(c_e1[2 * S:].data_ptr(), E, w["aspect_ratio_embedder_w1"].data_ptr(),
                                                  w["aspect_ratio_embedder_b1"].data_ptr(), t_emb_buf.data_ptr(), S, E,
                                                  E, D, 2 * E, 1, 1, st), "ar_mlp2")
                launches += 7
            _lib.check(lib.ecadk_small_linear(t_emb_buf.data_ptr(), D, w["ada_w"].data_ptr(), w["ada_b"].data_ptr(),
                                              temb6_buf.data_ptr(), St, D, 6 * D, 6 * D, 0, 1, 0, st), "adaln_linear")
            launches += 5
        ws["args"].temb6 = temb6_buf.data_ptr()

        # caption projection + per-block K/V: step-invariant, done once per generation (:315-321; hoisted)
        mask = encoder_attention_mask
        # identity = (pointer, shape); the projections are also redone at step 0 and after reset_cache(), so a new
        # generation never sees stale keys even if the allocator hands back the same address
        text_key = (encoder_hidden_states.data_ptr(), tuple(encoder_hidden_states.shape),
                    None if mask is None else (mask.data_ptr(), tuple(mask.shape)))
        if self._text_key != text_key or self.cache_schedule.curr_step == 0:
            enc = encoder_hidden_states.to(device=dev)
            enc_bf, enc_h, enc_p = (ws[n][: S * T] for n in ("enc_bf", "enc_h", "enc_p"))
            if enc.dtype == torch.float32:
                enc = enc.contiguous()
                _lib.check(lib.ecadk_cast_f32_bf16(enc.data_ptr(), enc_bf.data_ptr(), enc.numel(), st), "cast")
                launches += 1
            else:
                enc_bf.copy_(enc.reshape(S * T, -1))
            _lib.gemm_bias(enc_bf, w["cap_w1"], w["cap_b1"], enc_h, gelu=True)
            _lib.gemm_bias(enc_h, w["cap_w2"], w["cap_b2"], enc_p, gelu=False)
            n_l = C.c_int(0)
            _lib.check(lib.ecadk_pixart_text_kv(self._handle, ws["enc_p"].data_ptr(), S, T, TEXT_PAD,
                                                C.cast(ws["k2_ptrs"], C.POINTER(C.c_void_p)),
                                                C.cast(ws["v2_ptrs"], C.POINTER(C.c_void_p)), C.byref(n_l), st),
                       "text_kv")
            launches += 2 + n_l.value
            # mask -> additive bias (:255-291); padding keys get -inf
            if mask is None:
                ws["text_bias"][:S].zero_()
                ws["text_bias"][:S, T:] = float("-inf")
            elif mask.ndim == 2:
                m32 = mask.to(device=dev, dtype=torch.float32).contiguous()
                _lib.check(lib.ecadk_mask_bias(m32.data_ptr(), ws["text_bias"].data_ptr(), S, T, TEXT_PAD, st),
                           "mask_bias")
                launches += 1
            else:  # already a (S,1,T) bias
                ws["text_bias"][:S, :T] = mask.to(device=dev, dtype=torch.float32).reshape(S, T)
                ws["text_bias"][:S, T:] = float("-inf")
            self._text_key = text_key

        # 2. blocks under the decision row of the current step
        executed = self._decide()
        self.last_executed = executed
        ex = np.ascontiguousarray(executed.reshape(-1))
        dead = self._dead_stores(executed)
        if (~executed.astype(np.bool_) & self._has_cache & ~self._cache_written).any():
            raise RuntimeError("a cache slot whose store was skipped as dead is being reused: the schedule changed "
                               "mid-generation (set skip_dead_cache_stores = False for such flows)")
        self.last_dead = dead
        self._dead_flat = np.ascontiguousarray(dead.reshape(-1))
        ws["args"].cache_dead = self._dead_flat.ctypes.data_as(C.POINTER(C.c_uint8))
        n_l = C.c_int(0)
        if not self._tensor_blocks:
            _lib.check(lib.ecadk_pixart_blocks(self._handle, C.byref(ws["args"]),
                                               ex.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(n_l), st),
                       "pixart_blocks")
            launches += n_l.value
        else:
            # C executor for the runs of ordinary blocks, Python composition for the blocks with tensor functions
            self._tensor_ran = {}
            begin = 0
            for b in sorted(self._tensor_blocks) + [cfg.num_layers]:
                if b > begin:
                    _lib.check(lib.ecadk_pixart_blocks_range(self._handle, C.byref(ws["args"]),
                                                             ex.ctypes.data_as(C.POINTER(C.c_uint8)), begin, b,
                                                             C.byref(n_l), st), "pixart_blocks_range")
                    launches += n_l.value
                if b < cfg.num_layers:
                    l0 = self.launches
                    self._run_tensor_block(b, ws, S, n_real, N, temb6_buf, temb_stride, encoder_hidden_states, mask)
                    launches += self.launches - l0
                    self.launches = l0
                begin = b + 1
        self._has_cache |= executed.astype(np.bool_)
        self._cache_written = np.where(executed.astype(np.bool_), ~dead.astype(np.bool_), self._cache_written)
        if self._tensor_blocks:
            self.last_executed = executed.copy()
            for b, ran in self._tensor_ran.items():
                self.last_executed[b] = ran
        if self._tgate_average:
            # cached_transformer_block.py:443-449: at gate_step - 1 the cache keeps (uncond + text) / 2
            if S % 2:
                raise ValueError("TGATE averaging needs the CFG pair (an even number of samples)")
            half = (S // 2) * N * D
            for b in self._tgate_average:
                _lib.check(lib.ecadk_average_halves(ws["cache"][b * 3 + 1].data_ptr(), half, st), "average_halves")
            launches += len(self._tgate_average)

        # 3. output (:332-376)
        _lib.check(lib.ecadk_final_layer_padded(ws["x"].data_ptr(), w["final_table"].data_ptr(),
                                                t_emb_buf.data_ptr(), emb_stride, w["final_w"].data_ptr(),
                                                w["final_b"].data_ptr(), ws["h"].data_ptr(), ws["out"].data_ptr(), S,
                                                hp, wp, N, D, cfg.out_channels, cfg.norm_eps, st), "final_layer")
        launches += 2
        self.launches += launches
        out = ws["out"][:S]
        out = out.clone() if hidden_states.dtype == torch.float32 else out.to(hidden_states.dtype)
        if not return_dict:
            return (out,)
        return Transformer2DModelOutput(sample=out)

    def _run_tensor_block(self, b: int, ws: dict, S: int, N: int, NP: int, temb6: torch.Tensor, temb_stride: int,
                          encoder_hidden_states, mask) -> None:
        """One block executed sub-block by sub-block, restating CachedTransformerBlock.forward
        (cached_transformer_block.py:208-324) around user tensor functions: LN + adaLN modulate -> compute_attn("attn1")
        -> gate * out + x -> compute_attn("attn2") on the un-normalised stream -> out + x -> LN + modulate ->
        compute_ff -> gate * out + x.  Sub-blocks without a user function go through the tensor-level defaults."""
        from .registry import compute_attn_cached_tensor, compute_ff_cached_tensor

        # N tokens are what the functions see; the stream and the scratch buffers hold NP >= N rows per sample
        D, M = self.cfg.inner_dim, S * NP
        t_attn, attn_kw, t_ff, ff_kw = self._tensor_blocks[b]
        attn_fn = t_attn if t_attn is not None else compute_attn_cached_tensor
        ff_fn = t_ff if t_ff is not None else compute_ff_cached_tensor
        tab = self.blocks_w[b]["scale_shift_table"]  # [6, D] fp32: shift/scale/gate msa, shift/scale/gate mlp
        x, h, xb = ws["x"][:M], ws["h"][:M], ws["xb"][:M]
        proxy = B200BlockProxy(self, b, ws, S, N, NP)

        def as_bf16(t):
            if t.numel() != S * N * D:
                raise ValueError(f"custom compute function of block {b} returned {tuple(t.shape)}, expected [{S}, {N}, {D}]")
            return proxy._as_operand(t)  # bf16 [S * NP, D], zero padding rows

        def seen(buf):  # what a function receives: the real tokens of a [S * NP, D] buffer
            return buf.view(S, NP, D)[:, :N]

        # attn1 on LN1 + modulate (:208-246)
        _lib.residual_ln(x, NP, h=h, shift_table=tab[0], scale_table=tab[1], shift_temb=temb6[:, 0 * D:],
                         scale_temb=temb6[:, 1 * D:], temb_stride=temb_stride, eps=self.cfg.norm_eps)
        o1 = as_bf16(attn_fn(proxy, "attn1", seen(h), None, None, **attn_kw))
        # x += gate_msa * out; the bf16 shadow of the updated stream is attn2's input (:264-289: no norm for PixArt)
        _lib.residual_ln(x, NP, reuse=[(o1, tab[2], temb6[:, 2 * D:])], xb=xb, temb_stride=temb_stride,
                         eps=self.cfg.norm_eps)
        o2 = as_bf16(attn_fn(proxy, "attn2", seen(xb), encoder_hidden_states, mask, **attn_kw))
        # x += out; LN2 + modulate (:306-310)
        _lib.residual_ln(x, NP, reuse=[(o2, None, None)], h=h, shift_table=tab[3], scale_table=tab[4],
                         shift_temb=temb6[:, 3 * D:], scale_temb=temb6[:, 4 * D:], temb_stride=temb_stride,
                         eps=self.cfg.norm_eps)
        o3 = as_bf16(ff_fn(proxy, seen(h), **ff_kw))
        _lib.residual_ln(x, NP, reuse=[(o3, tab[5], temb6[:, 5 * D:])], temb_stride=temb_stride, eps=self.cfg.norm_eps)
        self.launches += 4
        # decisions are observed, not made, for this block: "executed" = the function ran the block's own module
        self._tensor_ran[b] = proxy.ran

    def hint_timestep(self, value: float | None) -> None:
        """Host value of the (batch-shared) timestep of the NEXT forward; consumed by that forward."""
        self._timestep_host_hint = None if value is None else float(value)
