"""Random-init weights of the named architecture (BASELINE.json: "random-init weights of the named architecture").

There is no network for checkpoints, so the protocol is the one the reference itself uses for MAC counting on
FLUX (``skip_transformer_block_init=True``, /root/reference/ecad/transformer_2d_models/flux_transformer_2d_edited.py:75-83):
construct the architecture and keep its constructor initialisation.  Keys follow
``diffusers.PixArtTransformer2DModel.state_dict()`` so a real PixArt-alpha/sigma checkpoint loads through the
same path; init scales are the constructor's (nn.Linear / nn.Conv2d Kaiming-uniform(a=sqrt 5) => U(+-1/sqrt(fan_in))
for weight and bias, ``scale_shift_table = randn/sqrt(D)``), SURVEY.md section 8d.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class PixArtConfig:
    """Mirror of the reference's constructor defaults (pixart_transformer_2d_edited.py:25-45)."""

    num_attention_heads: int = 16
    attention_head_dim: int = 72
    in_channels: int = 4
    out_channels: int = 8
    num_layers: int = 28
    cross_attention_dim: int = 1152
    sample_size: int = 32
    patch_size: int = 2
    norm_eps: float = 1e-6
    caption_channels: int = 4096
    interpolation_scale: int | None = None
    use_additional_conditions: bool | None = None

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def resolved_interpolation_scale(self) -> int:
        return self.interpolation_scale if self.interpolation_scale is not None else max(self.sample_size // 64, 1)

    @property
    def resolved_additional_conditions(self) -> bool:
        if self.use_additional_conditions is None:
            return self.sample_size == 128
        return self.use_additional_conditions


# Hugging Face names a schedule JSON may carry in ``config.transformer_weights`` (the reference passes them to
# ``from_pretrained``, image_generator.py:172-186) -> the architecture they denote.  There are no checkpoints offline:
# the name selects the CONFIG a random-init model is built with (sample size -> position-table base size and
# interpolation scale; micro-condition embedders only on PixArt-alpha 1024-MS).
KNOWN_PIXART_WEIGHTS: dict[str, PixArtConfig] = {
    "PixArt-alpha/PixArt-XL-2-256x256": PixArtConfig(sample_size=32),
    "PixArt-alpha/PixArt-XL-2-512x512": PixArtConfig(sample_size=64),
    "PixArt-alpha/PixArt-XL-2-1024-MS": PixArtConfig(sample_size=128),
    "PixArt-alpha/PixArt-Sigma-XL-2-256x256": PixArtConfig(sample_size=32, use_additional_conditions=False),
    "PixArt-alpha/PixArt-Sigma-XL-2-512-MS": PixArtConfig(sample_size=64, use_additional_conditions=False),
    "PixArt-alpha/PixArt-Sigma-XL-2-1024-MS": PixArtConfig(sample_size=128, use_additional_conditions=False),
}


def _linear(sd, name, out_f, in_f, gen, fan_in=None):
    bound = 1.0 / math.sqrt(fan_in or in_f)
    sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    sd[name + ".bias"] = (torch.rand(out_f, generator=gen) * 2 - 1) * bound


def random_init_state_dict(cfg: PixArtConfig = PixArtConfig(), seed: int = 0) -> dict[str, torch.Tensor]:
    """fp32 CPU state dict, deterministic in ``seed``."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    D = cfg.inner_dim
    sd: dict[str, torch.Tensor] = {}
    p = cfg.patch_size
    fan = cfg.in_channels * p * p
    b = 1.0 / math.sqrt(fan)
    sd["pos_embed.proj.weight"] = (torch.rand(D, cfg.in_channels, p, p, generator=gen) * 2 - 1) * b
    sd["pos_embed.proj.bias"] = (torch.rand(D, generator=gen) * 2 - 1) * b
    _linear(sd, "adaln_single.emb.timestep_embedder.linear_1", D, 256, gen)
    _linear(sd, "adaln_single.emb.timestep_embedder.linear_2", D, D, gen)
    if cfg.resolved_additional_conditions:
        s = D // 3
        for nm in ("resolution_embedder", "aspect_ratio_embedder"):
            _linear(sd, f"adaln_single.emb.{nm}.linear_1", s, 256, gen)
            _linear(sd, f"adaln_single.emb.{nm}.linear_2", s, s, gen)
    _linear(sd, "adaln_single.linear", 6 * D, D, gen)
    _linear(sd, "caption_projection.linear_1", D, cfg.caption_channels, gen)
    _linear(sd, "caption_projection.linear_2", D, D, gen)
    for i in range(cfg.num_layers):
        pre = f"transformer_blocks.{i}"
        sd[pre + ".scale_shift_table"] = torch.randn(6, D, generator=gen) / D**0.5
        for attn, kv_in in (("attn1", D), ("attn2", cfg.cross_attention_dim)):
            _linear(sd, f"{pre}.{attn}.to_q", D, D, gen)
            _linear(sd, f"{pre}.{attn}.to_k", D, kv_in, gen)
            _linear(sd, f"{pre}.{attn}.to_v", D, kv_in, gen)
            _linear(sd, f"{pre}.{attn}.to_out.0", D, D, gen)
        _linear(sd, f"{pre}.ff.net.0.proj", 4 * D, D, gen)
        _linear(sd, f"{pre}.ff.net.2", D, 4 * D, gen)
    sd["scale_shift_table"] = torch.randn(2, D, generator=gen) / D**0.5
    _linear(sd, "proj_out", p * p * cfg.out_channels, D, gen)
    return sd


def synthetic_prompt_embeddings(batch: int, text_tokens: int = 120, channels: int = 4096, seed: int = 1,
                                dtype: torch.dtype = torch.float32) -> dict[str, torch.Tensor]:
    """Synthetic T5-like caption embeddings in the reference's ``PixArtPromptEmbedding`` layout
    (/root/reference/ecad/types.py:14-18): randn*0.2, per-prompt valid length U{8..T}, and ONE null embedding
    (valid length 2) repeated over the batch like the reference's "" embedding
    (/root/reference/ecad/image_generators/pixart_image_generator.py:237-242)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    emb = torch.randn(batch, text_tokens, channels, generator=gen) * 0.2
    lens = torch.randint(8, text_tokens + 1, (batch,), generator=gen)
    mask = (torch.arange(text_tokens)[None, :] < lens[:, None]).to(torch.int64)
    neg = (torch.randn(1, text_tokens, channels, generator=gen) * 0.2).repeat(batch, 1, 1)
    neg_mask = (torch.arange(text_tokens)[None, :] < 2).to(torch.int64).repeat(batch, 1)
    return {
        "prompt_embeds": emb.to(dtype),
        "prompt_attention_mask": mask,
        "negative_prompt_embeds": neg.to(dtype),
        "negative_prompt_attention_mask": neg_mask,
    }


@dataclass(frozen=True)
class FluxConfig:
    """FLUX.1-dev architecture (diffusers FluxTransformer2DModel defaults)."""

    num_attention_heads: int = 24
    attention_head_dim: int = 128
    num_layers: int = 19
    num_single_layers: int = 38
    in_channels: int = 64
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    axes_dims_rope: tuple[int, int, int] = (16, 56, 56)
    guidance_embeds: bool = True

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


def flux_random_init_state_dict(cfg: FluxConfig = FluxConfig(), seed: int = 0) -> dict[str, torch.Tensor]:
    """Constructor-initialised FLUX weights keyed like diffusers' FluxTransformer2DModel.state_dict()
    (the reference's own random-init protocol: skip_transformer_block_init,
    /root/reference/ecad/transformer_2d_models/flux_transformer_2d_edited.py:75-83)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    D, hd = cfg.inner_dim, cfg.attention_head_dim
    sd: dict[str, torch.Tensor] = {}
    _linear(sd, "x_embedder", D, cfg.in_channels, gen)
    _linear(sd, "context_embedder", D, cfg.joint_attention_dim, gen)
    embs = ["timestep_embedder"] + (["guidance_embedder"] if cfg.guidance_embeds else [])
    for nm in embs:
        _linear(sd, f"time_text_embed.{nm}.linear_1", D, 256, gen)
        _linear(sd, f"time_text_embed.{nm}.linear_2", D, D, gen)
    _linear(sd, "time_text_embed.text_embedder.linear_1", D, cfg.pooled_projection_dim, gen)
    _linear(sd, "time_text_embed.text_embedder.linear_2", D, D, gen)
    for i in range(cfg.num_layers):
        pre = f"transformer_blocks.{i}"
        _linear(sd, f"{pre}.norm1.linear", 6 * D, D, gen)
        _linear(sd, f"{pre}.norm1_context.linear", 6 * D, D, gen)
        for nm in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
            _linear(sd, f"{pre}.attn.{nm}", D, D, gen)
        for nm in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            sd[f"{pre}.attn.{nm}.weight"] = torch.ones(hd)
        for ff in ("ff", "ff_context"):
            _linear(sd, f"{pre}.{ff}.net.0.proj", 4 * D, D, gen)
            _linear(sd, f"{pre}.{ff}.net.2", D, 4 * D, gen)
    for i in range(cfg.num_single_layers):
        pre = f"single_transformer_blocks.{i}"
        _linear(sd, f"{pre}.norm.linear", 3 * D, D, gen)
        _linear(sd, f"{pre}.proj_mlp", 4 * D, D, gen)
        _linear(sd, f"{pre}.proj_out", D, 5 * D, gen)
        for nm in ("to_q", "to_k", "to_v"):
            _linear(sd, f"{pre}.attn.{nm}", D, D, gen)
        for nm in ("norm_q", "norm_k"):
            sd[f"{pre}.attn.{nm}.weight"] = torch.ones(hd)
    _linear(sd, "norm_out.linear", 2 * D, D, gen)
    _linear(sd, "proj_out", cfg.in_channels, D, gen)
    return sd


def load_diffusers_state_dict(path) -> dict[str, torch.Tensor]:
    """State dict of a diffusers-format transformer directory (or its parent pipeline directory): single or sharded
    ``diffusion_pytorch_model*.safetensors`` (with ``.index.json``) or a ``.bin``.  Local files only - the reference's
    ``from_pretrained`` (pixart_transformer_2d_edited.py:104-117, flux_transformer_2d_edited.py:104-150) resolves hub
    names, which needs a network this environment does not have."""
    import json
    from pathlib import Path

    root = Path(path)
    for d in (root, root / "transformer"):
        if not d.is_dir():
            continue
        index = d / "diffusion_pytorch_model.safetensors.index.json"
        single = d / "diffusion_pytorch_model.safetensors"
        legacy = d / "diffusion_pytorch_model.bin"
        if index.exists() or single.exists():
            from safetensors.torch import load_file

            files = sorted(set(json.loads(index.read_text())["weight_map"].values())) if index.exists() else [single.name]
            sd: dict[str, torch.Tensor] = {}
            for f in files:
                sd.update(load_file(str(d / f), device="cpu"))
            return sd
        if legacy.exists():
            return torch.load(legacy, map_location="cpu")
    raise FileNotFoundError(f"no diffusers-format transformer weights under {root}")


def pixart_config_from_pretrained(path) -> PixArtConfig:
    """``config.json`` of a diffusers PixArtTransformer2DModel directory (or its parent pipeline directory) ->
    PixArtConfig.  Fields this implementation is specialised for must hold their PixArt values; anything else raises
    instead of silently building the wrong model (position table scale, micro-condition embedders)."""
    import json
    from pathlib import Path

    root = Path(path)
    for d in (root, root / "transformer"):
        f = d / "config.json"
        if f.is_file() and ((d / "diffusion_pytorch_model.safetensors").exists()
                            or (d / "diffusion_pytorch_model.safetensors.index.json").exists()
                            or (d / "diffusion_pytorch_model.bin").exists()):
            raw = json.loads(f.read_text())
            break
    else:
        raise FileNotFoundError(f"no transformer config.json next to the weights under {root}")
    required = {"norm_type": "ada_norm_single", "activation_fn": "gelu-approximate", "attention_bias": True,
                "norm_elementwise_affine": False, "dropout": 0.0}
    for key, want in required.items():
        if key in raw and raw[key] != want:
            raise ValueError(f"unsupported PixArt config: {key} = {raw[key]!r} (this path implements {want!r})")
    for key in ("num_embeds_ada_norm", "attention_type"):
        if key in raw and raw[key] not in (None, 1000, "default"):
            raise ValueError(f"unsupported PixArt config: {key} = {raw[key]!r}")
    fields = {"num_attention_heads", "attention_head_dim", "in_channels", "out_channels", "num_layers",
              "cross_attention_dim", "sample_size", "patch_size", "norm_eps", "caption_channels",
              "interpolation_scale", "use_additional_conditions"}
    cfg = PixArtConfig(**{k: v for k, v in raw.items() if k in fields})
    if cfg.cross_attention_dim != cfg.inner_dim:
        raise ValueError("cross_attention_dim must equal heads * head_dim on the PixArt path")
    return cfg
