"""CPU fp32 restatement of the VAE decode step of the reference's pipelines - TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU legs may import this module; the product path
(ecad_b200/vae.py -> libecad_b200.so) never does.

What it restates: ``image = self.vae.decode(latents / self.vae.config.scaling_factor, return_dict=False)[0]``
(/root/reference/ecad/pipelines/pass_through.py:382-385) followed by ``image_processor.postprocess`` (:393-396).  The
module behind ``self.vae`` is diffusers 0.30.3 ``AutoencoderKL`` (requirements.txt:8) - NOT vendored in the reference
and not installable here - so the algorithm below is restated from its published definition (SD / SDXL VAE decoder):

  z -> post_quant_conv (1x1) -> Decoder:
       conv_in 3x3 (4 -> 512)
       UNetMidBlock2D: ResnetBlock2D(512) -> Attention(heads 1, dim_head 512, GroupNorm(32), residual) -> ResnetBlock2D(512)
       4 x UpDecoderBlock2D over reversed block_out_channels (512, 512, 256, 128): 3 ResnetBlock2D each
            (the first maps prev -> out, with a 1x1 conv_shortcut when the widths differ), then for all but the last
            block Upsample2D = nearest 2x + conv 3x3
       conv_norm_out GroupNorm(32, eps 1e-6) -> SiLU -> conv_out 3x3 (128 -> 3)
  ResnetBlock2D (no time embedding in the VAE): GroupNorm(32, eps 1e-6) -> SiLU -> conv1 3x3 -> GroupNorm -> SiLU ->
       conv2 3x3; output = (shortcut(x) + h) / output_scale_factor (= 1)
  Attention: h = GroupNorm(x) as tokens [B, HW, C]; q, k, v = Linear(h); softmax(q k^T / sqrt(C)) v; Linear; + x

PARITY: the reference ships no VAE tensors, images or hashes and diffusers cannot be imported here; the decoder is
pinned instead to the ldm decoder AutoencoderKL is converted from - `torchtitan.experiments.flux.model.autoencoder.
Decoder` is installed in this image - on identical weights re-keyed with diffusers' published conversion rules
(tests/test_third_party_anchors.py: rel. max error < 2e-5, FLUX variant: 16 latent channels, shift factor, no
post_quant_conv).  `post_quant_conv` (a 1x1 convolution) and `postprocess` stay recalled.  The CUDA path is compared
against this file on identical random-init weights.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class OracleVaeConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215
    shift_factor: float = 0.0          # FLUX: latents / scaling_factor + shift_factor (FluxPipeline.__call__)
    use_post_quant_conv: bool = True   # FLUX VAE config: use_post_quant_conv = False


def _gn(x, sd, name, cfg, silu):
    y = F.group_norm(x, cfg.norm_num_groups, sd[name + ".weight"], sd[name + ".bias"], cfg.norm_eps)
    return F.silu(y) if silu else y


def _conv(x, sd, name, padding):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], padding=padding)


def _resnet(x, sd, pre, cfg):
    h = _conv(_gn(x, sd, pre + ".norm1", cfg, True), sd, pre + ".conv1", 1)
    h = _conv(_gn(h, sd, pre + ".norm2", cfg, True), sd, pre + ".conv2", 1)
    if pre + ".conv_shortcut.weight" in sd:
        x = _conv(x, sd, pre + ".conv_shortcut", 0)
    return x + h


def _attention(x, sd, pre, cfg):
    b, c, hh, ww = x.shape
    t = _gn(x.reshape(b, c, hh * ww), sd, pre + ".group_norm", cfg, False).transpose(1, 2)  # [B, HW, C]
    q = F.linear(t, sd[pre + ".to_q.weight"], sd[pre + ".to_q.bias"])
    k = F.linear(t, sd[pre + ".to_k.weight"], sd[pre + ".to_k.bias"])
    v = F.linear(t, sd[pre + ".to_v.weight"], sd[pre + ".to_v.bias"])
    p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1)
    o = F.linear(p @ v, sd[pre + ".to_out.0.weight"], sd[pre + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(b, c, hh, ww)


@torch.no_grad()
def vae_decode(sd: dict, latents: torch.Tensor, cfg: OracleVaeConfig = OracleVaeConfig(), denormalize: bool = False):
    """latents fp32 [B, 4, h, w] (as the denoising loop leaves them) -> image fp32 [B, 3, 8h, 8w]."""
    sd = {k: v.float() for k, v in sd.items()}
    x = latents.float() / cfg.scaling_factor + cfg.shift_factor
    if cfg.use_post_quant_conv:
        x = _conv(x, sd, "post_quant_conv", 0)
    x = _conv(x, sd, "decoder.conv_in", 1)
    x = _resnet(x, sd, "decoder.mid_block.resnets.0", cfg)
    x = _attention(x, sd, "decoder.mid_block.attentions.0", cfg)
    x = _resnet(x, sd, "decoder.mid_block.resnets.1", cfg)
    n_up = len(cfg.block_out_channels)
    for i in range(n_up):
        for j in range(cfg.layers_per_block + 1):
            x = _resnet(x, sd, f"decoder.up_blocks.{i}.resnets.{j}", cfg)
        if i < n_up - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(x, sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", 1)
    x = _gn(x, sd, "decoder.conv_norm_out", cfg, True)
    x = _conv(x, sd, "decoder.conv_out", 1)
    if denormalize:
        x = (x / 2 + 0.5).clamp(0, 1)
    return x
