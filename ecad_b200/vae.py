"""VAE decode on the B200: the step after the denoising loop.

Replaces ``image = self.vae.decode(latents / self.vae.config.scaling_factor, return_dict=False)[0]`` and
``image_processor.postprocess`` (/root/reference/ecad/pipelines/pass_through.py:382-396); the module behind
``self.vae`` is diffusers' ``AutoencoderKL`` (SD VAE for PixArt-alpha, SDXL VAE for PixArt-sigma, the 16-channel VAE
for FLUX.1: same decoder architecture, different ``scaling_factor`` / ``shift_factor`` / latent width).  ``B200VaeDecoder`` takes a state dict in diffusers' naming
(``post_quant_conv.*``, ``decoder.*``), keeps the weights resident as bf16 GEMM operands and runs every layer in
libecad_b200.so - there is no torch fallback:

* activations: zero-bordered NHWC bf16 ``[B, H+2, W+2, C]`` (the border is the convolutions' padding);
* 3x3 / 1x1 convolutions: implicit GEMMs on the tcgen05 kernels (``ecadk_conv_nhwc``: nine shifted TMA loads of the
  bordered image accumulate into one TMEM tile; bias, ResNet residual and border mask in the epilogue);
* GroupNorm + SiLU, nearest 2x upsampling, latent preparation (scaling + post_quant_conv), read-out: glue kernels;
* mid-block attention (one head of width 512 over H*W tokens): projections and both attention products on the GEMM
  kernels (``P V`` through ``V^T = W_v H^T``, the value bias added after - softmax rows sum to one), softmax as a
  row kernel over fp32 scores.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import _lib


@dataclass(frozen=True)
class VaeConfig:
    """AutoencoderKL config fields the decoder reads (defaults: stabilityai/sd-vae-ft-ema, the PixArt-alpha VAE)."""

    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215
    shift_factor: float = 0.0          # FLUX VAE: latents / scaling_factor + shift_factor
    use_post_quant_conv: bool = True   # FLUX VAE: False

    @property
    def up_widths(self) -> tuple:
        return tuple(reversed(self.block_out_channels))

    @staticmethod
    def flux() -> "VaeConfig":
        """black-forest-labs/FLUX.1-dev `vae/config.json`: 16 latent channels, no (post_)quant_conv."""
        return VaeConfig(latent_channels=16, scaling_factor=0.3611, shift_factor=0.1159, use_post_quant_conv=False)


def decoder_layer_names(cfg: VaeConfig = VaeConfig()) -> dict[str, tuple]:
    """name -> shape of every parameter ``B200VaeDecoder`` reads (diffusers AutoencoderKL naming)."""
    out: dict[str, tuple] = {}

    def conv(name, cout, cin, k):
        out[name + ".weight"] = (cout, cin, k, k)
        out[name + ".bias"] = (cout,)

    def norm(name, c):
        out[name + ".weight"] = (c,)
        out[name + ".bias"] = (c,)

    def linear(name, cout, cin):
        out[name + ".weight"] = (cout, cin)
        out[name + ".bias"] = (cout,)

    def resnet(pre, cin, cout):
        norm(pre + ".norm1", cin)
        conv(pre + ".conv1", cout, cin, 3)
        norm(pre + ".norm2", cout)
        conv(pre + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(pre + ".conv_shortcut", cout, cin, 1)

    widths = cfg.up_widths
    top = widths[0]
    if cfg.use_post_quant_conv:
        conv("post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    conv("decoder.conv_in", top, cfg.latent_channels, 3)
    resnet("decoder.mid_block.resnets.0", top, top)
    norm("decoder.mid_block.attentions.0.group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        linear(f"decoder.mid_block.attentions.0.{n}", top, top)
    resnet("decoder.mid_block.resnets.1", top, top)
    prev = top
    for i, wdt in enumerate(widths):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else wdt, wdt)
        if i < len(widths) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", wdt, wdt, 3)
        prev = wdt
    norm("decoder.conv_norm_out", widths[-1])
    conv("decoder.conv_out", cfg.out_channels, widths[-1], 3)
    return out


def random_init_vae_state_dict(cfg: VaeConfig = VaeConfig(), seed: int = 0) -> dict[str, torch.Tensor]:
    """fp32 CPU state dict of the decoder, deterministic in ``seed`` (no pretrained weights offline): torch's default
    Conv2d / Linear init (uniform +-1/sqrt(fan_in)); norm scales 1 + 0.1 N(0,1), norm shifts 0.1 N(0,1) so that the
    affine part of GroupNorm is exercised."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: dict[str, torch.Tensor] = {}
    for name, shape in decoder_layer_names(cfg).items():
        if len(shape) == 1 and (".norm" in name or "group_norm" in name or "conv_norm_out" in name):
            base = 1.0 if name.endswith(".weight") else 0.0
            sd[name] = base + 0.1 * torch.randn(shape, generator=gen)
        else:
            wname = name[: name.rfind(".")] + ".weight"
            wshape = decoder_layer_names(cfg)[wname]
            fan_in = math.prod(wshape[1:])
            sd[name] = (torch.rand(shape, generator=gen) * 2 - 1) / math.sqrt(fan_in)
    return sd


def _pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def pack_upsample_conv(w: torch.Tensor, cin_p: int, cout_p: int) -> torch.Tensor:
    """3x3 kernel [cout, cin, 3, 3] of the convolution that follows a nearest 2x upsampling -> the four 2x2 kernels
    ``[4 (a*2+b), cout_p, 4 (ry*2+rx) * cin_p + c]`` that act on the ORIGINAL image (ecadk_conv_up2x_nhwc): output row
    ``2y + a`` reads upsampled rows ``2y + a + ky - 1``, i.e. source rows ``y - 1, y, y`` (a = 0) or ``y, y, y + 1``
    (a = 1); kernel rows that land on the same source row are summed (in fp32, before the bf16 rounding)."""
    cout, cin = w.shape[:2]
    groups = {0: ([0], [1, 2]), 1: ([0, 1], [2])}  # parity -> kernel indices of source offsets (first, second)
    out = torch.zeros(4, cout_p, 4, cin_p)
    wf = w.detach().float()
    for a in range(2):
        for b in range(2):
            for ry in range(2):
                for rx in range(2):
                    acc = torch.zeros(cout, cin)
                    for ky in groups[a][ry]:
                        for kx in groups[b][rx]:
                            acc += wf[:, :, ky, kx]
                    out[a * 2 + b, :cout, ry * 2 + rx, :cin] = acc
    return out.reshape(4, cout_p, 4 * cin_p)


class B200VaeDecoder:
    """``decode(latents) -> image``: AutoencoderKL.decode(latents / scaling_factor) on the GPU through the C ABI."""

    def __init__(self, state_dict: dict[str, torch.Tensor], cfg: VaeConfig = VaeConfig(), device: str | int = "cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("B200VaeDecoder needs a CUDA device (there is no CPU fallback)")
        _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        missing = [k for k in decoder_layer_names(cfg) if k not in state_dict]
        if missing:
            raise KeyError(f"VAE state dict lacks {len(missing)} decoder tensors, e.g. {missing[:3]}")
        if any(w % cfg.norm_num_groups for w in cfg.block_out_channels):
            raise ValueError("block_out_channels must be multiples of norm_num_groups")
        self._w: dict[str, torch.Tensor] = {}
        sd = state_dict
        dev = self.device

        def f32(t):
            return t.detach().to(device=dev, dtype=torch.float32).contiguous()

        def conv(name):
            w = sd[name + ".weight"].detach().float()
            cout, cin, k, _ = w.shape
            cin_p, cout_p = _pad_to(cin, 64), _pad_to(cout, 128)
            # [cout, cin, ky, kx] -> [cout_p, (ky*3 + kx) * cin_p + c], zero rows / channels as padding
            wp = torch.zeros(cout_p, k * k, cin_p)
            wp[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, k * k, cin)
            self._w[name + ".weight"] = wp.reshape(cout_p, k * k * cin_p).to(device=dev, dtype=torch.bfloat16).contiguous()
            b = torch.zeros(cout_p)
            b[:cout] = sd[name + ".bias"].detach().float()
            self._w[name + ".bias"] = b.to(dev)

        # Upsample2D convolutions run on the original image as four 2x2 kernels (16 instead of 36 tap-products per
        # input pixel, no upsampled tensor); fused_upsample = False keeps upsample kernel + 3x3 convolution (A/B, tests)
        self.fused_upsample = True
        for name, shape in decoder_layer_names(cfg).items():
            if not name.endswith(".weight"):
                continue
            base = name[: -len(".weight")]
            if ".upsamplers." in name:
                cout, cin = shape[:2]
                self._w[base + ".weight4"] = pack_upsample_conv(sd[name], _pad_to(cin, 64), _pad_to(cout, 128)).to(
                    device=dev, dtype=torch.bfloat16).contiguous()
            if base == "post_quant_conv":
                self._w[base + ".weight"] = f32(sd[name].reshape(cfg.latent_channels, cfg.latent_channels))
                self._w[base + ".bias"] = f32(sd[base + ".bias"])
            elif len(shape) == 4:
                conv(base)
            elif len(shape) == 2:  # attention projections
                self._w[name] = sd[name].detach().to(device=dev, dtype=torch.bfloat16).contiguous()
                self._w[base + ".bias"] = f32(sd[base + ".bias"])
            else:
                self._w[name] = f32(sd[name])
                self._w[base + ".bias"] = f32(sd[base + ".bias"])
        if cfg.latent_channels not in (4, 16) or cfg.out_channels != 3:
            raise ValueError("the latent preparation / read-out kernels are written for 4 or 16 latent and 3 image channels")
        self.launches = 0

    # ------------------------------------------------------------------------------------------------
    def _new(self, b, h, w, c):
        return torch.empty(b, h + 2, w + 2, c, device=self.device, dtype=torch.bfloat16)

    def _gn(self, x, name, h, w, silu=True):
        out = torch.empty_like(x)
        _lib.groupnorm_nhwc(x, self._w[name + ".weight"], self._w[name + ".bias"], out, self._scratch, h, w,
                            groups=self.cfg.norm_num_groups, eps=self.cfg.norm_eps, silu=silu)
        self.launches += 3
        return out

    def _conv(self, x, name, h, w, taps=9, residual=None):
        wgt = self._w[name + ".weight"]
        out = self._new(x.shape[0], h, w, wgt.shape[0])
        _lib.conv_nhwc(x, wgt, self._w[name + ".bias"], out, h, w, taps, residual=residual)
        self.launches += 1
        return out

    def _resnet(self, x, pre, h, w):
        a = self._conv(self._gn(x, pre + ".norm1", h, w), pre + ".conv1", h, w)
        a = self._gn(a, pre + ".norm2", h, w)
        sc = self._conv(x, pre + ".conv_shortcut", h, w, taps=1) if (pre + ".conv_shortcut.weight") in self._w else x
        return self._conv(a, pre + ".conv2", h, w, residual=sc)

    def _attention(self, x, pre, h, w):
        b, c = x.shape[0], x.shape[-1]
        n = h * w
        if n % 4:
            raise ValueError(f"mid-block attention needs h*w % 4 == 0, got {h}x{w}")
        npad = _pad_to(n, 128)  # token rows per sample: zero rows pad the keys to the GEMM tile, masked in the softmax
        bf = torch.bfloat16
        t = torch.zeros(b, npad, c, device=self.device, dtype=bf) if npad != n else \
            torch.empty(b, n, c, device=self.device, dtype=bf)
        _lib.groupnorm_nhwc(x, self._w[pre + ".group_norm.weight"], self._w[pre + ".group_norm.bias"], t, self._scratch, h,
                            w, groups=self.cfg.norm_num_groups, eps=self.cfg.norm_eps, silu=False, unpadded_out=True)
        t2 = t.view(b * npad, c)
        q = torch.empty(b * npad, c, device=self.device, dtype=bf)
        k = torch.empty_like(q)
        _lib.gemm_bias(t2, self._w[pre + ".to_q.weight"], self._w[pre + ".to_q.bias"], q)
        _lib.gemm_bias(t2, self._w[pre + ".to_k.weight"], self._w[pre + ".to_k.bias"], k)
        scores = torch.empty(npad, npad, device=self.device, dtype=torch.float32)
        probs = torch.empty(npad, npad, device=self.device, dtype=bf)
        vt = torch.empty(c, npad, device=self.device, dtype=bf)
        o = torch.empty(b * npad, c, device=self.device, dtype=bf)
        lib = _lib.load()
        stream = _lib.stream_ptr()
        for s in range(b):
            rows = slice(s * npad, (s + 1) * npad)
            # scores = q k^T (fp32 out), probabilities in bf16 (padding keys get 0)
            _lib.check(lib.ecadk_gemm_bias_f32(q[rows].data_ptr(), k[rows].data_ptr(), None, scores.data_ptr(), npad, npad,
                                               c, npad, npad, stream), "vae attention scores")
            _lib.softmax_rows(scores, probs, 1.0 / math.sqrt(c), valid_cols=n)
            # V^T = W_v H^T without the bias; out = P V + b_v (rows of P sum to one)
            _lib.gemm_bias(self._w[pre + ".to_v.weight"], t2[rows], None, vt)
            _lib.gemm_bias(probs, vt, self._w[pre + ".to_v.bias"], o[rows])
        out_tok = torch.empty(b, npad, c, device=self.device, dtype=bf)
        _lib.gemm_bias(o, self._w[pre + ".to_out.0.weight"], self._w[pre + ".to_out.0.bias"], out_tok.view(b * npad, c))
        out = torch.empty_like(x)
        _lib.vae_add_tokens(x, out_tok, out, h, w)
        self.launches += 7 + 4 * b
        return out

    # bytes of the largest activation a single decode pass may hold (bf16 [b, 8h+2, 8w+2, 2*C_last]); batches above it
    # are decoded in slices (diffusers' `enable_slicing`, sized to the 180 GB of a B200: a pass keeps ~5 such tensors)
    max_activation_bytes = 8 << 30

    @torch.no_grad()
    def decode(self, latents: torch.Tensor, denormalize: bool = False) -> torch.Tensor:
        """latents fp32 ``[B, 4, h, w]`` as the denoising loop leaves them (NOT yet divided by ``scaling_factor``) ->
        image fp32 ``[B, 3, 8h, 8w]``; ``denormalize`` applies ``(x / 2 + 0.5).clamp(0, 1)``."""
        if latents.dim() != 4 or latents.shape[1] != self.cfg.latent_channels:
            raise ValueError(f"latents must be [B, {self.cfg.latent_channels}, h, w], got {tuple(latents.shape)}")
        up = 2 ** (len(self.cfg.block_out_channels) - 1)
        per_sample = (latents.shape[2] * up + 2) * (latents.shape[3] * up + 2) * 2 * self.cfg.up_widths[-1] * 2
        chunk = max(1, int(self.max_activation_bytes // per_sample))
        if latents.shape[0] > chunk:
            return torch.cat([self._decode(latents[i:i + chunk], denormalize) for i in range(0, latents.shape[0], chunk)])
        with _lib.nvtx_range(f"vae decode B={latents.shape[0]}"):
            return self._decode(latents, denormalize)

    def _decode(self, latents: torch.Tensor, denormalize: bool) -> torch.Tensor:
        cfg = self.cfg
        z = latents.to(device=self.device, dtype=torch.float32).contiguous()
        b, _, h, w = z.shape
        # GroupNorm partial sums at the largest resolution of the decode
        up = 2 ** (len(cfg.block_out_channels) - 1)
        self._scratch = torch.empty(_lib.groupnorm_scratch_bytes(b, h * up, w * up, cfg.norm_num_groups) + 64,
                                    device=self.device, dtype=torch.uint8)
        x = _lib.vae_prepare_latents(z, self._w.get("post_quant_conv.weight"), self._w.get("post_quant_conv.bias"),
                                     1.0 / cfg.scaling_factor, self._new(b, h, w, 64), shift=cfg.shift_factor)
        self.launches += 1
        x = self._conv(x, "decoder.conv_in", h, w)
        x = self._resnet(x, "decoder.mid_block.resnets.0", h, w)
        x = self._attention(x, "decoder.mid_block.attentions.0", h, w)
        x = self._resnet(x, "decoder.mid_block.resnets.1", h, w)
        n_up = len(cfg.block_out_channels)
        for i in range(n_up):
            for j in range(cfg.layers_per_block + 1):
                with _lib.nvtx_range(f"up_blocks.{i}.resnets.{j} {h}x{w}"):
                    x = self._resnet(x, f"decoder.up_blocks.{i}.resnets.{j}", h, w)
            if i < n_up - 1:
                name = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                if self.fused_upsample:
                    w4 = self._w[name + ".weight4"]
                    x = _lib.conv_up2x_nhwc(x, w4, self._w[name + ".bias"], self._new(b, 2 * h, 2 * w, w4.shape[1]), h, w)
                    h, w = 2 * h, 2 * w
                    self.launches += 5
                else:
                    up = _lib.upsample2x_nhwc(x, self._new(b, 2 * h, 2 * w, x.shape[-1]), h, w)
                    h, w = 2 * h, 2 * w
                    self.launches += 1
                    x = self._conv(up, name, h, w)
        x = self._gn(x, "decoder.conv_norm_out", h, w)
        y = torch.empty(b, h + 2, w + 2, 32, device=self.device, dtype=torch.bfloat16)
        _lib.conv_nhwc(x, self._w["decoder.conv_out.weight"], self._w["decoder.conv_out.bias"], y, h, w, 9,
                       out_cols=cfg.out_channels)
        image = torch.empty(b, cfg.out_channels, h, w, device=self.device, dtype=torch.float32)
        _lib.vae_finish(y, image, h, w, denormalize=denormalize)
        self.launches += 2
        return image

    @staticmethod
    def flops(batch: int, h: int, w: int, cfg: VaeConfig = VaeConfig()) -> float:
        """Algorithmic FLOPs (2 x MACs of the real channels; convolutions, attention projections and products) of one
        decode of ``batch`` latents of ``h x w``."""
        widths = cfg.up_widths
        top = widths[0]
        macs = 0.0

        def conv(cin, cout, hh, ww, k=3):
            return float(hh * ww) * cin * cout * k * k

        def resnet(cin, cout, hh, ww):
            m = conv(cin, cout, hh, ww) + conv(cout, cout, hh, ww)
            return m + (conv(cin, cout, hh, ww, 1) if cin != cout else 0.0)

        macs += (h * w * cfg.latent_channels**2 if cfg.use_post_quant_conv else 0.0) + conv(cfg.latent_channels, top, h, w)
        macs += 2 * resnet(top, top, h, w)
        n = h * w
        macs += 4.0 * n * top * top + 2.0 * n * n * top
        prev, hh, ww = top, h, w
        for i, wdt in enumerate(widths):
            for j in range(cfg.layers_per_block + 1):
                macs += resnet(prev if j == 0 else wdt, wdt, hh, ww)
            if i < len(widths) - 1:
                hh, ww = 2 * hh, 2 * ww
                macs += conv(wdt, wdt, hh, ww)  # the reference's count (3x3 on the upsampled image)
            prev = wdt
        macs += conv(widths[-1], cfg.out_channels, hh, ww)
        return 2.0 * macs * batch
