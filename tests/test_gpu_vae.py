"""VAE decode: every new C-ABI entry point against torch fp32 on the same bf16-rounded
inputs, then the whole decoder against the CPU oracle (oracle/vae_oracle.py) on identical random-init weights."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bordered(x_nchw):
    """fp32 NCHW -> zero-bordered NHWC bf16 on the GPU."""
    b, c, h, w = x_nchw.shape
    out = torch.zeros(b, h + 2, w + 2, c, dtype=torch.bfloat16, device="cuda")
    out[:, 1:-1, 1:-1, :] = x_nchw.permute(0, 2, 3, 1).to(device="cuda", dtype=torch.bfloat16)
    return out


def _interior(x_bordered):
    return x_bordered[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2).float()


def _border_is_zero(t):
    return (float(t[:, 0].abs().max()) == 0 and float(t[:, -1].abs().max()) == 0 and
            float(t[:, :, 0].abs().max()) == 0 and float(t[:, :, -1].abs().max()) == 0)


def _pack_conv(w):
    cout, cin, k, _ = w.shape
    return w.permute(0, 2, 3, 1).reshape(cout, k * k * cin).to(device="cuda", dtype=torch.bfloat16).contiguous()


@pytest.mark.parametrize("b,h,w,cin,cout,taps,res", [
    (2, 16, 16, 64, 128, 9, False),      # small: one-CTA kernels
    (3, 32, 32, 512, 512, 9, True),      # mid-block shape, residual
    (2, 40, 24, 128, 256, 9, False),     # non-square, rows not a multiple of 128
    (2, 32, 32, 512, 256, 1, False),     # 1x1 shortcut
    (24, 64, 64, 128, 128, 9, True),     # enough rows for the 2-CTA kernels; C_out = 128: taps of a row share one A tile
    (9, 64, 64, 256, 128, 9, False),     # same path, two channel blocks per tap
    (24, 64, 64, 128, 256, 9, False),    # 2-CTA kernels, nine loads per block
])
def test_conv_nhwc(cuda_device, b, h, w, cin, cout, taps, res):
    from ecad_b200 import _lib
    g = torch.Generator().manual_seed(b * 1000 + h + cin + cout + taps)
    k = 3 if taps == 9 else 1
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    bias = torch.randn(cout, generator=g)
    r = torch.randn(b, cout, h, w, generator=g) if res else None
    xb, wb = _bordered(x), _pack_conv(wt)
    rb = _bordered(r) if res else None
    out = torch.full((b, h + 2, w + 2, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.conv_nhwc(xb, wb, bias.cuda(), out, h, w, taps, residual=rb)
    torch.cuda.synchronize()
    ref = F.conv2d(_interior(xb), wt.to(torch.bfloat16).float().cuda(), bias.cuda(), padding=k // 2)
    if res:
        ref = ref + _interior(rb)
    assert torch.isfinite(out.float()).all()
    assert _border_is_zero(out)
    err = float((_interior(out) - ref).abs().max() / ref.abs().max())
    assert err < 1e-2, err  # bf16 output rounding (2^-8 relative)


@pytest.mark.parametrize("b,h,w,cin,cout", [(2, 16, 12, 128, 128), (3, 32, 32, 512, 512), (20, 64, 64, 256, 256)])
def test_conv_up2x_nhwc(cuda_device, b, h, w, cin, cout):
    """Upsample2D (nearest 2x + 3x3 conv) as four 2x2-tap convolutions of the original image."""
    from ecad_b200 import _lib
    from ecad_b200.vae import pack_upsample_conv
    g = torch.Generator().manual_seed(b + h + cin)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(torch.bfloat16).float()
    bias = torch.randn(cout, generator=g)
    xb = _bordered(x)
    w4 = pack_upsample_conv(wt, cin, cout).to(device="cuda", dtype=torch.bfloat16).contiguous()
    out = torch.full((b, 2 * h + 2, 2 * w + 2, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.conv_up2x_nhwc(xb, w4, bias.cuda(), out, h, w)
    torch.cuda.synchronize()
    ref = F.conv2d(F.interpolate(_interior(xb), scale_factor=2.0, mode="nearest"), wt.cuda(), bias.cuda(), padding=1)
    assert torch.isfinite(out.float()).all() and _border_is_zero(out)
    # the pre-summed 2x2 kernels are rounded to bf16 once more than the 3x3 kernel: 2^-8 relative on sums of <= 4 terms
    err = float((_interior(out) - ref).abs().max() / ref.abs().max())
    assert err < 1.5e-2, err


def test_conv_nhwc_few_output_channels(cuda_device):
    """conv_out: 3 real channels out of 128 weight rows, written into a 32-column buffer."""
    from ecad_b200 import _lib
    g = torch.Generator().manual_seed(5)
    b, h, w, cin = 2, 24, 24, 128
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(3, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    bias = torch.randn(3, generator=g)
    wp = torch.zeros(128, cin, 3, 3)
    wp[:3] = wt
    bp = torch.zeros(128)
    bp[:3] = bias
    xb = _bordered(x)
    out = torch.full((b, h + 2, w + 2, 32), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.conv_nhwc(xb, _pack_conv(wp), bp.cuda(), out, h, w, 9, out_cols=3)
    image = torch.empty(b, 3, h, w, device="cuda")
    _lib.vae_finish(out, image, h, w)
    torch.cuda.synchronize()
    ref = F.conv2d(_interior(xb), wt.to(torch.bfloat16).float().cuda(), bias.cuda(), padding=1)
    assert float((image - ref).abs().max() / ref.abs().max()) < 1e-2
    image2 = torch.empty_like(image)
    _lib.vae_finish(out, image2, h, w, denormalize=True)
    assert torch.allclose(image2, (image / 2 + 0.5).clamp(0, 1), atol=1e-6)


@pytest.mark.parametrize("c,silu,unpadded", [(512, True, False), (256, True, False), (128, True, False), (512, False, True)])
def test_groupnorm_nhwc(cuda_device, c, silu, unpadded):
    from ecad_b200 import _lib
    g = torch.Generator().manual_seed(c)
    b, h, w = 3, 24, 40
    x = torch.randn(b, c, h, w, generator=g) * 2.0 + 0.5
    gamma = 1 + 0.1 * torch.randn(c, generator=g)
    beta = 0.1 * torch.randn(c, generator=g)
    xb = _bordered(x)
    scratch = torch.empty(_lib.groupnorm_scratch_bytes(b, h, w) + 64, dtype=torch.uint8, device="cuda")
    out = (torch.empty(b, h * w, c, dtype=torch.bfloat16, device="cuda") if unpadded
           else torch.full_like(xb, float("nan")))
    _lib.groupnorm_nhwc(xb, gamma.cuda(), beta.cuda(), out, scratch, h, w, silu=silu, unpadded_out=unpadded)
    torch.cuda.synchronize()
    ref = F.group_norm(_interior(xb), 32, gamma.cuda(), beta.cuda(), 1e-6)
    if silu:
        ref = F.silu(ref)
    if unpadded:
        got = out.view(b, h, w, c).permute(0, 3, 1, 2).float()
    else:
        assert _border_is_zero(out)
        got = _interior(out)
    assert float((got - ref).abs().max()) < 3e-2  # bf16 rounding of values up to ~4
    assert float((got - ref).abs().mean()) < 3e-3


def test_upsample_softmax_prepare_add(cuda_device):
    from ecad_b200 import _lib
    g = torch.Generator().manual_seed(3)
    b, h, w, c = 2, 12, 20, 128
    x = torch.randn(b, c, h, w, generator=g)
    xb = _bordered(x)
    up = torch.full((b, 2 * h + 2, 2 * w + 2, c), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.upsample2x_nhwc(xb, up, h, w)
    assert _border_is_zero(up)
    assert torch.equal(_interior(up), F.interpolate(_interior(xb), scale_factor=2.0, mode="nearest"))
    # softmax rows
    s = torch.randn(300, 1024, generator=g).cuda() * 30
    p = torch.empty(300, 1024, dtype=torch.bfloat16, device="cuda")
    _lib.softmax_rows(s, p, 1.0 / math.sqrt(512))
    ref = torch.softmax(s / math.sqrt(512), dim=-1)
    assert float((p.float() - ref).abs().max()) < 4e-3 * float(ref.max())
    _lib.softmax_rows(s, p, 1.0 / math.sqrt(512), valid_cols=1000)  # 24 padding keys: probability 0
    ref = torch.softmax(s[:, :1000] / math.sqrt(512), dim=-1)
    assert float(p[:, 1000:].abs().max()) == 0
    assert float((p[:, :1000].float() - ref).abs().max()) < 4e-3 * float(ref.max())
    # latent preparation
    z = torch.randn(b, 4, h, w, generator=g)
    pw = torch.randn(4, 4, generator=g)
    pb = torch.randn(4, generator=g)
    lat = torch.full((b, h + 2, w + 2, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.vae_prepare_latents(z.cuda(), pw.cuda(), pb.cuda(), 1 / 0.18215, lat)
    assert _border_is_zero(lat) and float(lat[..., 4:].abs().max()) == 0
    ref = F.conv2d(z / 0.18215, pw.view(4, 4, 1, 1), pb).cuda()
    assert float((_interior(lat)[:, :4] - ref).abs().max() / ref.abs().max()) < 1e-2
    # residual add of attention tokens
    tok = torch.randn(b, h * w, c, generator=g).to(device="cuda", dtype=torch.bfloat16)
    out = torch.full_like(xb, float("nan"))
    _lib.vae_add_tokens(xb, tok, out, h, w)
    assert _border_is_zero(out)
    ref = _interior(xb) + tok.view(b, h, w, c).permute(0, 3, 1, 2).float()
    assert float((_interior(out) - ref).abs().max()) < 3e-2


@pytest.mark.parametrize("hw,batch", [(16, 2), (32, 1)])
def test_vae_decode_matches_oracle(cuda_device, hw, batch):
    """Whole decoder, full width (512/512/256/128), random-init weights: 128x128 (batch 2) and 256x256 images.
    Tolerance: the decoder is ~30 bf16 layers deep; measured max error 1-2 % of the output range, asserted at 4 %
    (max) and 0.6 % (mean) of the oracle's max magnitude."""
    from ecad_b200.vae import B200VaeDecoder, VaeConfig, random_init_vae_state_dict
    from oracle.vae_oracle import OracleVaeConfig, vae_decode
    cfg = VaeConfig()
    sd = random_init_vae_state_dict(cfg, seed=0)
    # both sides see bf16-representable convolution / projection weights (the GPU path stores them in bf16)
    sd = {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 and not k.startswith("post_quant") else v) for k, v in sd.items()}
    lat = torch.randn(batch, 4, hw, hw, generator=torch.Generator().manual_seed(1)) * 0.9
    dec = B200VaeDecoder(sd, cfg)
    img = dec.decode(lat.cuda())
    torch.cuda.synchronize()
    ref = vae_decode(sd, lat, OracleVaeConfig())
    assert img.shape == (batch, 3, 8 * hw, 8 * hw) == tuple(ref.shape)
    assert torch.isfinite(img).all()
    scale = float(ref.abs().max())
    err = (img.cpu() - ref).abs()
    assert float(err.max()) < 4e-2 * scale, (float(err.max()), scale)
    assert float(err.mean()) < 6e-3 * scale, (float(err.mean()), scale)
    cos = float(F.cosine_similarity(img.cpu().flatten(), ref.flatten(), dim=0))
    assert cos > 0.999, cos
    # determinism, and the denormalised form
    img2 = dec.decode(lat.cuda(), denormalize=True)
    assert torch.allclose(img2, (img / 2 + 0.5).clamp(0, 1), atol=1e-6)
    # sliced decode (batch above the activation budget): the same computation sample by sample; kernels chosen by
    # problem size (1-CTA / 2-CTA tiles, shared-row taps) sum the taps in a different order, so not bit-identical
    if batch > 1:
        dec.max_activation_bytes = 1
        sliced = dec.decode(lat.cuda())
        assert sliced.shape == img.shape and float((sliced - img).abs().max()) < 2e-2 * scale
        dec.max_activation_bytes = type(dec).max_activation_bytes
    # the un-fused form of Upsample2D (upsample kernel + 3x3 convolution) agrees with the fused one
    dec.fused_upsample = False
    img3 = dec.decode(lat.cuda())
    assert float((img3 - img).abs().max()) < 4e-2 * scale and float((img3.cpu() - ref).abs().max()) < 4e-2 * scale


def test_generator_output_types(cuda_device, tmp_path):
    """`output_type` of the generators (pass_through.py:382-396): "latent" (default) / "pt" / "np" / "pil" from the same
    seed; the decoded forms equal an explicit decode of the latents; generate_from_saved_prompts writes .jpg files for
    PIL output like the reference (image_generator.py:423-440)."""
    import numpy as np
    from ecad_b200.dataset import PIXART_KEYS
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator, B200PixArtSigmaImageGenerator
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings

    L, steps = 2, 3
    cfg = PixArtConfig(num_layers=L)
    flags = np.random.default_rng(2).random((steps, L, 3)) < 0.6
    emb = synthetic_prompt_embeddings(2, seed=4)
    gen = B200PixArtAlphaImageGenerator(cache_schedule=PixArtCacheSchedule.from_numpy(flags, steps, L), state_dict=random_init_state_dict(cfg, 1),
                                        model_config=cfg)
    lat = gen.generate_images(emb)[0]
    assert lat.shape == (2, 4, 32, 32)
    pt = gen.generate_images(emb, output_type="pt")[0]
    assert pt.shape == (2, 3, 256, 256) and float(pt.min()) >= 0 and float(pt.max()) <= 1
    assert torch.equal(pt, gen.create_vae().decode(lat, denormalize=True))
    arr = gen.generate_images(emb, output_type="np")[0]
    assert arr.shape == (2, 256, 256, 3) and np.allclose(arr, pt.permute(0, 2, 3, 1).cpu().numpy())
    pil = gen.generate_images(emb, output_type="pil")[0]
    assert len(pil) == 2 and pil[0].size == (256, 256)
    assert gen.vae_config.scaling_factor == 0.18215
    assert B200PixArtSigmaImageGenerator.default_vae_config.scaling_factor == 0.13025
    # file-driven API with PIL output
    src, dst = tmp_path / "prompts", tmp_path / "images"
    src.mkdir()
    for i in range(2):
        torch.save({k: emb[k][i:i + 1] for k in PIXART_KEYS}, src / f"p{i}.pt")
    gen.output_type = "pil"
    gen.generate_from_saved_prompts(src, dst, batch_size=2)
    assert sorted(p.name for p in dst.glob("**/*.jpg")) == ["p0__image_seed:000.jpg", "p1__image_seed:000.jpg"]
    ms = gen.generate_images_timed(emb)  # timed INCLUDING the decode, like the reference's latency figures
    assert ms > 0


def test_flux_vae_variant(cuda_device):
    """The FLUX.1 VAE: 16 latent channels, `latents / scaling_factor + shift_factor`, no post_quant_conv - latent
    preparation alone, then the whole decoder against the oracle (128 x 96 image)."""
    from ecad_b200 import _lib
    from ecad_b200.vae import B200VaeDecoder, VaeConfig, decoder_layer_names, random_init_vae_state_dict
    from oracle.vae_oracle import OracleVaeConfig, vae_decode
    cfg = VaeConfig.flux()
    assert cfg.latent_channels == 16 and not cfg.use_post_quant_conv
    assert "post_quant_conv.weight" not in decoder_layer_names(cfg)
    assert decoder_layer_names(cfg)["decoder.conv_in.weight"] == (512, 16, 3, 3)
    g = torch.Generator().manual_seed(2)
    z = torch.randn(2, 16, 16, 12, generator=g)
    lat = torch.full((2, 18, 14, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.vae_prepare_latents(z.cuda(), None, None, 1 / cfg.scaling_factor, lat, shift=cfg.shift_factor)
    assert _border_is_zero(lat) and float(lat[..., 16:].abs().max()) == 0
    ref = (z / cfg.scaling_factor + cfg.shift_factor).cuda()
    assert float((_interior(lat)[:, :16] - ref).abs().max() / ref.abs().max()) < 1e-2
    sd = random_init_vae_state_dict(cfg, seed=1)
    sd = {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 else v) for k, v in sd.items()}
    img = B200VaeDecoder(sd, cfg).decode(z.cuda())
    torch.cuda.synchronize()
    oref = vae_decode(sd, z, OracleVaeConfig(latent_channels=16, scaling_factor=cfg.scaling_factor,
                                             shift_factor=cfg.shift_factor, use_post_quant_conv=False))
    assert img.shape == (2, 3, 128, 96) == tuple(oref.shape)
    scale = float(oref.abs().max())
    err = (img.cpu() - oref).abs()
    assert float(err.max()) < 4e-2 * scale and float(err.mean()) < 6e-3 * scale
    assert float(F.cosine_similarity(img.cpu().flatten(), oref.flatten(), dim=0)) > 0.999


def test_flux_generator_decodes_packed_latents(cuda_device):
    """B200FluxImageGenerator(output_type="pt"): packed latents [B, N, 64] -> unpack -> FLUX VAE decode, equal to an
    explicit unpack + decode of the latent output of the same seed."""
    from ecad_b200.image_generator import B200FluxImageGenerator
    from ecad_b200.schedule import FluxCacheSchedule
    from ecad_b200.weights import FluxConfig, flux_random_init_state_dict
    from test_gpu_flux_parity import SMALL, _embeds, _schedule_flags

    cfg = FluxConfig(**SMALL)
    steps, rows = 3, cfg.num_layers + cfg.num_single_layers
    sched = FluxCacheSchedule.from_numpy(_schedule_flags(steps, rows), steps, cfg.num_layers, cfg.num_single_layers,
                                         "rand", top_level_config={"height": 256, "width": 192})
    gen = B200FluxImageGenerator(cache_schedule=sched, state_dict=flux_random_init_state_dict(cfg, seed=0), model_config=cfg)
    emb = _embeds(2, 64, SMALL, seed=4)
    lat = gen.generate_images(emb)[0]
    assert lat.shape == (2, 16 * 12, 64)
    img = gen.generate_images(emb, output_type="pt")[0]
    assert img.shape == (2, 3, 256, 192) and float(img.min()) >= 0 and float(img.max()) <= 1
    z = lat.view(2, 16, 12, 16, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(2, 16, 32, 24)  # FluxPipeline._unpack_latents
    assert torch.equal(img, gen.create_vae().decode(z, denormalize=True))
    assert gen.vae_config.scaling_factor == 0.3611 and gen.vae_config.shift_factor == 0.1159
