"""The CUDA path against INDEPENDENT third-party implementations, without the repo's own oracle in between
(tests/test_third_party_anchors.py pins the oracles to the same code on the CPU): Black Forest Labs' FLUX model and
the ldm VAE decoder as vendored by `torchtitan.experiments.flux`, evaluated in fp32 on the host on the same weights
(re-keyed with diffusers' published conversion rules) and the same inputs.  Bars as everywhere: bf16 kernels vs fp32,
rel. max-abs <= 1e-2 on a model output, cosine >= 0.999."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from anchor_util import bfl_state_dict, ldm_decoder_state_dict

pytestmark = pytest.mark.gpu


def test_flux_cuda_forward_matches_the_bfl_model(cuda_device):
    tt = pytest.importorskip("torchtitan.experiments.flux.model.model")
    from torchtitan.experiments.flux.model.args import FluxModelArgs

    from ecad_b200.flux_pipeline import latent_image_ids
    from ecad_b200.flux_transformer import B200FluxTransformer2D
    from ecad_b200.schedule import FluxCacheSchedule
    from ecad_b200.transformer import SequentialDiTScheduler
    from ecad_b200.weights import FluxConfig, flux_random_init_state_dict

    H, hd, L, LS, ctx, pooled, axes = 4, 128, 2, 3, 256, 64, (16, 56, 56)
    cfg = FluxConfig(num_attention_heads=H, attention_head_dim=hd, num_layers=L, num_single_layers=LS,
                     joint_attention_dim=ctx, pooled_projection_dim=pooled, axes_dims_rope=axes, guidance_embeds=False)
    sd = flux_random_init_state_dict(cfg, seed=5)
    g = torch.Generator().manual_seed(11)
    for k in sd:
        if "norm_" in k and k.endswith(".weight") and sd[k].ndim == 1:
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g)
    D = H * hd
    ref_model = tt.FluxModel(FluxModelArgs(in_channels=64, out_channels=64, vec_in_dim=pooled, context_in_dim=ctx,
                                           hidden_size=D, num_heads=H, depth=L, depth_single_blocks=LS, axes_dim=axes))
    ref_model.load_state_dict(bfl_state_dict(sd, L, LS, D), strict=True)
    for m in ref_model.modules():
        if isinstance(m, torch.nn.RMSNorm):
            m.eps = 1e-6  # BFL's own RMSNorm (and diffusers') add 1e-6; torchtitan's nn.RMSNorm default differs
    ref_model.eval()

    B, T, hh, ww = 2, 64, 16, 12  # 192 image + 64 text tokens = 256 joint tokens
    N = hh * ww
    img = torch.randn(B, N, 64, generator=g)
    txt = torch.randn(B, T, ctx, generator=g) * 0.2
    y = torch.randn(B, pooled, generator=g) * 0.2
    t = torch.tensor([0.83, 0.27])
    img_ids, txt_ids = latent_image_ids(B, hh, ww), torch.zeros(B, T, 3)
    with torch.no_grad():
        ref = ref_model(img, img_ids, txt, txt_ids, t, y)

    sched = FluxCacheSchedule.from_numpy(np.ones((1, L + LS, 3), bool), 1, L, LS, "dense")
    model = B200FluxTransformer2D(sd, cfg, SequentialDiTScheduler(1), sched)
    got = model(img, txt, y, t, img_ids, txt_ids, None, return_dict=False)[0].float().cpu()
    assert got.shape == ref.shape == (B, N, 64)
    rel = float((got - ref).abs().max() / ref.abs().max())
    cos = float(F.cosine_similarity(got.flatten().double(), ref.flatten().double(), dim=0))
    assert rel <= 1e-2 and cos >= 0.9999, (rel, cos)


def test_vae_cuda_decode_matches_the_ldm_decoder(cuda_device):
    ae = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")
    from ecad_b200.vae import B200VaeDecoder, VaeConfig, random_init_vae_state_dict

    cfg = VaeConfig.flux()
    sd = random_init_vae_state_dict(cfg, seed=6)
    sd = {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 else v) for k, v in sd.items()}  # weights as the GPU holds them
    p = ae.AutoEncoderParams()
    assert (p.z_channels, p.ch, tuple(p.ch_mult), p.scale_factor, p.shift_factor) == \
        (cfg.latent_channels, cfg.block_out_channels[0], (1, 2, 4, 4), cfg.scaling_factor, cfg.shift_factor)
    dec = ae.Decoder(ch=p.ch, out_ch=p.out_ch, ch_mult=list(p.ch_mult), num_res_blocks=p.num_res_blocks,
                     in_channels=p.in_channels, resolution=p.resolution, z_channels=p.z_channels)
    dec.load_state_dict(ldm_decoder_state_dict(sd), strict=True)
    dec.eval()
    z = torch.randn(2, 16, 16, 12, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = dec(z / p.scale_factor + p.shift_factor)  # AutoEncoder.decode
    img = B200VaeDecoder(sd, cfg).decode(z.cuda()).cpu()
    assert img.shape == ref.shape == (2, 3, 128, 96)
    scale = float(ref.abs().max())
    err = (img - ref).abs()
    assert float(err.max()) < 4e-2 * scale and float(err.mean()) < 6e-3 * scale  # the bars of tests/test_gpu_vae.py
    assert float(F.cosine_similarity(img.flatten(), ref.flatten(), dim=0)) > 0.999
