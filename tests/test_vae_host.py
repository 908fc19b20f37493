"""CPU checks of the VAE oracle and of the host-side weight packing (no GPU)."""
import math

import torch
import torch.nn.functional as F

from ecad_b200.vae import B200VaeDecoder, VaeConfig, decoder_layer_names, random_init_vae_state_dict
from oracle.vae_oracle import OracleVaeConfig, vae_decode


def test_decoder_parameter_inventory():
    names = decoder_layer_names()
    # SD VAE decoder: 1 + 1 convs, 2 + 12 resnets (2 with a shortcut), 3 upsampler convs, 1 attention, out norm + conv
    assert names["decoder.conv_in.weight"] == (512, 4, 3, 3)
    assert names["decoder.conv_out.weight"] == (3, 128, 3, 3)
    assert names["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"] == (256, 512, 1, 1)
    assert names["decoder.up_blocks.3.resnets.0.conv_shortcut.weight"] == (128, 256, 1, 1)
    assert "decoder.up_blocks.0.resnets.0.conv_shortcut.weight" not in names
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in names
    assert sum(1 for n in names if n.endswith(".weight")) == 2 + 14 * 4 + 2 + 3 + 5 + 1 + 1
    n_params = sum(math.prod(s) for s in names.values())
    assert 49_000_000 < n_params < 50_000_000  # the SD decoder has 49.5 M parameters


def test_oracle_shapes_and_determinism():
    cfg = VaeConfig(block_out_channels=(32, 32, 64, 64))  # narrow: seconds on CPU
    sd = random_init_vae_state_dict(cfg, seed=3)
    assert set(sd) == set(decoder_layer_names(cfg))
    lat = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    ocfg = OracleVaeConfig(block_out_channels=cfg.block_out_channels)
    a = vae_decode(sd, lat, ocfg)
    b = vae_decode(sd, lat, ocfg)
    assert a.shape == (2, 3, 64, 64) and torch.equal(a, b) and torch.isfinite(a).all()
    # samples are independent
    c = vae_decode(sd, lat[1:], ocfg)
    assert torch.allclose(a[1:], c, atol=1e-5)
    d = vae_decode(sd, lat, ocfg, denormalize=True)
    assert float(d.min()) >= 0 and float(d.max()) <= 1


def test_conv_weight_packing_is_the_tap_major_gemm_operand():
    """The packed [cout, (ky*3+kx)*cin + c] matrix times the shifted-row im2col of a bordered NHWC image equals
    F.conv2d - the identity the implicit-GEMM kernel relies on (checked here with plain torch on the CPU)."""
    g = torch.Generator().manual_seed(0)
    b, cin, cout, h, w = 2, 8, 5, 6, 7
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g)
    packed = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    xb = torch.zeros(b, h + 2, w + 2, cin)
    xb[:, 1:-1, 1:-1] = x.permute(0, 2, 3, 1)
    rows = xb.reshape(-1, cin)
    pitch = w + 2
    m = rows.shape[0]
    acc = torch.zeros(m, cout)
    for tap in range(9):
        ky, kx = divmod(tap, 3)
        off = (ky - 1) * pitch + (kx - 1)
        shifted = torch.zeros_like(rows)  # rows outside the matrix read as zero (TMA out-of-bounds fill)
        lo, hi = max(0, -off), min(m, m - off)
        shifted[lo:hi] = rows[lo + off:hi + off]
        acc += shifted @ packed[:, tap * cin:(tap + 1) * cin].T
    got = acc.view(b, h + 2, w + 2, cout)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
    assert torch.allclose(got, F.conv2d(x, wt, padding=1), atol=1e-4)


def test_decoder_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        return
    try:
        B200VaeDecoder(random_init_vae_state_dict(VaeConfig(block_out_channels=(32, 32, 64, 64))))
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("expected RuntimeError")


def test_flops_model_matches_published_order_of_magnitude():
    # SD VAE decode of a 64 x 64 latent (512 x 512 image): ~1.24 TMACs = 2.5 TFLOP (widely quoted figure)
    f = B200VaeDecoder.flops(1, 64, 64)
    assert 2.3e12 < f < 2.7e12


def test_upsample_conv_equals_four_2x2_convs_on_the_original_image():
    """pack_upsample_conv: conv3x3(nearest_2x(x)) == interleave of four 2x2-tap convolutions of x (the identity
    ecadk_conv_up2x_nhwc relies on), checked with plain torch on the CPU, including the image border."""
    from ecad_b200.vae import pack_upsample_conv
    g = torch.Generator().manual_seed(0)
    b, cin, cout, h, w = 2, 5, 7, 6, 9
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), wt, padding=1)
    w4 = pack_upsample_conv(wt, cin, cout).view(4, cout, 4, cin)
    xp = F.pad(x, (1, 1, 1, 1))
    got = torch.zeros(b, cout, 2 * h, 2 * w)
    for a in range(2):
        for bb in range(2):
            acc = torch.zeros(b, cout, h, w)
            for ry in range(2):
                for rx in range(2):
                    oy, ox = ry + a - 1, rx + bb - 1  # source offset of this tap
                    src = xp[:, :, 1 + oy:1 + oy + h, 1 + ox:1 + ox + w]
                    acc += torch.einsum("bchw,oc->bohw", src, w4[a * 2 + bb, :, ry * 2 + rx, :])
            got[:, :, a::2, bb::2] = acc
    assert torch.allclose(got, ref, atol=1e-4)
