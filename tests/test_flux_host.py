"""Host-side FLUX logic (no GPU): flow-match sigmas / Euler coefficients, latent packing, position ids and RoPE tables
against the oracle's restatement of diffusers' EmbedND; ctypes struct layout of the FLUX executor arguments."""
import ctypes as C

import numpy as np
import torch

from ecad_b200 import _lib
from ecad_b200.flux_pipeline import FlowMatchEulerDiscrete, calculate_shift, latent_image_ids, pack_latents
from ecad_b200.flux_transformer import rope_tables
from oracle.flux_oracle import embed_nd, flux_sigmas


def test_flow_match_sigmas_match_oracle():
    for steps, n in ((20, 256), (20, 4096), (7, 192)):
        s = FlowMatchEulerDiscrete()
        c = s.config
        s.set_timesteps(steps, mu=calculate_shift(n, c.base_image_seq_len, c.max_image_seq_len, c.base_shift, c.max_shift))
        ref = flux_sigmas(steps, n)
        assert np.array_equal(s.sigmas, ref)
        assert np.allclose(s.timesteps.numpy(), ref[:-1] * 1000)
        for i in range(steps):
            assert s.step_coefficient() == float(ref[i + 1]) - float(ref[i])
            s.advance()
    assert s.sigmas[-1] == 0.0


def test_flow_match_shift_known_answers():
    """FLUX.1-dev scheduler_config: base_shift 0.5 at 256 image tokens, max_shift 1.15 at 4096 (NOT the 1.16 default of
    calculate_shift's signature, which the pipeline never uses); linear in between."""
    import math

    c = FlowMatchEulerDiscrete().config
    assert (c.base_shift, c.max_shift, c.base_image_seq_len, c.max_image_seq_len) == (0.5, 1.15, 256, 4096)

    def mu(n):
        return calculate_shift(n, c.base_image_seq_len, c.max_image_seq_len, c.base_shift, c.max_shift)

    assert abs(mu(256) - 0.5) < 1e-12 and abs(mu(4096) - 1.15) < 1e-12
    assert abs(mu(1024) - (0.5 + 768 * 0.65 / 3840)) < 1e-12
    s = FlowMatchEulerDiscrete()
    s.set_timesteps(20, mu=mu(4096))
    # sigma' = e^mu / (e^mu + 1/sigma - 1): first sigma stays 1, the last one is e^1.15 / (e^1.15 + 19)
    assert s.sigmas[0] == 1.0
    assert abs(float(s.sigmas[19]) - math.exp(1.15) / (math.exp(1.15) + 19.0)) < 1e-7
    assert abs(float(s.sigmas[9]) - math.exp(1.15) / (math.exp(1.15) + (1 / 0.55 - 1))) < 1e-7


def test_pack_latents_and_ids():
    x = torch.arange(2 * 16 * 4 * 6, dtype=torch.float32).reshape(2, 16, 4, 6)
    p = pack_latents(x)
    assert p.shape == (2, 6, 64)
    # token (i, j) holds the 2x2 patch of every channel, channel-major
    assert torch.equal(p[1, 1 * 3 + 2].reshape(16, 2, 2), x[1, :, 2:4, 4:6])
    ids = latent_image_ids(2, 2, 3)
    assert ids.shape == (2, 6, 3) and ids[0, 5].tolist() == [0.0, 1.0, 2.0] and torch.equal(ids[0], ids[1])


def test_rope_tables_match_embed_nd():
    ids = torch.cat([torch.zeros(5, 3), latent_image_ids(1, 3, 4)[0]], dim=0)
    cos, sin = rope_tables(ids, (16, 56, 56))
    ref = embed_nd(ids[None], (16, 56, 56))[0, 0]  # [S, 64, 2, 2] = [[cos, -sin], [sin, cos]]
    assert cos.shape == (17, 64)
    assert torch.allclose(cos, ref[..., 0, 0], atol=1e-7) and torch.allclose(sin, ref[..., 1, 0], atol=1e-7)
    assert torch.allclose(-sin, ref[..., 0, 1], atol=1e-7)


def test_flux_args_struct_layout():
    # mirrors include/ecad_b200.h: 3 ints (+pad), 14 pointers, int (+pad), 2 pointers, 2 pointer arrays, dead mask
    assert C.sizeof(_lib.EcadkFluxArgs) == 16 + 14 * 8 + 8 + 5 * 8 + 8  # ... + rope_sample_stride (+pad)
    assert C.sizeof(_lib.EcadkFluxDesc) == 20
    assert C.sizeof(_lib.EcadkFluxDoubleWeights) == 20 * 8 and C.sizeof(_lib.EcadkFluxSingleWeights) == 8 * 8
    assert _lib.EcadkFluxArgs.mod_stride.offset == 16 + 14 * 8
