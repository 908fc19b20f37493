"""FLUX oracle (oracle/flux_oracle.py) on a tiny configuration: its observed execution trace under the reference's
shipped FLUX schedules reproduces the reference's recorded per-step MACs, and the cached-block semantics of
cached_flux_transformer_block.py hold (pair caching of the joint attention, pre-GELU proj_mlp cache, reset)."""
import gzip
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from ecad_b200.macs import FluxShape, flux_macs_per_step
from ecad_b200.schedule import trace_decisions
from ecad_b200.weights import FluxConfig, flux_random_init_state_dict
from oracle.flux_oracle import FluxOracle, FluxOracleConfig, FluxOracleSchedule, generate_flux_latents

GOLDEN = Path(__file__).resolve().parent / "golden" / "flux_schedules.json.gz"
TINY = dict(num_attention_heads=2, attention_head_dim=16, in_channels=8, joint_attention_dim=24,
            pooled_projection_dim=12, axes_dims_rope=(4, 6, 6))


@pytest.fixture(scope="module")
def rows():
    with gzip.open(GOLDEN, "rb") as f:
        return json.loads(f.read())["rows"]


@pytest.fixture(scope="module")
def tiny_sd():
    return flux_random_init_state_dict(FluxConfig(**TINY), seed=0)


def _flags(r):
    n = r["S"] * (r["NB"] + r["NS"]) * 3
    return np.unpackbits(np.frombuffer(bytes.fromhex(r["bits"]), np.uint8))[:n].reshape(r["S"], r["NB"] + r["NS"], 3) \
        .astype(bool)


def _inputs(batch=1, n_img=16, n_txt=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(batch, n_img, 8, generator=g)
    emb = torch.randn(batch, n_txt, 24, generator=g) * 0.2
    pooled = torch.randn(batch, 12, generator=g) * 0.2
    side = int(n_img**0.5)
    img_ids = torch.zeros(side, side, 3)
    img_ids[..., 1] = torch.arange(side)[:, None]
    img_ids[..., 2] = torch.arange(side)[None, :]
    img_ids = img_ids.reshape(1, n_img, 3).repeat(batch, 1, 1)
    txt_ids = torch.zeros(batch, n_txt, 3)
    return lat, emb, pooled, img_ids, txt_ids


@pytest.mark.parametrize("suffix", ["flux_256/ours_fast.json", "flux_256/ours_faster.json",
                                    "gen_recompute_all_every_n/recompute_all_every_003.json"])
def test_flux_oracle_trace_reproduces_golden_macs(rows, tiny_sd, suffix):
    r = [r for r in rows if r["path"].endswith(suffix) and r["macs"] is not None][0]
    flags = _flags(r)
    model = FluxOracle(tiny_sd, FluxOracleConfig(**TINY), FluxOracleSchedule.from_flags(flags, r["NB"], r["NS"]))
    generate_flux_latents(model, *_inputs()[1:3], _inputs()[0], *_inputs()[3:], r["S"])
    trace = model.trace.to_numpy(r["S"], r["NB"] + r["NS"])
    assert np.array_equal(trace, trace_decisions(flags))
    macs = flux_macs_per_step(trace, FluxShape(tokens=r["tokens"]))
    assert list(map(int, macs)) == r["macs"]
    assert model.cache_schedule.curr_step == 0  # reset after the last step
    assert all(v is None for c in model.double_cache + model.single_cache for v in c.values())


def test_flux_cache_semantics(tiny_sd):
    cfg = FluxOracleConfig(**TINY)
    flags = np.ones((2, 57, 3), bool)
    flags[1, 0, 0] = False       # reuse the joint attention of double block 0 at step 1
    flags[1, 19 + 3, 1] = False  # reuse proj_mlp of single block 3 at step 1
    sched = FluxOracleSchedule.from_flags(flags, 19, 38)
    m = FluxOracle(tiny_sd, cfg, sched)
    lat, emb, pooled, img_ids, txt_ids = _inputs(batch=2)
    g = torch.full((2,), 5.0)
    m.forward(lat, emb, pooled, torch.full((2,), 1.0), img_ids, txt_ids, g)
    a0, c0 = m.double_cache[0]["attn"].clone(), m.double_cache[0]["context_attn"].clone()
    mlp0 = m.single_cache[3]["proj_mlp"].clone()
    assert a0.shape == (2, 16, 32) and c0.shape == (2, 6, 32)   # cached after to_out / to_add_out, image/text split
    assert mlp0.shape == (2, 22, 128) and float(mlp0.min()) < -0.17  # pre-GELU (GELU(tanh) is bounded below by -0.17)
    sched.per_step_callback(0, 1000.0)
    m.forward(lat * 0.5, emb, pooled, torch.full((2,), 0.8), img_ids, txt_ids, g)
    assert torch.equal(m.double_cache[0]["attn"], a0) and torch.equal(m.double_cache[0]["context_attn"], c0)
    assert torch.equal(m.single_cache[3]["proj_mlp"], mlp0)
    tr = m.trace.to_numpy(2, 57)
    assert tr[1, 0].tolist() == [0, 1, 1] and tr[1, 22].tolist() == [1, 0, 1]
    assert tr.sum() == 2 * 57 * 3 - 2


def test_flux_all_true_equals_uncached_and_rope_is_a_rotation(tiny_sd):
    from oracle.flux_oracle import apply_rope, embed_nd
    cfg = FluxOracleConfig(**TINY)
    lat, emb, pooled, img_ids, txt_ids = _inputs()
    flags = np.ones((3, 57, 3), bool)
    m = FluxOracle(tiny_sd, cfg, FluxOracleSchedule.from_flags(flags, 19, 38))
    out = generate_flux_latents(m, emb, pooled, lat, img_ids, txt_ids, 3)
    assert out.shape == lat.shape and torch.isfinite(out).all() and not m.warnings
    rope = embed_nd(torch.cat((txt_ids, img_ids), dim=1), cfg.axes_dims_rope)
    assert rope.shape == (1, 1, 22, 8, 2, 2)
    x = torch.randn(1, 2, 22, 16)
    y = apply_rope(x, rope)
    assert torch.allclose(y.norm(dim=-1), x.norm(dim=-1), atol=1e-5)  # rotations preserve the norm
    assert torch.allclose(y[:, :, :6], x[:, :, :6], atol=1e-6)         # text ids are all zero -> identity
