"""The CPU oracle itself: semantic invariants of the reference's cached block (invariants 1-5)
on a tiny configuration (same code path, small dims so the whole file runs in seconds), and its decision trace
against the reference's golden MAC vectors."""
import numpy as np
import pytest
import torch

from ecad_b200.macs import PixArtShape, macs_per_step
from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings
from golden_util import flags_of, row_by_path
from oracle.pixart_oracle import (OracleConfig, OracleDPMSolver, OracleSchedule, PixArtOracle, generate_latents)

TINY = dict(num_attention_heads=2, attention_head_dim=8, num_layers=28, cross_attention_dim=16, sample_size=8,
            caption_channels=32)


@pytest.fixture(scope="module")
def tiny_sd():
    return random_init_state_dict(PixArtConfig(**TINY), seed=0)


def _run(sd, flags, custom=None, gate_step=None, batch=1, seed=0):
    emb = synthetic_prompt_embeddings(batch, text_tokens=12, channels=32, seed=1)
    model = PixArtOracle(sd, OracleConfig(**TINY), OracleSchedule.from_flags(flags, custom))
    noise = torch.randn(batch, 4, 8, 8, generator=torch.Generator().manual_seed(seed))
    out = generate_latents(model, emb["prompt_embeds"], emb["prompt_attention_mask"], emb["negative_prompt_embeds"],
                           emb["negative_prompt_attention_mask"], noise, flags.shape[0], tgate_gate_step=gate_step,
                           record_steps=True)
    return out, model


@pytest.mark.parametrize("path", [
    "schedules_in_paper/pixart_alpha_256/ours_fast.json",
    "schedules_in_paper/pixart_alpha_256/ours_fastest.json",
    "alpha_cache_schedules/gen_recompute_all_every_n/recompute_all_every_002.json",
    "alpha_cache_schedules/gen_tgate/tgate_m_010_sp_001_fi_001_warmup_002.json",
    "alpha_cache_schedules/gen_tgate_without_ca_avg/tgate_without_ca_avg_m_010_sp_001_fi_001_warmup_002.json",
    "population_initialization/pixart_alpha_256x256/gen_000/candidates/cand_017.json",
])
def test_oracle_decision_trace_reproduces_golden_macs(tiny_sd, path):
    """Running the oracle's cached forward (tiny dims, real schedule) marks exactly the sub-blocks whose MACs the
    reference recorded at full size."""
    hits = [r for r in __import__("golden_util").rows() if r["path"].endswith(path.split("/")[-1])
            and path.split("/")[-2] in r["path"]]
    r = hits[0]
    pipe = (r["config"] or {}).get("pipeline") or {}
    gate = pipe.get("kwargs", {}).get("gate_step") if pipe.get("name") == "tgate" else None
    custom = None
    if r["custom"]:
        custom = {"name": r["custom"]["attn"], "kwargs": {"gate_step": r["custom"]["gate_step"]}}
    _, model = _run(tiny_sd, flags_of(r), custom, gate)
    trace = model.trace.to_numpy(20, 28)
    macs = macs_per_step(trace, PixArtShape(tokens=r["tokens"]), 2, gate)
    assert list(map(int, macs)) == r["macs"]


def test_default_schedule_equals_uncached_model(tiny_sd):
    """All-true schedule == never reading a cache (alpha_cache_schedules/gen_default/default.json semantics)."""
    flags = np.ones((6, 28, 3), bool)
    a, model = _run(tiny_sd, flags)
    assert model.trace.to_numpy(6, 28).all() and not model.warnings
    # perturbing the caches between steps cannot change anything when every flag is True
    emb = synthetic_prompt_embeddings(1, text_tokens=12, channels=32, seed=1)
    m2 = PixArtOracle(tiny_sd, OracleConfig(**TINY), OracleSchedule.from_flags(flags))
    orig = m2.block_forward

    def poisoned(b, *args):
        for c in m2.caches:
            for n in ("attn1", "attn2", "ff"):
                if getattr(c, n) is not None:
                    setattr(c, n, torch.full_like(getattr(c, n), 1e6))
        return orig(b, *args)

    m2.block_forward = poisoned
    noise = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    b = generate_latents(m2, emb["prompt_embeds"], emb["prompt_attention_mask"], emb["negative_prompt_embeds"],
                         emb["negative_prompt_attention_mask"], noise, 6)
    assert torch.equal(a["latents"], b["latents"])


def test_cache_holds_ungated_output_and_is_regated_each_step(tiny_sd):
    """Invariant 2: the cache stores the UN-gated sub-block output; on reuse the CURRENT step's gate multiplies it."""
    flags = np.ones((2, 28, 3), bool)
    flags[1, 0, 0] = False  # reuse attn1 of block 0 at step 1
    cfg = OracleConfig(**TINY)
    sched = OracleSchedule.from_flags(flags)
    m = PixArtOracle(tiny_sd, cfg, sched)
    emb = synthetic_prompt_embeddings(2, text_tokens=12, channels=32, seed=2)
    x = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(1))
    m.forward(x, emb["prompt_embeds"], torch.tensor([999, 999]), None, emb["prompt_attention_mask"])
    cached = m.caches[0].attn1.clone()
    sched.per_step_callback(0, 999)
    # step 1 at a different timestep: gate differs, cached tensor is returned untouched and re-cached
    m.forward(x, emb["prompt_embeds"], torch.tensor([500, 500]), None, emb["prompt_attention_mask"])
    assert torch.equal(m.caches[0].attn1, cached)
    assert m.trace.to_numpy(2, 28)[1, 0].tolist() == [0, 1, 1]


def test_warning_and_recompute_on_empty_cache(tiny_sd):
    flags = np.zeros((1, 28, 3), bool)
    _, model = _run(tiny_sd, flags)
    assert model.trace.to_numpy(1, 28).all()
    assert len(model.warnings) == 84 and model.warnings[0] == "WARNING: No cached attn1 found. Recomputing."


def test_reset_after_last_step(tiny_sd):
    flags = np.ones((3, 28, 3), bool)
    _, model = _run(tiny_sd, flags)
    assert model.cache_schedule.curr_step == 0
    assert all(c.attn1 is None and c.attn2 is None and c.ff is None for c in model.caches)


def test_mask_to_bias_and_masked_keys_do_not_matter(tiny_sd):
    cfg = OracleConfig(**TINY)
    m = PixArtOracle(tiny_sd, cfg, OracleSchedule.from_flags(np.ones((1, 28, 3), bool)))
    mask = torch.tensor([[1, 1, 0, 0]])
    assert m.mask_to_bias(mask).tolist() == [[[0.0, 0.0, -10000.0, -10000.0]]]
    emb = synthetic_prompt_embeddings(1, text_tokens=12, channels=32, seed=4)
    x = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(2))
    t = torch.tensor([400])
    a = m.forward(x, emb["prompt_embeds"], t, None, emb["prompt_attention_mask"])
    e2 = emb["prompt_embeds"].clone()
    L = int(emb["prompt_attention_mask"].sum())
    e2[:, L:] = 123.0  # garbage in masked positions
    m.reset_cache()
    b = m.forward(x, e2, t, None, emb["prompt_attention_mask"])
    assert torch.allclose(a, b, atol=1e-5, rtol=1e-5)


def test_dpm_solver_timesteps_and_exact_last_step():
    s = OracleDPMSolver(20)
    assert s.timesteps.tolist() == [999, 949, 899, 849, 799, 749, 699, 649, 599, 549, 500, 450, 400, 350, 300, 250,
                                    200, 150, 100, 50]
    assert float(s.sigmas[-1]) == 0.0 and len(s.sigmas) == 21
    # with an exact epsilon the last step returns x0 itself
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(1, 4, 8, 8, generator=g)
    eps = torch.randn(1, 4, 8, 8, generator=g)
    x = None
    for i in range(20):
        a, sg = s._alpha_sigma(s.sigmas[i])
        if x is None:
            x = a * x0 + sg * eps
        x = s.step((x - a * x0) / sg, x)
    assert torch.allclose(x, x0, atol=1e-4)


def test_precision_policy_bf16_operands_fp32_stream(tiny_sd):
    """The precision study behind DESIGN.md: rounding GEMM operands/activations to bf16 while keeping the residual
    stream fp32 stays far inside the tolerance; this is the policy the CUDA path implements."""
    flags = np.ones((8, 28, 3), bool)
    flags[2:6, ::2] = False
    bf = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
    ref, _ = _run(tiny_sd, flags)
    emb = synthetic_prompt_embeddings(1, text_tokens=12, channels=32, seed=1)
    sd_b = {k: (bf(v) if (v.ndim >= 2 and "scale_shift" not in k) else v) for k, v in tiny_sd.items()}
    m = PixArtOracle(sd_b, OracleConfig(**TINY), OracleSchedule.from_flags(flags), round_act=bf, round_res=None)
    noise = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    got = generate_latents(m, emb["prompt_embeds"], emb["prompt_attention_mask"], emb["negative_prompt_embeds"],
                           emb["negative_prompt_attention_mask"], noise, 8)
    cos = torch.nn.functional.cosine_similarity(got["latents"].flatten(), ref["latents"].flatten(), dim=0)
    assert float(cos) > 0.999
