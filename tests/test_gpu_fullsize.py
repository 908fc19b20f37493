"""Parity at BASELINE.json's FULL sizes, where the CPU oracle cannot run the whole workload: size-independent
properties plus oracle spot checks of individual prompts.

  * config 2 (population evaluation, 100 prompts = 200 samples per forward): samples are independent, so two prompts
    of the batch-100 generation must match the oracle run on those two prompts alone, and the whole batch must match
    the same prompts generated four at a time by the same library (batch invariance).
  * config 5 (FLUX.1-dev width and depth, 1024 px, batch 4; 12 B parameters - no oracle run possible): a step that
    reuses every component on identical inputs must reproduce the dense step (cache bookkeeping across 30 GB of cache
    slots); the two halves of a batch fed identical inputs must agree; nothing may be NaN.
"""
import numpy as np
import pytest
import torch

from golden_util import flags_of, row_by_path, schedule_of

pytestmark = pytest.mark.gpu


def _cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))


def test_config2_batch100_matches_oracle_on_sampled_prompts_and_is_batch_invariant(cuda_device):
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings
    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle, generate_latents

    row = row_by_path("population_initialization/pixart_alpha_256x256/gen_000/candidates/cand_003.json")
    flags = flags_of(row)
    sd = random_init_state_dict(PixArtConfig(), 0)
    B = 100
    emb = synthetic_prompt_embeddings(B, seed=1)
    trace = []
    gen = B200PixArtAlphaImageGenerator(
        cache_schedule=schedule_of(row), start_seed=0, state_dict=sd,
        additional_callbacks=[lambda s, t, **kw: trace.append(gen.diffusion_pipeline.transformer.last_executed.copy())])
    full = gen.generate_images(emb)[0].cpu()
    assert full.shape == (B, 4, 32, 32) and torch.isfinite(full).all()

    # (a) oracle on prompts 7 and 93 alone, with the noise rows the batch-100 draw gave them
    noise = torch.randn(B, 4, 32, 32, generator=torch.Generator().manual_seed(0))
    pick = [7, 93]
    sub = {k: v[pick] for k, v in emb.items()}
    model = PixArtOracle(sd, OracleConfig(), OracleSchedule.from_flags(flags))
    ref = generate_latents(model, sub["prompt_embeds"], sub["prompt_attention_mask"], sub["negative_prompt_embeds"],
                           sub["negative_prompt_attention_mask"], noise[pick], flags.shape[0])["latents"]
    assert np.array_equal(np.stack(trace), model.trace.to_numpy(flags.shape[0], 28)), "decisions differ from the oracle"
    got = full[pick]
    rel = float((got - ref).abs().max() / ref.abs().max())
    assert _cos(got, ref) >= 0.999 and rel <= 1e-2, (_cos(got, ref), rel)

    # (b) batch invariance: the same prompts four at a time (another GEMM kernel variant and grid)
    for lo in (0, 48, 96):
        part = {k: v[lo:lo + 4] for k, v in emb.items()}
        small = gen.diffusion_pipeline(
            prompt_embeds=part["prompt_embeds"], prompt_attention_mask=part["prompt_attention_mask"],
            negative_prompt_embeds=part["negative_prompt_embeds"],
            negative_prompt_attention_mask=part["negative_prompt_attention_mask"], latents=noise[lo:lo + 4].clone(),
            num_inference_steps=flags.shape[0], callback=gen._call_callbacks_wrapper)[0].cpu()
        # measured 0.0 (bit-identical) on the fixed build; up to 4.5e-4 over the staging race of attn_pair2_kernel that
        # tests/test_gpu_determinism.py now pins, 3e-3 to 4e-3 once the faster MMA issue made it frequent
        rel = float((small - full[lo:lo + 4]).abs().max() / full.abs().max())
        assert rel <= 1e-4 and _cos(small, full[lo:lo + 4]) >= 0.99999, (lo, rel)


def test_config5_flux_full_size_reuse_step_reproduces_dense_step(cuda_device):
    from ecad_b200.flux_pipeline import latent_image_ids
    from ecad_b200.flux_transformer import B200FluxTransformer2D
    from ecad_b200.schedule import FluxCacheSchedule
    from ecad_b200.transformer import SequentialDiTScheduler
    from ecad_b200.weights import FluxConfig

    free, _ = torch.cuda.mem_get_info()
    if free < 70e9:
        pytest.skip("needs ~60 GB of HBM (FLUX.1-dev weights + config-5 caches)")
    cfg = FluxConfig()
    flags = np.ones((3, 57, 3), bool)
    flags[1] = False  # step 1 reuses all 171 components
    sched = FluxCacheSchedule.from_numpy(flags, 3, 19, 38, "dense_reuse_dense")
    model = B200FluxTransformer2D.from_random_init(SequentialDiTScheduler(3), sched, cfg, seed=0, on_device=True)
    B, N, T = 4, 4096, 512
    g = torch.Generator(device="cuda").manual_seed(1)
    lat = torch.randn(2, N, 64, device="cuda", generator=g).repeat(2, 1, 1)       # samples 2,3 repeat samples 0,1
    emb = (torch.randn(2, T, 4096, device="cuda", generator=g) * 0.2).repeat(2, 1, 1)
    pooled = (torch.randn(2, 768, device="cuda", generator=g) * 0.2).repeat(2, 1)
    ids, tids = latent_image_ids(B, 64, 64), torch.zeros(B, T, 3)
    t, guid = torch.full((B,), 0.7, device="cuda"), torch.full((B,), 3.5, device="cuda")
    outs = []
    for step in range(3):
        o = model(lat, emb, pooled, t, ids, tids, guid, return_dict=False)[0].float().clone()
        assert torch.isfinite(o).all()
        assert int(model.last_executed.sum()) == (171 if step != 1 else 0)
        outs.append(o)
        sched.per_step_callback(step)
    # identical inputs: the all-reuse step re-adds the cached (bf16) component outputs -> same result up to that rounding
    rel = float((outs[1] - outs[0]).abs().max() / outs[0].abs().max())
    assert rel <= 1e-2 and _cos(outs[1], outs[0]) >= 0.9999, rel
    assert torch.equal(outs[2], outs[0])                 # recomputing everything again is deterministic
    assert torch.equal(outs[0][:2], outs[0][2:])         # sample independence at full size
    assert int(model.last_dead.sum()) > 0                # last step: every epilogue cache store is dead


@pytest.mark.parametrize("name,sample_size,batch,text_tokens", [("config3", 64, 16, 120), ("config4", 128, 8, 300)])
def test_config3_config4_full_model_forward_matches_oracle_on_sampled_rows(cuda_device, name, sample_size, batch,
                                                                            text_tokens):
    """BASELINE configs 3 (PixArt-alpha 512x512, batch 16) and 4 (PixArt-sigma 1024x1024, 300-token captions, batch 8)
    at FULL model depth and batch: a dense forward, then a forward that reuses everything.  The oracle runs only the
    CFG pair of one prompt (samples are independent); the other rows are covered by the repeat/independence check."""
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.transformer import B200PixArtTransformer2D, SequentialDiTScheduler
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings
    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle

    cfg = PixArtConfig(sample_size=sample_size, use_additional_conditions=False)
    sd = random_init_state_dict(cfg, seed=5)
    emb = synthetic_prompt_embeddings(batch, text_tokens=text_tokens, seed=6)
    lat = torch.randn(batch, 4, sample_size, sample_size, generator=torch.Generator().manual_seed(7))
    lat[batch - 1] = lat[0]                                  # last prompt repeats prompt 0 (independence check)
    for k in emb:
        emb[k][batch - 1] = emb[k][0]
    x_in = torch.cat([lat, lat])
    e_in = torch.cat([emb["negative_prompt_embeds"], emb["prompt_embeds"]])
    m_in = torch.cat([emb["negative_prompt_attention_mask"], emb["prompt_attention_mask"]])
    S = 2 * batch
    ts = torch.full((S,), 649, dtype=torch.int64)
    flags = np.ones((2, 28, 3), bool)
    flags[1] = False
    sched = PixArtCacheSchedule.from_numpy(flags, 2, 28, "dense_then_reuse")
    tr = B200PixArtTransformer2D(sd, cfg, SequentialDiTScheduler(2), sched)
    kw = dict(encoder_hidden_states=e_in.cuda(), encoder_attention_mask=m_in.cuda(), timestep=ts.cuda(),
              added_cond_kwargs={"resolution": None, "aspect_ratio": None}, return_dict=False)
    out0 = tr(x_in.cuda(), **kw)[0].cpu()
    assert tr.last_executed.all()
    sched.per_step_callback(0)
    out1 = tr(x_in.cuda(), **kw)[0].cpu()
    assert not tr.last_executed.any()

    rows = [0, batch]                                        # uncond + cond sample of prompt 0
    ocfg = OracleConfig(sample_size=sample_size, use_additional_conditions=False)
    oracle = PixArtOracle(sd, ocfg, OracleSchedule.from_flags(flags))
    ref0 = oracle.forward(x_in[rows], e_in[rows], ts[rows], None, m_in[rows])
    oracle.cache_schedule.per_step_callback(0)
    ref1 = oracle.forward(x_in[rows], e_in[rows], ts[rows], None, m_in[rows])
    for got, ref in ((out0, ref0), (out1, ref1)):
        rel = float((got[rows] - ref).abs().max() / ref.abs().max())
        assert rel <= 1e-2 and _cos(got[rows], ref) >= 0.9999, (name, rel, _cos(got[rows], ref))
    assert torch.isfinite(out0).all() and torch.isfinite(out1).all()
    assert torch.equal(out0[batch - 1], out0[0]) and torch.equal(out0[S - 1], out0[batch])  # repeated prompt, same result
