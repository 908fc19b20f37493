"""FLUX parity on the GPU: B200FluxTransformer2D / B200FluxImageGenerator (C ABI, sm_100a kernels) against the CPU
fp32 oracle (oracle/flux_oracle.py) on identical random-init weights, synthetic embeddings and CPU-generator noise.

Bars (BASELINE.json north_star): decisions bit-exact; per-step max|x - x_ref| / max|x_ref| <= 1e-2 for the model
output (bf16 kernels vs fp32 oracle); cosine similarity >= 0.999 on the final latent.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# a narrow FLUX (4 heads x 128 = 512 wide) so the CPU oracle finishes in seconds; same block structure
SMALL = dict(num_attention_heads=4, attention_head_dim=128, num_layers=2, num_single_layers=3, in_channels=64,
             joint_attention_dim=256, pooled_projection_dim=64, axes_dims_rope=(16, 56, 56))


def _cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _embeds(batch, text_tokens, cfg, seed=1):
    g = torch.Generator().manual_seed(seed)
    return {"prompt_embeds": torch.randn(batch, text_tokens, cfg["joint_attention_dim"], generator=g) * 0.2,
            "pooled_prompt_embeds": torch.randn(batch, cfg["pooled_projection_dim"], generator=g) * 0.2}


def _schedule_flags(steps, rows, seed=0):
    """random reuse pattern incl. a component that is never recomputed after step 0 and a flag that asks for reuse at
    step 0 (cache empty -> the fallback recomputes, cached_flux_transformer_block.py:57-61)."""
    rng = np.random.default_rng(seed)
    flags = rng.random((steps, rows, 3)) < 0.55
    flags[0] = True
    flags[0, 1, 0] = False   # reuse requested with an empty cache: must still execute
    flags[0, rows - 1, 1] = False
    flags[1:, 0, 1] = False  # full_ff of block 0 reused for the whole generation
    flags[2] = False         # a step that reuses everything
    return flags


def test_flux_cached_generation_matches_oracle(cuda_device):
    from ecad_b200.image_generator import B200FluxImageGenerator
    from ecad_b200.schedule import FluxCacheSchedule
    from ecad_b200.weights import FluxConfig, flux_random_init_state_dict
    from oracle.flux_oracle import FluxOracle, FluxOracleConfig, FluxOracleSchedule, generate_flux_latents
    from ecad_b200.flux_pipeline import latent_image_ids, pack_latents

    cfg = FluxConfig(**SMALL)
    steps, rows = 6, cfg.num_layers + cfg.num_single_layers
    flags = _schedule_flags(steps, rows)
    sd = flux_random_init_state_dict(cfg, seed=0)
    B, T, height, width = 2, 64, 256, 192  # N = 16 * 12 = 192 image tokens, S = 256
    emb = _embeds(B, T, SMALL)
    trace, per_step = [], []

    def grab(step, timestep, **kw):
        tr = gen.diffusion_pipeline.transformer
        trace.append(tr.last_executed.copy())
        per_step.append(tr._ws["out"].view(B, -1, 64).float().cpu().clone())

    gen = B200FluxImageGenerator(
        cache_schedule=FluxCacheSchedule.from_numpy(flags, steps, cfg.num_layers, cfg.num_single_layers, "rand",
                                                    top_level_config={"height": height, "width": width}),
        start_seed=0, state_dict=sd, model_config=cfg, additional_callbacks=[grab])
    got = gen.generate_images(emb, images_per_prompt=1)[0].cpu()
    tr = gen.diffusion_pipeline.transformer
    assert tr.launches > 0 and len(tr.warnings) == 2  # the two forced cache misses warn like the reference

    ocfg = FluxOracleConfig(**SMALL)
    model = FluxOracle(sd, ocfg, FluxOracleSchedule.from_flags(flags, cfg.num_layers, cfg.num_single_layers))
    outs = []
    fwd = model.forward
    model.forward = lambda *a, **k: (outs.append(fwd(*a, **k)), outs[-1])[1]
    noise = torch.randn(B, 16, 2 * (height // 16), 2 * (width // 16), generator=torch.Generator().manual_seed(0))
    lat0 = pack_latents(noise)
    ids = latent_image_ids(B, height // 16, width // 16)
    ref = generate_flux_latents(model, emb["prompt_embeds"], emb["pooled_prompt_embeds"], lat0, ids,
                                torch.zeros(B, T, 3), steps, guidance_scale=5.0)
    assert np.array_equal(np.stack(trace), model.trace.to_numpy(steps, rows)), "decision trace differs from the oracle"
    assert len(model.warnings) == 2
    for s in range(steps):
        assert _rel(per_step[s], outs[s]) <= 1e-2, (s, _rel(per_step[s], outs[s]))
    assert _cos(got, ref) >= 0.999 and _rel(got, ref) <= 1e-2, (_cos(got, ref), _rel(got, ref))
    # a second generation on the same resident model starts from an empty cache again (reset LAST callback)
    trace.clear(); per_step.clear()
    got2 = gen.generate_images(emb, images_per_prompt=1)[0].cpu()
    assert torch.equal(got, got2)


def test_flux_full_width_forward_matches_oracle(cuda_device):
    """FLUX.1-dev width (24 heads x 128 = 3072, T5 4096 channels) with one double-stream and one single-stream block:
    a dense forward, then a forward that reuses every component."""
    from ecad_b200.flux_pipeline import latent_image_ids
    from ecad_b200.flux_transformer import B200FluxTransformer2D
    from ecad_b200.schedule import FluxCacheSchedule
    from ecad_b200.transformer import SequentialDiTScheduler
    from ecad_b200.weights import FluxConfig, flux_random_init_state_dict
    from oracle.flux_oracle import FluxOracle, FluxOracleConfig, FluxOracleSchedule

    kw = dict(num_layers=1, num_single_layers=1)
    cfg = FluxConfig(**kw)
    sd = flux_random_init_state_dict(cfg, seed=3)
    flags = np.ones((2, 2, 3), bool)
    flags[1] = False
    B, N, T = 1, 256, 256
    g = torch.Generator().manual_seed(7)
    lat = torch.randn(B, N, 64, generator=g)
    emb = _embeds(B, T, dict(joint_attention_dim=4096, pooled_projection_dim=768), seed=8)
    ids, tids = latent_image_ids(B, 16, 16), torch.zeros(B, T, 3)
    t, guid = torch.full((B,), 0.9), torch.full((B,), 3.5)

    sched = FluxCacheSchedule.from_numpy(flags, 2, 1, 1, "dense_then_reuse")
    model = B200FluxTransformer2D(sd, cfg, SequentialDiTScheduler(2), sched)
    oracle = FluxOracle(sd, FluxOracleConfig(**kw), FluxOracleSchedule.from_flags(flags, 1, 1))
    for step, scale in enumerate((1.0, 0.7)):
        got = model(lat * scale, emb["prompt_embeds"], emb["pooled_prompt_embeds"], t, ids, tids, guid,
                    return_dict=False)[0].float().cpu()
        ref = oracle.forward(lat * scale, emb["prompt_embeds"], emb["pooled_prompt_embeds"], t, ids, tids, guid)
        assert model.last_executed.sum() == (6 if step == 0 else 0)
        assert _rel(got, ref) <= 1e-2 and _cos(got, ref) >= 0.9999, (step, _rel(got, ref), _cos(got, ref))
        sched.per_step_callback(step)
        oracle.cache_schedule.per_step_callback(step)


def test_flux_full_width_shipped_schedule_generation(cuda_device):
    """FLUX.1-dev WIDTH (24 heads x 128 = 3072, T5 4096 channels, 512 text tokens) with 4 double-stream + 8 single-stream
    blocks, a cached generation under the decision rows of the paper's shipped schedule
    (schedules/schedules_in_paper/flux_256/ours_fast.json: rows of double blocks 0-3 and single blocks 0-7, first 5
    steps) at 256x256 so the CPU oracle finishes, batch 2 with PER-SAMPLE img_ids (the second sample's positions are a
    shifted crop): decisions bit-exact, every step's model output and the final latents within the bf16 bars."""
    import gzip
    import json
    from pathlib import Path

    from ecad_b200.flux_pipeline import FlowMatchEulerDiscrete, calculate_shift, latent_image_ids, pack_latents
    from ecad_b200.flux_transformer import B200FluxTransformer2D
    from ecad_b200.schedule import FluxCacheSchedule
    from ecad_b200.transformer import SequentialDiTScheduler
    from ecad_b200.weights import FluxConfig, flux_random_init_state_dict
    from oracle.flux_oracle import FluxOracle, FluxOracleConfig, FluxOracleSchedule

    rows = json.loads(gzip.open(Path(__file__).parent / "golden" / "flux_schedules.json.gz").read())["rows"]
    r = [r for r in rows if r["path"] == "schedules_in_paper/flux_256/ours_fast.json"][0]
    full = (np.unpackbits(np.frombuffer(bytes.fromhex(r["bits"]), np.uint8))[: r["S"] * 57 * 3]
            .reshape(r["S"], 57, 3).astype(bool))
    ND, NS, steps = 4, 8, 5
    flags = np.concatenate([full[:steps, :ND], full[:steps, 19:19 + NS]], axis=1)
    assert 0.2 < flags[1:].mean() < 0.9  # the slice really mixes executed and reused components

    kw = dict(num_layers=ND, num_single_layers=NS)
    cfg = FluxConfig(**kw)
    sd = flux_random_init_state_dict(cfg, seed=2)
    B, T, hw = 2, 512, 16  # 256x256 px -> 16 x 16 = 256 image tokens
    N = hw * hw
    emb = _embeds(B, T, dict(joint_attention_dim=4096, pooled_projection_dim=768), seed=4)
    lat = pack_latents(torch.randn(B, 16, 2 * hw, 2 * hw, generator=torch.Generator().manual_seed(6)))
    ids = latent_image_ids(B, hw, hw)
    ids[1, :, 1] += 7.0   # sample 1: another crop of the position grid -> its own rotation table
    ids[1, :, 2] += 3.0
    tids = torch.zeros(B, T, 3)
    guid = torch.full((B,), 3.5)
    sched_host = FlowMatchEulerDiscrete()
    c = sched_host.config
    sched_host.set_timesteps(steps, mu=calculate_shift(N, c.base_image_seq_len, c.max_image_seq_len, c.base_shift,
                                                      c.max_shift))
    sig = sched_host.sigmas

    sched = FluxCacheSchedule.from_numpy(flags, steps, ND, NS, "ours_fast[4+8 blocks, 5 steps]")
    model = B200FluxTransformer2D(sd, cfg, SequentialDiTScheduler(steps), sched)
    oracle = FluxOracle(sd, FluxOracleConfig(**kw), FluxOracleSchedule.from_flags(flags, ND, NS))
    x_gpu, x_ref = lat.clone(), lat.clone()
    for s in range(steps):
        t = torch.full((B,), float(sig[s]))
        got = model(x_gpu.cuda(), emb["prompt_embeds"].cuda(), emb["pooled_prompt_embeds"].cuda(), t.cuda(), ids, tids,
                    guid.cuda(), return_dict=False)[0].float().cpu()
        ref = oracle.forward(x_ref, emb["prompt_embeds"], emb["pooled_prompt_embeds"], t, ids, tids, guid)
        assert np.array_equal(model.last_executed, oracle.trace.to_numpy(steps, ND + NS)[s]), s
        assert _rel(got, ref) <= 1e-2, (s, _rel(got, ref))
        for b in range(B):  # per sample: a wrong rotation table for sample 1 must not hide behind sample 0
            assert _rel(got[b], ref[b]) <= 1e-2, (s, b)
        dt = float(sig[s + 1]) - float(sig[s])
        x_gpu, x_ref = x_gpu + dt * got, x_ref + dt * ref
        sched.per_step_callback(s)
        oracle.cache_schedule.per_step_callback(s)
    assert _cos(x_gpu, x_ref) >= 0.999 and _rel(x_gpu, x_ref) <= 1e-2
    # the per-sample tables matter: the two samples use different rotations
    assert model._ws["args"].rope_sample_stride == (N + T) * 64
