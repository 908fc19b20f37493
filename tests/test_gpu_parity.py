"""Model-level parity on the GPU: the CUDA path (through the C ABI and the reference-shaped host API) against the
CPU fp32 oracle on identical random-init weights, synthetic caption embeddings and fixed-seed noise.

Bars (BASELINE.json north_star):
  * per-step / per-block compute-or-reuse decisions: BIT-EXACT
  * latents: bf16 tensor-core operands + fp32 accumulators / fp32 residual stream vs the fp32 oracle:
      per-step   max|x - x_ref| / max|x_ref| <= 1e-2   (asserted at 5e-3; measured 3.5e-3)
      final      cosine similarity >= 0.999            (asserted at 0.9999; measured ~0.99999)
"""
import numpy as np
import pytest
import torch

from golden_util import flags_of, row_by_path, schedule_of

pytestmark = pytest.mark.gpu

# north_star's bars are 1e-2 per step / cosine 0.999; the assertions are held tighter (measured on B200: 3.5e-3 per
# step, cosine 0.99999) so that a numerical regression is caught long before it reaches the stated tolerance
PER_STEP_REL_MAXABS = 5e-3
FINAL_COS = 0.9999


@pytest.fixture(scope="module")
def weights():
    from ecad_b200.weights import PixArtConfig, random_init_state_dict
    return random_init_state_dict(PixArtConfig(), seed=0)


def _oracle_run(sd, flags, embeds, latents, custom=None, gate_step=None):
    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle, generate_latents
    torch.set_num_threads(max(1, torch.get_num_threads()))
    model = PixArtOracle(sd, OracleConfig(), OracleSchedule.from_flags(flags, custom))
    out = generate_latents(model, embeds["prompt_embeds"], embeds["prompt_attention_mask"],
                           embeds["negative_prompt_embeds"], embeds["negative_prompt_attention_mask"],
                           latents, flags.shape[0], tgate_gate_step=gate_step, record_steps=True)
    return out, model.trace.to_numpy(flags.shape[0], flags.shape[1])


def _cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))


def test_single_forward_dense(cuda_device, weights):
    """One dense forward (step 0, all sub-blocks executed) == the oracle's forward."""
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.transformer import B200PixArtTransformer2D, SequentialDiTScheduler
    from ecad_b200.weights import PixArtConfig, synthetic_prompt_embeddings
    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle

    emb = synthetic_prompt_embeddings(2, seed=3)
    g = torch.Generator().manual_seed(11)
    lat = torch.randn(2, 4, 32, 32, generator=g)
    x_in = torch.cat([lat, lat])
    e_in = torch.cat([emb["negative_prompt_embeds"], emb["prompt_embeds"]])
    m_in = torch.cat([emb["negative_prompt_attention_mask"], emb["prompt_attention_mask"]])
    ts = torch.full((4,), 949, dtype=torch.int64)

    flags = np.ones((20, 28, 3), bool)
    ref = PixArtOracle(weights, OracleConfig(), OracleSchedule.from_flags(flags)).forward(x_in, e_in, ts, None, m_in)

    tr = B200PixArtTransformer2D(weights, PixArtConfig(), SequentialDiTScheduler(20), PixArtCacheSchedule.default())
    out = tr(x_in.cuda(), encoder_hidden_states=e_in.cuda(), encoder_attention_mask=m_in.cuda(), timestep=ts.cuda(),
             added_cond_kwargs={"resolution": None, "aspect_ratio": None}, return_dict=False)[0].cpu()
    assert out.shape == ref.shape == (4, 8, 32, 32)
    assert tr.last_executed.all()
    rel = float((out - ref).abs().max() / ref.abs().max())
    assert rel < 5e-3, rel
    assert _cos(out, ref) > 0.99999


@pytest.mark.parametrize("schedule_file,batch", [
    ("schedules_in_paper/pixart_alpha_256/ours_fast.json", 1),
    ("alpha_cache_schedules/gen_default/default.json", 1),
    ("schedules_in_paper/pixart_alpha_256/ours_fastest.json", 2)])
def test_generation_matches_oracle(cuda_device, weights, schedule_file, batch):
    """BASELINE config 1: PixArt-alpha 256x256, 20 DPM steps, cached schedule, through the ImageGenerator API."""
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.weights import synthetic_prompt_embeddings

    row = row_by_path(schedule_file)
    flags = flags_of(row)
    emb = synthetic_prompt_embeddings(batch, seed=1)

    traces, per_step = [], []

    def spy(step, timestep, latents=None, **kw):
        traces.append(gen.diffusion_pipeline.transformer.last_executed.copy())
        per_step.append(latents.detach().cpu().clone())

    gen = B200PixArtAlphaImageGenerator(cache_schedule=schedule_of(row), start_seed=0, state_dict=weights,
                                        additional_callbacks=[spy])
    got = gen.generate_images(emb, images_per_prompt=1)[0].cpu()

    noise = torch.randn(batch, 4, 32, 32, generator=torch.Generator().manual_seed(0))
    ref, ref_trace = _oracle_run(weights, flags, emb, noise)

    # decisions: bit-exact, every step, every block, every component
    assert np.array_equal(np.stack(traces), ref_trace)
    # per-step bound
    for s, (a, b) in enumerate(zip(per_step, ref["per_step"])):
        rel = float((a - b).abs().max() / b.abs().max())
        assert rel <= PER_STEP_REL_MAXABS, (s, rel)
    assert _cos(got, ref["latents"]) >= FINAL_COS
    # the generator resets counters and caches after the last step (image_generator.py:193-202)
    assert gen.cache_schedule.curr_step == 0
    assert not gen.diffusion_pipeline.transformer._has_cache.any()


@pytest.mark.parametrize("sample_size,text_tokens", [(64, 120), (32, 300), (32, 200), (64, 200)])
def test_single_forward_other_shapes(cuda_device, sample_size, text_tokens):
    """BASELINE configs 3 / 4 shapes on one forward: PixArt-alpha 512x512 (N = 1024 image tokens, streamed
    self-attention) and PixArt-sigma's 300-token captions (cross-attention keys padded to 384); 200-token captions
    (keys padded to 256 WITH a mask bias: the one key count the 256-query kernel has no room for - streamed)."""
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.transformer import B200PixArtTransformer2D, SequentialDiTScheduler
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings
    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle

    cfg = PixArtConfig(sample_size=sample_size, num_layers=4)
    sd = random_init_state_dict(cfg, seed=2)
    emb = synthetic_prompt_embeddings(1, text_tokens=text_tokens, seed=3)
    lat = torch.randn(1, 4, sample_size, sample_size, generator=torch.Generator().manual_seed(4))
    x_in = torch.cat([lat, lat])
    e_in = torch.cat([emb["negative_prompt_embeds"], emb["prompt_embeds"]])
    m_in = torch.cat([emb["negative_prompt_attention_mask"], emb["prompt_attention_mask"]])
    ts = torch.full((2,), 749, dtype=torch.int64)
    flags = np.ones((2, 4, 3), bool)
    ocfg = OracleConfig(sample_size=sample_size, num_layers=4)
    ref = PixArtOracle(sd, ocfg, OracleSchedule.from_flags(flags)).forward(x_in, e_in, ts, None, m_in)
    tr = B200PixArtTransformer2D(sd, cfg, SequentialDiTScheduler(2), PixArtCacheSchedule.default(2, 4))
    out = tr(x_in.cuda(), encoder_hidden_states=e_in.cuda(), encoder_attention_mask=m_in.cuda(), timestep=ts.cuda(),
             added_cond_kwargs={"resolution": None, "aspect_ratio": None}, return_dict=False)[0].cpu()
    assert out.shape == ref.shape == (2, 8, sample_size, sample_size)
    rel = float((out - ref).abs().max() / ref.abs().max())
    assert rel < 5e-3, rel
    assert _cos(out, ref) > 0.99999


def test_single_forward_1024ms_additional_conditions(cuda_device):
    """PixArt-alpha 1024-MS (sample_size 128): resolution / aspect-ratio micro-conditions, interpolation_scale 2,
    N = 4096 image tokens - one forward of a 2-block model against the oracle."""
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.transformer import B200PixArtTransformer2D, SequentialDiTScheduler
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings
    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle

    cfg = PixArtConfig(sample_size=128, num_layers=2)
    assert cfg.resolved_additional_conditions and cfg.resolved_interpolation_scale == 2
    sd = random_init_state_dict(cfg, seed=8)
    emb = synthetic_prompt_embeddings(1, seed=3)
    lat = torch.randn(1, 4, 128, 128, generator=torch.Generator().manual_seed(4))
    x_in = torch.cat([lat, lat])
    e_in = torch.cat([emb["negative_prompt_embeds"], emb["prompt_embeds"]])
    m_in = torch.cat([emb["negative_prompt_attention_mask"], emb["prompt_attention_mask"]])
    ts = torch.full((2,), 399, dtype=torch.int64)
    added = {"resolution": torch.tensor([[1024.0, 768.0], [1024.0, 768.0]]),
             "aspect_ratio": torch.tensor([[1024.0 / 768.0], [1024.0 / 768.0]])}
    flags = np.ones((1, 2, 3), bool)
    ref = PixArtOracle(sd, OracleConfig(sample_size=128, num_layers=2), OracleSchedule.from_flags(flags)) \
        .forward(x_in, e_in, ts, added, m_in)
    tr = B200PixArtTransformer2D(sd, cfg, SequentialDiTScheduler(1), PixArtCacheSchedule.default(1, 2))
    out = tr(x_in.cuda(), encoder_hidden_states=e_in.cuda(), encoder_attention_mask=m_in.cuda(), timestep=ts.cuda(),
             added_cond_kwargs={k: v.cuda() for k, v in added.items()}, return_dict=False)[0].cpu()
    assert out.shape == ref.shape == (2, 8, 128, 128)
    assert float((out - ref).abs().max() / ref.abs().max()) < 5e-3
    assert _cos(out, ref) > 0.99999
    with pytest.raises(ValueError, match="added_cond_kwargs"):
        tr(x_in.cuda(), encoder_hidden_states=e_in.cuda(), encoder_attention_mask=m_in.cuda(), timestep=ts.cuda(),
           added_cond_kwargs=None, return_dict=False)


def test_cached_generation_512px_small_model(cuda_device):
    """A cached multi-step generation at 512x512 (N = 1024 tokens) on a 6-block model: exercises the streamed
    self-attention together with the reuse / fused-residual paths at a second resolution."""
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings
    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle, generate_latents

    L, steps = 6, 5
    cfg = PixArtConfig(sample_size=64, num_layers=L)
    sd = random_init_state_dict(cfg, seed=5)
    rng = np.random.default_rng(3)
    flags = rng.random((steps, L, 3)) < 0.45
    emb = synthetic_prompt_embeddings(1, seed=9)
    traces = []
    gen = B200PixArtAlphaImageGenerator(
        cache_schedule=PixArtCacheSchedule.from_numpy(flags, steps, L, "rand512"), start_seed=0, state_dict=sd,
        model_config=cfg,
        additional_callbacks=[lambda s, t, **kw: traces.append(gen.diffusion_pipeline.transformer.last_executed.copy())])
    got = gen.generate_images(emb, images_per_prompt=1)[0].cpu()
    assert got.shape == (1, 4, 64, 64)
    model = PixArtOracle(sd, OracleConfig(sample_size=64, num_layers=L), OracleSchedule.from_flags(flags))
    noise = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(0))
    ref = generate_latents(model, emb["prompt_embeds"], emb["prompt_attention_mask"], emb["negative_prompt_embeds"],
                           emb["negative_prompt_attention_mask"], noise, steps)["latents"]
    assert np.array_equal(np.stack(traces), model.trace.to_numpy(steps, L))
    assert float((got - ref).abs().max() / ref.abs().max()) <= PER_STEP_REL_MAXABS
    assert _cos(got, ref) >= FINAL_COS


@pytest.mark.parametrize("schedule_file,want_gate", [
    ("alpha_cache_schedules/gen_tgate/tgate_m_010_sp_003_fi_001_warmup_002.json", 10),
    ("alpha_cache_schedules/gen_tgate_m_k_expanded/tgate_m_006_sp_001_fi_001_warmup_002.json", 6),
    ("alpha_cache_schedules/gen_tgate/tgate_m_015_sp_005_fi_001_warmup_002.json", 15)])
def test_tgate_generation_matches_oracle(cuda_device, weights, schedule_file, want_gate):
    """TGATE (ecad/pipelines/tgate.py + compute_attn_tgate): CFG pair until the gate step, the cross-attention cache
    averaged at gate_step - 1, then the null embedding alone with attn2 always served from the averaged cache.  Three
    shipped schedules with different gate steps and self-attention / feed-forward sharing periods."""
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.weights import synthetic_prompt_embeddings

    row = row_by_path(schedule_file)
    flags = flags_of(row)
    gate = row["config"]["pipeline"]["kwargs"]["gate_step"]
    custom = {"name": row["custom"]["attn"], "kwargs": {"gate_step": row["custom"]["gate_step"]}}
    emb = synthetic_prompt_embeddings(2, seed=7)
    traces, per_step = [], []

    def spy(step, timestep, latents=None, **kw):
        traces.append(gen.diffusion_pipeline.transformer.last_executed.copy())
        per_step.append(latents.detach().cpu().clone())

    gen = B200PixArtAlphaImageGenerator(cache_schedule=schedule_of(row), start_seed=0, state_dict=weights,
                                        additional_callbacks=[spy])
    assert gen.gate_step == gate == want_gate
    got = gen.generate_images(emb, images_per_prompt=1)[0].cpu()
    noise = torch.randn(2, 4, 32, 32, generator=torch.Generator().manual_seed(0))
    ref, ref_trace = _oracle_run(weights, flags, emb, noise, custom=custom, gate_step=gate)
    assert np.array_equal(np.stack(traces), ref_trace)
    assert not np.stack(traces)[gate:, :, 1].any()
    for s, (a, b) in enumerate(zip(per_step, ref["per_step"])):
        rel = float((a - b).abs().max() / b.abs().max())
        assert rel <= PER_STEP_REL_MAXABS, (s, rel)
    assert _cos(got, ref["latents"]) >= FINAL_COS
    # a second generation on the same generator starts from the full CFG batch again
    got2 = gen.generate_images(emb, images_per_prompt=1)[0].cpu()
    assert torch.equal(got, got2)


def test_second_generation_reuses_resident_model(cuda_device, weights):
    """Two seeds per prompt + a schedule swap on the resident model give the same result as fresh generators."""
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.weights import synthetic_prompt_embeddings

    emb = synthetic_prompt_embeddings(1, seed=5)
    fast = schedule_of(row_by_path("schedules_in_paper/pixart_alpha_256/ours_fast.json"))
    faster = schedule_of(row_by_path("schedules_in_paper/pixart_alpha_256/ours_faster.json"))
    gen = B200PixArtAlphaImageGenerator(cache_schedule=fast, start_seed=3, seed_step=2, state_dict=weights)
    a = [t.cpu() for t in gen.generate_images(emb, images_per_prompt=2)]
    gen.set_schedule(faster)
    b = gen.generate_images(emb, images_per_prompt=1)[0].cpu()
    gen.set_schedule(fast)
    a2 = gen.generate_images(emb, images_per_prompt=1)[0].cpu()
    assert torch.equal(a[0], a2)  # deterministic, no state leaks across schedules/generations
    assert not torch.equal(a[0], a[1])  # seed stepping
    assert not torch.equal(a[0], b)


@pytest.mark.parametrize("px", [256, 320])
def test_tensor_level_custom_compute_functions(cuda_device, weights, px):
    """(px = 320: 400 image tokens run padded to 512 rows per sample; the functions still see ``[S, 400, D]``.)
    A schedule JSON that names USER-registered compute functions with the reference's tensor signatures
    (cached_transformer_block.py:141-149,161-165) is honoured: the same two Python functions - written against the
    reference's block attribute surface (``block.attn1/attn2/ff``, ``block.cached_*_output``, ``block.cache_schedule``,
    ``block.block_num``) - run unchanged on the B200 path (block proxy over the C ABI) and on the oracle; decisions of
    the untouched blocks stay bit-exact, latents stay inside the bf16 bars."""
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.registry import ComputeAttnRegistry, ComputeFFRegistry
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.weights import synthetic_prompt_embeddings
    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle, generate_latents

    def compute_attn_blend(block, attn, hidden_states, encoder_hidden_states, attention_mask, alpha=0.5, **kw):
        """recompute -> blend the fresh output with the previous one (temporal smoothing); else reuse."""
        cached = getattr(block, f"cached_{attn}_output")
        if block.cache_schedule.get_recompute(block.block_num, attn) or cached is None:
            out = getattr(block, attn)(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                       attention_mask=attention_mask)
            if cached is not None:
                out = alpha * out.float() + (1.0 - alpha) * cached.float()
        else:
            out = cached
        setattr(block, f"cached_{attn}_output", out)
        return out

    def compute_ff_every_other(block, norm_hidden_states, period=2, **kw):
        """ignores the schedule flag: recompute every `period`-th step, otherwise reuse"""
        step = block.cache_schedule.curr_step
        if step % period == 0 or block.cached_ff_output is None:
            block.cached_ff_output = block.ff(norm_hidden_states)
        return block.cached_ff_output

    ComputeAttnRegistry.register_tensor(compute_attn_blend)
    ComputeFFRegistry.register_tensor(compute_ff_every_other)
    PixArtOracle.custom_attn_fns = {"compute_attn_blend": compute_attn_blend}
    PixArtOracle.custom_ff_fns = {"compute_ff_every_other": compute_ff_every_other}
    try:
        steps = 8
        row = row_by_path("schedules_in_paper/pixart_alpha_256/ours_fast.json")
        flags = flags_of(row)[:steps]
        custom_blocks = {3: ("attn",), 4: ("attn", "ff"), 17: ("ff",), 27: ("attn", "ff")}

        def build(cls):
            sched = cls.from_flags(flags) if cls is OracleSchedule else cls.from_numpy(flags, steps, 28, "custom")
            for s in range(steps):
                for b, kinds in custom_blocks.items():
                    e = sched.schedule[s][str(b)]
                    if "attn" in kinds:
                        e["custom_compute_attn"] = {"name": "Compute_Attn_Blend", "kwargs": {"alpha": 0.75}}
                    if "ff" in kinds:
                        e["custom_compute_ff"] = {"name": "compute_ff_every_other", "kwargs": {"period": 3}}
            return sched

        emb = synthetic_prompt_embeddings(2, seed=5)
        traces, per_step = [], []

        def spy(step, timestep, latents=None, **kw):
            traces.append(gen.diffusion_pipeline.transformer.last_executed.copy())
            per_step.append(latents.detach().cpu().clone())

        gen = B200PixArtAlphaImageGenerator(cache_schedule=build(PixArtCacheSchedule), start_seed=0, state_dict=weights,
                                            additional_callbacks=[spy])
        got = gen.generate_images(emb, images_per_prompt=1, height=px, width=px)[0].cpu()

        model = PixArtOracle(weights, OracleConfig(), build(OracleSchedule))
        noise = torch.randn(2, 4, px // 8, px // 8, generator=torch.Generator().manual_seed(0))
        ref = generate_latents(model, emb["prompt_embeds"], emb["prompt_attention_mask"], emb["negative_prompt_embeds"],
                               emb["negative_prompt_attention_mask"], noise, steps, record_steps=True)
        # executed = "the block's own module ran": identical on both sides, custom blocks included
        assert np.array_equal(np.stack(traces), model.trace.to_numpy(steps, 28))
        ran = np.stack(traces)
        assert ran[0, 4, 2] == 1 and ran[1, 4, 2] == 0 and ran[3, 4, 2] == 1  # the period-3 rule, not the schedule flag
        for s, (a, b) in enumerate(zip(per_step, ref["per_step"])):
            rel = float((a - b).abs().max() / b.abs().max())
            assert rel <= PER_STEP_REL_MAXABS, (s, rel)
        assert _cos(got, ref["latents"]) >= FINAL_COS
        # and the functions really changed the result (vs the same flags with default functions)
        base, _ = _oracle_run(weights, flags, emb, noise)
        assert float((ref["latents"] - base["latents"]).abs().max() / base["latents"].abs().max()) > 1e-4
    finally:
        ComputeAttnRegistry._tensor_registry.pop("compute_attn_blend", None)
        ComputeFFRegistry._tensor_registry.pop("compute_ff_every_other", None)
        PixArtOracle.custom_attn_fns, PixArtOracle.custom_ff_fns = {}, {}


def test_sigma_cached_generation_shipped_schedule(cuda_device, weights):
    """PixArt-sigma 256x256 (BASELINE config 4's model family: 300 text tokens -> cross-attention keys padded to 384,
    streamed), a full 20-step generation under the paper's shipped schedule
    schedules/schedules_in_paper/pixart_sigma_256/ours_fast.json, through the sigma generator entry."""
    from ecad_b200.image_generator import B200PixArtSigmaImageGenerator
    from ecad_b200.weights import synthetic_prompt_embeddings

    row = row_by_path("schedules_in_paper/pixart_sigma_256/ours_fast.json")
    flags = flags_of(row)
    emb = synthetic_prompt_embeddings(1, text_tokens=B200PixArtSigmaImageGenerator.text_tokens, seed=11)
    traces, per_step = [], []

    def spy(step, timestep, latents=None, **kw):
        traces.append(gen.diffusion_pipeline.transformer.last_executed.copy())
        per_step.append(latents.detach().cpu().clone())

    gen = B200PixArtSigmaImageGenerator(cache_schedule=schedule_of(row), start_seed=0, state_dict=weights,
                                        additional_callbacks=[spy])
    assert gen.text_tokens == 300 and not gen.model_config.resolved_additional_conditions
    got = gen.generate_images(emb, images_per_prompt=1)[0].cpu()
    noise = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(0))
    ref, ref_trace = _oracle_run(weights, flags, emb, noise)
    assert np.array_equal(np.stack(traces), ref_trace)
    assert 0.2 < ref_trace[1:].mean() < 0.8  # a genuinely cached run
    for s, (a, b) in enumerate(zip(per_step, ref["per_step"])):
        rel = float((a - b).abs().max() / b.abs().max())
        assert rel <= PER_STEP_REL_MAXABS, (s, rel)
    assert _cos(got, ref["latents"]) >= FINAL_COS


def test_sigma_1024px_cached_multistep_reduced_depth(cuda_device):
    """BASELINE config 4 shape (PixArt-sigma 1024x1024: N = 4096 image tokens, 300 text tokens, no micro-conditions) on
    a reduced-depth model (3 blocks) so the CPU oracle finishes: a 4-step CACHED run - dense first step, then reuse /
    recompute mixes incl. a step that reuses everything - with the 4096-key streamed self-attention, the 384-key
    streamed cross-attention and the lazy reuse path all at the full token count."""
    from ecad_b200.image_generator import B200PixArtSigmaImageGenerator
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings
    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle, generate_latents

    L, steps = 3, 4
    cfg = PixArtConfig(sample_size=128, num_layers=L, use_additional_conditions=False)
    sd = random_init_state_dict(cfg, seed=6)
    flags = np.ones((steps, L, 3), bool)
    flags[1] = [[False, True, False], [True, False, True], [False, False, True]]
    flags[2] = False
    flags[3] = [[True, False, False], [False, True, False], [True, True, True]]
    emb = synthetic_prompt_embeddings(1, text_tokens=300, seed=12)
    traces, per_step = [], []

    def spy(step, timestep, latents=None, **kw):
        traces.append(gen.diffusion_pipeline.transformer.last_executed.copy())
        per_step.append(latents.detach().cpu().clone())

    gen = B200PixArtSigmaImageGenerator(cache_schedule=PixArtCacheSchedule.from_numpy(flags, steps, L, "sigma1024"),
                                        start_seed=0, state_dict=sd, model_config=cfg, additional_callbacks=[spy])
    got = gen.generate_images(emb, images_per_prompt=1)[0].cpu()
    assert got.shape == (1, 4, 128, 128)
    ocfg = OracleConfig(sample_size=128, num_layers=L, use_additional_conditions=False)
    model = PixArtOracle(sd, ocfg, OracleSchedule.from_flags(flags))
    noise = torch.randn(1, 4, 128, 128, generator=torch.Generator().manual_seed(0))
    ref = generate_latents(model, emb["prompt_embeds"], emb["prompt_attention_mask"], emb["negative_prompt_embeds"],
                           emb["negative_prompt_attention_mask"], noise, steps, record_steps=True)
    assert np.array_equal(np.stack(traces), model.trace.to_numpy(steps, L))
    assert np.array_equal(np.stack(traces)[1:], flags[1:].astype(np.uint8))  # caches exist after step 0
    for s, (a, b) in enumerate(zip(per_step, ref["per_step"])):
        rel = float((a - b).abs().max() / b.abs().max())
        assert rel <= PER_STEP_REL_MAXABS, (s, rel)
    assert _cos(got, ref["latents"]) >= FINAL_COS


@pytest.mark.parametrize("hl,wl", [(24, 24), (40, 40), (24, 40), (48, 48)])
def test_token_counts_that_are_not_a_multiple_of_256(cuda_device, hl, wl):
    """192 / 320 / 384 px and a non-square size: 144 / 400 / 240 / 576 image tokens run PADDED to 256 / 512 / 256 /
    768 rows per sample (zero padding rows, masked as self-attention keys, dropped by the unpatchify epilogue) and
    must give the oracle's result on the real tokens - dense step, then a step that reuses sub-blocks."""
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.transformer import B200PixArtTransformer2D, SequentialDiTScheduler
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings
    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle

    cfg = PixArtConfig(num_layers=3)
    sd = random_init_state_dict(cfg, seed=2)
    emb = synthetic_prompt_embeddings(2, seed=3)
    lat = torch.randn(2, 4, hl, wl, generator=torch.Generator().manual_seed(4))
    x_in = torch.cat([lat, lat])
    e_in = torch.cat([emb["negative_prompt_embeds"], emb["prompt_embeds"]])
    m_in = torch.cat([emb["negative_prompt_attention_mask"], emb["prompt_attention_mask"]])
    flags = np.ones((2, 3, 3), bool)
    flags[1, 0, 0] = flags[1, 1, 1] = flags[1, 2, 2] = flags[1, 2, 0] = False
    oracle = PixArtOracle(sd, OracleConfig(num_layers=3), OracleSchedule.from_flags(flags))
    sched = PixArtCacheSchedule.from_numpy(flags, 2, 3)
    tr = B200PixArtTransformer2D(sd, cfg, SequentialDiTScheduler(2), sched)
    for step, t in enumerate((949, 899)):
        ts = torch.full((4,), t, dtype=torch.int64)
        ref = oracle.forward(x_in, e_in, ts, None, m_in)
        out = tr(x_in.cuda(), encoder_hidden_states=e_in.cuda(), encoder_attention_mask=m_in.cuda(),
                 timestep=ts.cuda(), added_cond_kwargs={"resolution": None, "aspect_ratio": None},
                 return_dict=False)[0].cpu()
        assert out.shape == ref.shape == (4, 8, hl, wl)
        assert np.array_equal(tr.last_executed.astype(bool), flags[step] if step else np.ones((3, 3), bool))
        assert torch.isfinite(out).all()
        rel = float((out - ref).abs().max() / ref.abs().max())
        assert rel < 5e-3, (step, rel)
        assert _cos(out, ref) > 0.99999
        sched.per_step_callback(step)
        oracle.cache_schedule.per_step_callback(step)


def test_generation_at_320px_matches_oracle(cuda_device):
    """A cached 6-step generation at 320 x 320 (400 tokens -> 512 padded rows) through the generator API
    (``generate_images(height=, width=)``): decisions bit-exact, per-step latents within the bar."""
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings

    L, steps = 4, 6
    cfg = PixArtConfig(num_layers=L)
    sd = random_init_state_dict(cfg, seed=5)
    rng = np.random.default_rng(3)
    flags = rng.random((steps, L, 3)) < 0.5
    flags[0] = True
    emb = synthetic_prompt_embeddings(2, seed=1)
    traces, per_step = [], []

    def spy(step, timestep, latents=None, **kw):
        traces.append(gen.diffusion_pipeline.transformer.last_executed.copy())
        per_step.append(latents.detach().cpu().clone())

    gen = B200PixArtAlphaImageGenerator(cache_schedule=PixArtCacheSchedule.from_numpy(flags, steps, L, "px320"),
                                        start_seed=0, state_dict=sd, model_config=cfg, additional_callbacks=[spy])
    got = gen.generate_images(emb, height=320, width=320)[0].cpu()
    assert got.shape == (2, 4, 40, 40)

    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle, generate_latents
    model = PixArtOracle(sd, OracleConfig(num_layers=L), OracleSchedule.from_flags(flags))
    noise = torch.randn(2, 4, 40, 40, generator=torch.Generator().manual_seed(0))
    ref = generate_latents(model, emb["prompt_embeds"], emb["prompt_attention_mask"], emb["negative_prompt_embeds"],
                           emb["negative_prompt_attention_mask"], noise, steps, record_steps=True)
    assert np.array_equal(np.stack(traces), model.trace.to_numpy(steps, L))
    for s, (a, b) in enumerate(zip(per_step, ref["per_step"])):
        rel = float((a - b).abs().max() / b.abs().max())
        assert rel <= PER_STEP_REL_MAXABS, (s, rel)
    assert _cos(got, ref["latents"]) >= FINAL_COS
