"""Whole-generation CUDA graphs (ecad_b200/graphs.py): a replayed generation must equal the eager one bit for bit -
same kernels, same decisions, same inputs - for new prompts and seeds fed through the static buffers."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pixart_graph_replay_equals_eager(cuda_device):
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.schedule import schedule_from_packed
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings
    from golden_util import row_by_path

    row = row_by_path("schedules_in_paper/pixart_alpha_256/ours_fast.json")
    sd = random_init_state_dict(PixArtConfig(), 0)
    seen = []
    eager = B200PixArtAlphaImageGenerator(cache_schedule=schedule_from_packed(row), state_dict=sd)
    graphed = B200PixArtAlphaImageGenerator(cache_schedule=schedule_from_packed(row), state_dict=sd, use_cuda_graph=True,
                                            additional_callbacks=[lambda s, t, **kw: seen.append(s)])
    per_gen = []  # kernels per eager generation (the first one also fills the per-timestep adaLN-table cache)
    for it, (batch, seed) in enumerate([(2, 1), (2, 5), (2, 9)]):
        emb = synthetic_prompt_embeddings(batch, seed=seed)
        eager.start_seed = graphed.start_seed = 10 * it
        l0 = eager.diffusion_pipeline.transformer.launches if eager.diffusion_pipeline else 0
        a = eager.generate_images(emb)[0]
        per_gen.append(eager.diffusion_pipeline.transformer.launches - l0)
        b = graphed.generate_images(emb)[0]
        assert torch.equal(a, b), float((a - b).abs().max())
    g = graphed.diffusion_pipeline._graphs
    assert g.captures == 1 and g.replays == 3
    assert seen == list(range(20)) * 3  # the per-step callback protocol still runs once per step
    tr_e, tr_g = eager.diffusion_pipeline.transformer, graphed.diffusion_pipeline.transformer
    # launch accounting: one eager warm-up generation + 3 replays of the recorded (warm-cache) count
    assert per_gen[1] == per_gen[2] < per_gen[0]
    assert tr_g.launches == per_gen[0] + 3 * per_gen[1] and tr_e.launches == sum(per_gen)
    assert graphed.cache_schedule.curr_step == 0 and not tr_g._has_cache.any()
    # swapping the candidate schedule records a new graph
    flags = np.ones((20, 28, 3), bool)
    from ecad_b200.schedule import PixArtCacheSchedule
    dense = PixArtCacheSchedule.from_numpy(flags, 20, 28, "dense")
    eager.set_schedule(dense); graphed.set_schedule(dense)
    emb = synthetic_prompt_embeddings(2, seed=3)
    assert torch.equal(eager.generate_images(emb)[0], graphed.generate_images(emb)[0])
    assert g.captures == 2


def test_flux_graph_replay_equals_eager(cuda_device):
    from ecad_b200.image_generator import B200FluxImageGenerator
    from ecad_b200.schedule import FluxCacheSchedule
    from ecad_b200.weights import FluxConfig, flux_random_init_state_dict
    from test_gpu_flux_parity import SMALL, _embeds, _schedule_flags

    cfg = FluxConfig(**SMALL)
    steps, rows = 6, cfg.num_layers + cfg.num_single_layers
    sd = flux_random_init_state_dict(cfg, seed=0)

    def make(graph):
        sched = FluxCacheSchedule.from_numpy(_schedule_flags(steps, rows), steps, cfg.num_layers, cfg.num_single_layers,
                                             "rand", top_level_config={"height": 256, "width": 192})
        return B200FluxImageGenerator(cache_schedule=sched, state_dict=sd, model_config=cfg, use_cuda_graph=graph)

    eager, graphed = make(False), make(True)
    for it in range(2):
        emb = _embeds(2, 64, SMALL, seed=4 + it)
        eager.start_seed = graphed.start_seed = it
        a, b = eager.generate_images(emb)[0], graphed.generate_images(emb)[0]
        assert torch.equal(a, b), float((a - b).abs().max())
    assert graphed.diffusion_pipeline._graphs.captures == 1


def test_dead_cache_store_elimination_is_invisible(cuda_device):
    """Skipping the cache store of an executed sub-block whose slot is overwritten (or dropped) before any read must not
    change a single bit of the result - for the paper's schedule, a TGATE schedule and a FLUX schedule."""
    from ecad_b200.image_generator import B200FluxImageGenerator, B200PixArtAlphaImageGenerator
    from ecad_b200.schedule import FluxCacheSchedule, schedule_from_packed
    from ecad_b200.weights import (FluxConfig, PixArtConfig, flux_random_init_state_dict, random_init_state_dict,
                                   synthetic_prompt_embeddings)
    from golden_util import rows, row_by_path
    from test_gpu_flux_parity import SMALL, _embeds, _schedule_flags

    sd = random_init_state_dict(PixArtConfig(), 0)
    emb = synthetic_prompt_embeddings(2, seed=2)
    tgate = [r for r in rows() if (r.get("config") or {}).get("pipeline", {}).get("name") == "tgate" and r["NB"] == 28
             and r["tokens"] == 256][0]
    for row in (row_by_path("schedules_in_paper/pixart_alpha_256/ours_fast.json"), tgate):
        outs, dead_total = [], []
        for skip in (True, False):
            seen = []
            gen = B200PixArtAlphaImageGenerator(
                cache_schedule=schedule_from_packed(row), state_dict=sd,
                additional_callbacks=[lambda s, t, **kw: seen.append(int(gen.diffusion_pipeline.transformer.last_dead.sum()))])
            gen.create_diffusion_pipeline().transformer.skip_dead_cache_stores = skip
            outs.append(gen.generate_images(emb)[0])
            dead_total.append(sum(seen))
        assert torch.equal(outs[0], outs[1]), row["path"]
        assert dead_total[0] > 0 and dead_total[1] == 0, (row["path"], dead_total)

    cfg = FluxConfig(**SMALL)
    steps, nrows = 6, cfg.num_layers + cfg.num_single_layers
    fsd = flux_random_init_state_dict(cfg, seed=0)
    femb = _embeds(2, 64, SMALL, seed=4)
    outs = []
    for skip in (True, False):
        sched = FluxCacheSchedule.from_numpy(_schedule_flags(steps, nrows), steps, cfg.num_layers, cfg.num_single_layers,
                                             "rand", top_level_config={"height": 256, "width": 192})
        gen = B200FluxImageGenerator(cache_schedule=sched, state_dict=fsd, model_config=cfg)
        gen.create_diffusion_pipeline().transformer.skip_dead_cache_stores = skip
        outs.append(gen.generate_images(femb)[0])
    assert torch.equal(outs[0], outs[1])


def test_latency_metrics_block(cuda_device, tmp_path):
    """metrics.latency in the reference's layout (compute_latency.py:52-73), measured on the resident model."""
    import json

    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.metrics import annotate_schedule_file
    from ecad_b200.schedule import schedule_from_packed
    from ecad_b200.weights import synthetic_prompt_embeddings
    from golden_util import row_by_path

    row = row_by_path("schedules_in_paper/pixart_alpha_256/ours_fast.json")
    sched = schedule_from_packed(row)
    f = tmp_path / "ours_fast.json"
    f.write_text(json.dumps(sched.to_dict()))
    gen = B200PixArtAlphaImageGenerator(cache_schedule=sched)
    emb = {k: v.cuda() for k, v in synthetic_prompt_embeddings(4).items()}
    m = annotate_schedule_file(f, image_generator=gen, prompt_embeds=emb, num_samples=2, warmup_steps=1)
    lat = m["latency"]
    assert set(lat) == {"avg", "batch_size", "num_samples", "warmup_steps", "gpu", "warmups", "latencies", "output_type"}
    assert lat["output_type"] == "latent"  # the reference's figure includes the VAE decode: the block says what was timed
    assert lat["batch_size"] == 4 and len(lat["latencies"]) == 2 and len(lat["warmups"]) == 1 and 0 < lat["avg"] < 1000
    assert m["total_macs"] == row["total_macs"]
    assert json.loads(f.read_text())["metrics"]["latency"]["gpu"] == lat["gpu"]


def test_population_run_from_host_pipelines_copies_and_keeps_results(cuda_device):
    """PopulationEvaluator.run_from_host: inputs from pinned host memory, results to pinned host memory, copies on a
    side stream - results must equal the plain per-unit path."""
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.population import PopulationEvaluator
    from ecad_b200.schedule import schedule_from_packed
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings
    from golden_util import row_by_path

    rows = [row_by_path(f"population_initialization/pixart_alpha_256x256/gen_000/candidates/cand_{i:03d}.json")
            for i in (1, 2, 5)]
    sd = random_init_state_dict(PixArtConfig(), 0)
    gen = B200PixArtAlphaImageGenerator(cache_schedule=schedule_from_packed(rows[0]), state_dict=sd)
    emb_host = {k: v.pin_memory() for k, v in synthetic_prompt_embeddings(3, seed=5).items()}

    def run_unit(i, emb):
        gen.set_schedule(schedule_from_packed(rows[i]))
        return gen.generate_images(emb)[0]

    ev = PopulationEvaluator(0, 1, cuda_device)
    out = ev.run_from_host(range(3), run_unit, emb_host)
    ref = [run_unit(i, {k: v.cuda() for k, v in emb_host.items()}).cpu() for i in range(3)]
    assert len(out["host"]) == 3 and all(t.is_pinned() for t in out["host"])
    for got, dev, want in zip(out["host"], out["device"], ref):
        assert torch.equal(got, want) and torch.equal(dev.cpu(), want)
    assert not torch.equal(ref[0], ref[1])  # different candidates really produce different latents


def test_graph_cache_is_keyed_on_schedule_content(cuda_device):
    """A set_schedule() loop over GA candidates creates many short-lived schedule objects that all carry the same name
    ("from_numpy") and whose addresses get reused: the graph cache must key on what the schedule SAYS.  Same content
    (another object) -> the recorded graph is replayed; different content under the same name -> a new recording, and
    the replayed latents equal the eager ones in both cases."""
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings

    L, steps = 4, 4
    cfg = PixArtConfig(num_layers=L)
    sd = random_init_state_dict(cfg, 0)
    rng = np.random.default_rng(0)
    fa = rng.random((steps, L, 3)) < 0.5
    fb = fa.copy()
    fb[2, 1, 2] = not fb[2, 1, 2]
    eager = B200PixArtAlphaImageGenerator(cache_schedule=PixArtCacheSchedule.from_numpy(fa, steps, L), state_dict=sd,
                                          model_config=cfg)
    graphed = B200PixArtAlphaImageGenerator(cache_schedule=PixArtCacheSchedule.from_numpy(fa, steps, L), state_dict=sd,
                                            model_config=cfg, use_cuda_graph=True)
    emb = synthetic_prompt_embeddings(2, seed=4)
    outs = {}
    for tag, f in (("a", fa), ("b", fb), ("a again", fa), ("b again", fb)):
        # a NEW object every time, all named "from_numpy"
        eager.set_schedule(PixArtCacheSchedule.from_numpy(f, steps, L))
        graphed.set_schedule(PixArtCacheSchedule.from_numpy(f, steps, L))
        e, g = eager.generate_images(emb)[0], graphed.generate_images(emb)[0]
        assert torch.equal(e, g), tag
        outs[tag] = g
    assert not torch.equal(outs["a"], outs["b"])
    assert torch.equal(outs["a"], outs["a again"]) and torch.equal(outs["b"], outs["b again"])
    gg = graphed.diffusion_pipeline._graphs
    assert gg.captures == 2 and gg.replays == 4
    # a larger batch re-allocates the transformer workspace: the recorded graphs (raw pointers into it) are dropped
    emb3 = synthetic_prompt_embeddings(3, seed=4)
    eager.set_schedule(PixArtCacheSchedule.from_numpy(fa, steps, L))
    graphed.set_schedule(PixArtCacheSchedule.from_numpy(fa, steps, L))
    assert torch.equal(eager.generate_images(emb3)[0], graphed.generate_images(emb3)[0])
    assert torch.equal(eager.generate_images(emb)[0], graphed.generate_images(emb)[0])  # back to the small batch


def test_generate_from_saved_prompts_and_module_shell(cuda_device, tmp_path):
    """The file-driven generator API on the GPU (image_generator.py:366-421,442-487): a directory of saved prompt
    embeddings -> one latent file per (prompt, seed) equal to a direct generate_images call; time_image_generation
    returns one ms-per-image figure per batch.  Also: the transformer is an nn.Module shell a pipeline can hold."""
    from ecad_b200.dataset import PIXART_KEYS
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings

    L, steps = 3, 3
    cfg = PixArtConfig(num_layers=L)
    sd = random_init_state_dict(cfg, 1)
    flags = np.random.default_rng(1).random((steps, L, 3)) < 0.6
    emb = synthetic_prompt_embeddings(3, seed=8)
    src, dst = tmp_path / "prompts", tmp_path / "latents"
    for i in range(3):
        d = src / ("x" if i < 2 else "y")
        d.mkdir(parents=True, exist_ok=True)
        torch.save({k: emb[k][i:i + 1] for k in PIXART_KEYS}, d / f"p{i}.pt")
    gen = B200PixArtAlphaImageGenerator(cache_schedule=PixArtCacheSchedule.from_numpy(flags, steps, L), start_seed=5,
                                        seed_step=3, state_dict=sd, model_config=cfg)
    gen.generate_from_saved_prompts(src, dst, batch_size=2, images_per_prompt=2)
    files = sorted(p.relative_to(dst).as_posix() for p in dst.glob("**/*.pt"))
    assert files == sorted(f"{'x' if i < 2 else 'y'}/p{i}__image_seed:{s:03}.pt" for i in range(3) for s in (5, 8))
    direct = gen.generate_images({k: emb[k][:2] for k in PIXART_KEYS}, images_per_prompt=2)
    assert torch.equal(torch.load(dst / "x" / "p1__image_seed:008.pt"), direct[1][1].cpu())
    times = gen.time_image_generation(src, batch_size=2, num_batches=3)
    assert len(times) == 3 and all(t > 0 for t in times)

    tr = gen.diffusion_pipeline.transformer
    assert isinstance(tr, torch.nn.Module) and not tr.training and list(tr.parameters()) == []
    names = dict(tr.named_buffers())
    assert "block0_w_qkv1" in names and names["block0_w_qkv1"].dtype == torch.bfloat16 and len(tr.state_dict()) > 15 * L
    assert tr.to("cuda:0") is tr and tr.eval() is tr
    with pytest.raises(RuntimeError, match="bound to its device"):
        tr.to("cpu")
    with pytest.raises(RuntimeError, match="bound to its device"):
        tr.half()
    holder = torch.nn.ModuleDict({"transformer": tr})  # what a diffusers pipeline's register_modules needs
    assert holder["transformer"] is tr
    # the forward returns fresh tensors: two outputs kept across calls do not alias
    x = torch.randn(2, 4, 32, 32, device="cuda")
    kw = dict(encoder_hidden_states=emb["prompt_embeds"][:2].cuda(), encoder_attention_mask=emb["prompt_attention_mask"][:2].cuda(),
              timestep=torch.full((2,), 500, device="cuda"), added_cond_kwargs={"resolution": None, "aspect_ratio": None},
              return_dict=False)
    o1 = tr(x, **kw)[0]
    keep = o1.clone()
    tr.reset_cache()  # the step counter did not advance: start the second forward from empty caches too
    o2 = tr(x * 0.5, **kw)[0]
    assert o1.data_ptr() != o2.data_ptr() and torch.equal(o1, keep)
