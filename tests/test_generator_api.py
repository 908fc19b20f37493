"""Host-side pieces of the kept ImageGenerator API (no GPU): checkpoint config parsing, the saved-prompt loader with
the DataLoader contract, the file naming of generate_from_saved_prompts / time_image_generation driven through a stub
generator (the GPU run of the same methods is tests/test_gpu_graphs.py::test_generate_from_saved_prompts)."""
import json

import pytest
import torch

from ecad_b200.dataset import PIXART_KEYS, PromptEmbeddingDataset
from ecad_b200.image_generator import _SavedPromptMixin
from ecad_b200.weights import PixArtConfig, pixart_config_from_pretrained, synthetic_prompt_embeddings


def _write_prompts(root, n=5, tokens=16, channels=8):
    emb = synthetic_prompt_embeddings(n, text_tokens=tokens, channels=channels, seed=2)
    for i in range(n):
        d = root / ("a" if i < 3 else "b/c")
        d.mkdir(parents=True, exist_ok=True)
        torch.save({k: emb[k][i:i + 1] for k in PIXART_KEYS}, d / f"p{i}.pt")
    return emb


def test_pixart_config_from_checkpoint_dir(tmp_path):
    d = tmp_path / "pipe" / "transformer"
    d.mkdir(parents=True)
    torch.save({}, d / "diffusion_pytorch_model.bin")
    raw = {"_class_name": "PixArtTransformer2DModel", "sample_size": 128, "num_layers": 28, "attention_head_dim": 72,
           "num_attention_heads": 16, "in_channels": 4, "out_channels": 8, "cross_attention_dim": 1152,
           "caption_channels": 4096, "norm_type": "ada_norm_single", "interpolation_scale": 2,
           "use_additional_conditions": True, "activation_fn": "gelu-approximate", "norm_eps": 1e-6, "patch_size": 2}
    (d / "config.json").write_text(json.dumps(raw))
    cfg = pixart_config_from_pretrained(tmp_path / "pipe")
    assert cfg.sample_size == 128 and cfg.resolved_interpolation_scale == 2 and cfg.resolved_additional_conditions
    raw["use_additional_conditions"] = None
    raw["sample_size"] = 64
    raw["interpolation_scale"] = None
    (d / "config.json").write_text(json.dumps(raw))
    cfg = pixart_config_from_pretrained(d)
    assert cfg.sample_size == 64 and cfg.resolved_interpolation_scale == 1 and not cfg.resolved_additional_conditions
    raw["norm_type"] = "layer_norm"
    (d / "config.json").write_text(json.dumps(raw))
    with pytest.raises(ValueError):
        pixart_config_from_pretrained(d)
    with pytest.raises(FileNotFoundError):
        pixart_config_from_pretrained(tmp_path / "nothing")


def test_batches_follow_the_dataloader_contract(tmp_path):
    emb = _write_prompts(tmp_path)
    ds = PromptEmbeddingDataset(tmp_path)
    loader = ds.batches(2)
    assert len(loader) == 3
    first = list(loader)
    again = list(loader)  # re-iterable like a DataLoader
    assert [b["name"] for b in first] == [b["name"] for b in again] == [["p0", "p1"], ["p2", "p3"], ["p4"]]
    assert torch.equal(torch.cat([b["prompt_embeds"] for b in first]), emb["prompt_embeds"])
    shuffled = [n for b in ds.batches(2, shuffle=True, seed=3) for n in b["name"]]
    assert sorted(shuffled) == [f"p{i}" for i in range(5)]
    # FLUX-style dicts (other tensor keys) batch the same way
    fl = tmp_path / "flux"
    fl.mkdir()
    for i in range(3):
        torch.save({"prompt_embeds": torch.full((1, 4, 6), float(i)), "pooled_prompt_embeds": torch.full((1, 5), float(i)),
                    "text_ids": None}, fl / f"f{i}.pt")
    b = next(iter(PromptEmbeddingDataset(fl).batches(3)))
    assert b["prompt_embeds"].shape == (3, 4, 6) and b["pooled_prompt_embeds"].shape == (3, 5) and "text_ids" not in b


class _StubGenerator(_SavedPromptMixin):
    """Stands in for the GPU generator: 'images' encode (prompt value, seed index)."""

    def __init__(self):
        self.diffusion_pipeline, self.start_seed, self.seed_step, self.device = None, 7, 5, "cpu"
        self.created = 0
        self.timed = []

    def create_diffusion_pipeline(self):
        self.created += 1
        self.diffusion_pipeline = object()
        return self.diffusion_pipeline

    def generate_images(self, embeds, images_per_prompt=1, **kw):
        base = embeds["prompt_embeds"][:, 0, 0]
        return [base + 100.0 * i for i in range(images_per_prompt)]

    def generate_images_timed(self, embeds, **kw):
        self.timed.append(list(embeds["name"]))
        return float(len(self.timed))


def test_generate_from_saved_prompts_file_contract(tmp_path):
    src, dst = tmp_path / "in", tmp_path / "out"
    emb = _write_prompts(src)
    g = _StubGenerator()
    g.generate_from_saved_prompts(src, dst, batch_size=2, images_per_prompt=2, free_after=True)
    assert g.created == 1 and g.diffusion_pipeline is None  # created on demand, freed after
    files = sorted(p.relative_to(dst).as_posix() for p in dst.glob("**/*.pt"))
    # image_generator.py:400-409: <rel_path>/<name>__image_seed:<start + i*step, 3 digits>
    assert files == sorted(f"{'a' if i < 3 else 'b/c'}/p{i}__image_seed:{s:03}.pt" for i in range(5) for s in (7, 12))
    got = torch.load(dst / "b/c" / "p4__image_seed:012.pt")
    assert float(got) == pytest.approx(float(emb["prompt_embeds"][4, 0, 0]) + 100.0)
    g.generate_from_saved_prompts(src, tmp_path / "plain", batch_size=5, include_seed_in_name=False)
    assert (tmp_path / "plain" / "a" / "p0.pt").exists()


def test_time_image_generation_cycles_the_directory(tmp_path):
    src = tmp_path / "in"
    _write_prompts(src, n=3)
    g = _StubGenerator()
    times = g.time_image_generation(src, batch_size=2, num_batches=5)
    assert times == [1.0, 2.0, 3.0, 4.0, 5.0]
    assert g.timed == [["p0", "p1"], ["p2"], ["p0", "p1"], ["p2"], ["p0", "p1"]]
    with pytest.raises(ValueError):
        g.time_image_generation(tmp_path / "empty_dir_that_does_not_exist", num_batches=1)


# ---- the third plug-in point: PipelineRegistry (ecad/pipelines/load_pipeline.py:16-58) --------------------------
def test_pipeline_registry_names_and_fallback():
    from ecad_b200.flux_pipeline import B200FluxPipeline
    from ecad_b200.pipeline import B200PixArtPipeline, B200TGATEPipeline
    from ecad_b200.registry import PipelineRegistry, pipeline_from_pretrained

    # the reference's five names
    assert PipelineRegistry.get("pixart_alpha") is B200PixArtPipeline
    assert PipelineRegistry.get("pixart_sigma") is B200PixArtPipeline
    assert PipelineRegistry.get("pass_through") is B200PixArtPipeline
    assert PipelineRegistry.get("tgate") is B200TGATEPipeline
    assert PipelineRegistry.get("flux") is B200FluxPipeline
    # load_pipeline.py:25-41: unknown -> default -> None
    assert PipelineRegistry.get("nope") is None
    assert PipelineRegistry.get("nope", "flux") is B200FluxPipeline
    assert PipelineRegistry.get("nope", "neither") is None

    # load_pipeline.py:44-58: the default applies when the config has no name; kwargs are forwarded
    made = []

    @PipelineRegistry.register("recording")
    class Recording:
        @classmethod
        def from_pretrained(cls, *args, **kwargs):
            made.append((args, kwargs))
            return cls()

    try:
        f = pipeline_from_pretrained({"name": "recording", "kwargs": {"gate_step": 7}}, "pixart_alpha")
        assert isinstance(f("transformer", use_cuda_graph=True), Recording)
        assert made == [(("transformer",), {"use_cuda_graph": True, "gate_step": 7})]
        assert pipeline_from_pretrained({}, "recording").pipeline_class is Recording
        assert pipeline_from_pretrained(None, "pixart_sigma").pipeline_class is B200PixArtPipeline
        with pytest.raises(ValueError):
            pipeline_from_pretrained({"name": "nope"}, "pixart_alpha")  # a NAMED unknown pipeline does not fall back
    finally:
        PipelineRegistry._registry.pop("recording", None)


def test_tgate_pipeline_needs_its_gate_step():
    from ecad_b200.pipeline import B200TGATEPipeline

    class Tr:
        device = "cpu"

    with pytest.raises(ValueError, match="gate_step must be provided"):  # tgate.py:51-55
        B200TGATEPipeline.from_pretrained(Tr())
    assert B200TGATEPipeline.from_pretrained(Tr(), gate_step=8).gate_step == 8


def test_generator_reads_the_schedule_config(monkeypatch):
    """image_generator.py:172-191 + pixart_image_generator.py:78-82: transformer_weights, pipeline, height / width."""
    import numpy as np

    from ecad_b200 import image_generator as ig
    from ecad_b200.pipeline import B200PixArtPipeline, B200TGATEPipeline
    from ecad_b200.schedule import PixArtCacheSchedule

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    built = []
    monkeypatch.setattr(ig, "random_init_state_dict", lambda cfg, seed: built.append(cfg) or {})

    def sched(config):
        return PixArtCacheSchedule.from_numpy(np.ones((2, 28, 3), bool), 2, 28, top_level_config=config)

    g = ig.B200PixArtAlphaImageGenerator(cache_schedule=sched({}))
    assert (g.height, g.width) == (256, 256) and g.model_config.sample_size == 32 and g.gate_step is None
    assert g.pipeline_from_pretrained.pipeline_class is B200PixArtPipeline

    # the shipped gen_default_1024x1024 config block
    g = ig.B200PixArtAlphaImageGenerator(cache_schedule=sched(
        {"transformer_weights": "PixArt-alpha/PixArt-XL-2-1024-MS", "height": 1024, "width": 1024}))
    assert (g.height, g.width) == (1024, 1024)
    assert g.model_config.sample_size == 128 and g.model_config.resolved_additional_conditions
    assert built[-1] is g.model_config  # the random-init weights are drawn for THAT architecture

    g = ig.B200PixArtSigmaImageGenerator(cache_schedule=sched(
        {"transformer_weights": "PixArt-alpha/PixArt-Sigma-XL-2-1024-MS"}))
    assert g.model_config.sample_size == 128 and not g.model_config.resolved_additional_conditions
    assert (g.height, g.width) == (1024, 1024) and g.text_tokens == 300

    g = ig.B200PixArtAlphaImageGenerator(cache_schedule=sched({"pipeline": {"name": "tgate", "kwargs": {"gate_step": 1}}}))
    assert g.gate_step == 1 and g.pipeline_from_pretrained.pipeline_class is B200TGATEPipeline

    with pytest.raises(ValueError, match="transformer_weights"):
        ig.B200PixArtAlphaImageGenerator(cache_schedule=sched({"transformer_weights": "someone/else"}))
    # an explicit architecture wins over the name in the JSON
    small = PixArtConfig(num_layers=28, sample_size=64)
    g = ig.B200PixArtAlphaImageGenerator(cache_schedule=sched({"transformer_weights": "someone/else"}), model_config=small)
    assert g.model_config is small and (g.height, g.width) == (512, 512)
    # ... and then the JSON's height / width (they describe the checkpoint IT names) do not apply either: the seed
    # population `pixart_alpha_256x256/gen_000` carries the 1024-MS block and is evaluated on the 256 px model
    g = ig.B200PixArtAlphaImageGenerator(
        cache_schedule=sched({"transformer_weights": "PixArt-alpha/PixArt-XL-2-1024-MS", "height": 1024, "width": 1024}),
        state_dict={})
    assert (g.height, g.width) == (256, 256) and g.model_config.sample_size == 32
