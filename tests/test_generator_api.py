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
