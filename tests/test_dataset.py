import torch

from ecad_b200.dataset import PIXART_KEYS, PromptEmbeddingDataset
from ecad_b200.weights import synthetic_prompt_embeddings


def test_prompt_embedding_dataset_round_trip(tmp_path):
    emb = synthetic_prompt_embeddings(5, text_tokens=16, channels=8, seed=2)
    for i in range(5):
        d = tmp_path / ("a" if i < 3 else "b/c")
        d.mkdir(parents=True, exist_ok=True)
        # the reference saves each prompt with a leading batch dim of 1 and squeezes on load
        torch.save({k: emb[k][i:i + 1] for k in PIXART_KEYS} | {"unused": None}, d / f"p{i}.pt")
    ds = PromptEmbeddingDataset(tmp_path)
    assert len(ds) == 5
    item = ds[0]
    assert item["name"] == "p0" and item["relative_path"] == "a"
    assert item["prompt_embeds"].shape == (16, 8) and "unused" not in item
    assert ds[4]["relative_path"] == "b/c"
    batches = list(ds.batches(2))
    assert [len(b["name"]) for b in batches] == [2, 2, 1]
    got = torch.cat([b["prompt_embeds"] for b in batches])
    assert torch.equal(got, emb["prompt_embeds"])
    assert batches[0]["prompt_attention_mask"].dtype == torch.int64


def test_load_diffusers_state_dict_single_sharded_and_bin(tmp_path):
    import json

    import pytest
    import torch
    from safetensors.torch import save_file

    from ecad_b200.weights import load_diffusers_state_dict

    sd = {"a.weight": torch.arange(6, dtype=torch.float32).reshape(2, 3), "b.bias": torch.ones(4)}
    single = tmp_path / "single"
    single.mkdir()
    save_file(sd, str(single / "diffusion_pytorch_model.safetensors"))
    sharded = tmp_path / "pipe" / "transformer"
    sharded.mkdir(parents=True)
    save_file({"a.weight": sd["a.weight"]}, str(sharded / "diffusion_pytorch_model-00001-of-00002.safetensors"))
    save_file({"b.bias": sd["b.bias"]}, str(sharded / "diffusion_pytorch_model-00002-of-00002.safetensors"))
    (sharded / "diffusion_pytorch_model.safetensors.index.json").write_text(json.dumps({"weight_map": {
        "a.weight": "diffusion_pytorch_model-00001-of-00002.safetensors",
        "b.bias": "diffusion_pytorch_model-00002-of-00002.safetensors"}}))
    legacy = tmp_path / "legacy"
    legacy.mkdir()
    torch.save(sd, legacy / "diffusion_pytorch_model.bin")
    for d in (single, tmp_path / "pipe", legacy):
        got = load_diffusers_state_dict(d)
        assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    with pytest.raises(FileNotFoundError):
        load_diffusers_state_dict(tmp_path / "nothing")
