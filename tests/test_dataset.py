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
