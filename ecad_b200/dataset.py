"""Saved prompt-embedding files -> batches for the generator.

Mirror of /root/reference/ecad/dataset_utils/prompt_embedding_dataset.py:9-60 (one ``*.pt`` dict per prompt with
``prompt_embeds`` / ``prompt_attention_mask`` / ``negative_prompt_embeds`` / ``negative_prompt_attention_mask``,
ecad/types.py:14-18) plus the batching the reference gets from a DataLoader over it
(ecad/image_generators/image_generator.py:278-300).  Host-side only.
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterator

import torch

PIXART_KEYS = ("prompt_embeds", "prompt_attention_mask", "negative_prompt_embeds", "negative_prompt_attention_mask")


class PromptEmbeddingDataset:
    def __init__(self, embedding_dir: str | Path):
        self.embedding_dir = Path(embedding_dir)
        self.filenames = sorted(self.embedding_dir.glob("**/*.pt"))

    def _get_relative_path(self, idx: int) -> str:
        return str(self.filenames[idx].relative_to(self.embedding_dir).parent)

    def __len__(self) -> int:
        return len(self.filenames)

    def __getitem__(self, idx: int) -> dict[str, torch.Tensor | str]:
        loaded = torch.load(self.filenames[idx], weights_only=True, map_location="cpu")
        out: dict[str, torch.Tensor | str] = {
            "name": self.filenames[idx].stem,
            "relative_path": self._get_relative_path(idx),
        }
        for key, value in loaded.items():
            if value is None:
                continue
            out[key] = value.squeeze() if isinstance(value, torch.Tensor) else value
        return out

    def batches(self, batch_size: int, pin_memory: bool = False) -> Iterator[dict[str, torch.Tensor | list[str]]]:
        """Stack consecutive prompts into PixArtPromptEmbedding batches (drop nothing; the last batch may be short)."""
        for start in range(0, len(self), batch_size):
            items = [self[i] for i in range(start, min(start + batch_size, len(self)))]
            batch: dict[str, torch.Tensor | list[str]] = {
                "name": [it["name"] for it in items],
                "relative_path": [it["relative_path"] for it in items],
            }
            for key in PIXART_KEYS:
                t = torch.stack([it[key] for it in items])
                batch[key] = t.pin_memory() if pin_memory and torch.cuda.is_available() else t
            yield batch
