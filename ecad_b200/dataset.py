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

    def batches(self, batch_size: int, pin_memory: bool = False, shuffle: bool = False,
                seed: int | None = None) -> "EmbeddingBatches":
        """What the reference gets from ``DataLoader(dataset, batch_size, shuffle, pin_memory=True)``
        (image_generator.py:278-300): every tensor key of the per-prompt dicts stacked along a new batch dimension
        (PixArt: the four PIXART_KEYS; FLUX: prompt_embeds / pooled_prompt_embeds), string keys collected into lists;
        nothing is dropped, the last batch may be short.  The result has a ``len()`` like a DataLoader."""
        return EmbeddingBatches(self, batch_size, pin_memory, shuffle, seed)


class EmbeddingBatches:
    def __init__(self, dataset: PromptEmbeddingDataset, batch_size: int, pin_memory: bool = False,
                 shuffle: bool = False, seed: int | None = None):
        if batch_size < 1:
            raise ValueError("batch_size must be >= 1")
        self.dataset, self.batch_size, self.pin_memory, self.shuffle, self.seed = (dataset, batch_size, pin_memory,
                                                                                   shuffle, seed)

    def __len__(self) -> int:
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[dict[str, torch.Tensor | list[str]]]:
        order = list(range(len(self.dataset)))
        if self.shuffle:
            g = torch.Generator()
            if self.seed is not None:
                g.manual_seed(self.seed)
            order = torch.randperm(len(order), generator=g).tolist()
        pin = self.pin_memory and torch.cuda.is_available()
        for start in range(0, len(order), self.batch_size):
            items = [self.dataset[i] for i in order[start:start + self.batch_size]]
            batch: dict[str, torch.Tensor | list[str]] = {}
            for key in items[0]:
                vals = [it[key] for it in items]
                if isinstance(vals[0], torch.Tensor):
                    t = torch.stack(vals)
                    batch[key] = t.pin_memory() if pin else t
                else:
                    batch[key] = vals
            yield batch
