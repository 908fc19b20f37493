"""Whole-generation CUDA graphs for the launch-bound small-batch configuration (BASELINE config 1: batch 1).

A 20-step PixArt generation at batch 1 is ~2300 kernel launches of a few microseconds each; issued one by one through
ctypes the host is the bottleneck.  The executed/reused decisions of a (schedule, shape) pair are the same for every
generation - caches start empty and the reset callback runs last (ecad/image_generators/image_generator.py:193-202) -
so the whole denoising loop (all steps: embedders, blocks under each step's decision row, solver update) is recorded
ONCE into a CUDA graph over static input buffers and replayed with one launch per generation.  The reference has no
counterpart (it issues every torch op eagerly each step).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Callable

import torch


class GenerationGraphs:
    def __init__(self, max_entries: int = 4):
        self.max_entries = max_entries
        self._entries: OrderedDict[Any, tuple] = OrderedDict()
        self.captures = 0
        self.replays = 0

    def clear(self) -> None:
        self._entries.clear()

    def keep_only(self, key: Any) -> None:
        """Drop every recorded graph except ``key`` (the buffers the others point into have just been re-allocated)."""
        for k in [k for k in self._entries if k != key]:
            del self._entries[k]

    def run(self, key: Any, inputs: dict[str, Any], body: Callable[[dict[str, Any], Any], torch.Tensor],
            capture_callback, transformer) -> torch.Tensor:
        """``body(static_inputs, callback)`` runs the denoising loop in place on ``static_inputs["latents"]`` and
        returns the final latents tensor.  ``capture_callback`` must advance the schedule step counters and reset them
        (and the transformer cache validity) after the last step; it must not read device memory."""
        entry = self._entries.get(key)
        if entry is None:
            static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in inputs.items()}
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                body(static, capture_callback)  # eager warm-up: workspace allocation, kernel attributes
            cur.wait_stream(side)
            torch.cuda.synchronize()
            for k, v in inputs.items():
                if torch.is_tensor(v):
                    static[k].copy_(v)
            l0 = transformer.launches
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = body(static, capture_callback)
            launches = transformer.launches - l0
            transformer.launches = l0  # the capture pass recorded kernels, it did not run them
            entry = (graph, static, out, launches)
            self._entries[key] = entry
            self.captures += 1
            while len(self._entries) > self.max_entries:
                self._entries.popitem(last=False)
        graph, static, out, launches = entry
        for k, v in inputs.items():
            if torch.is_tensor(v):
                static[k].copy_(v, non_blocking=True)
        graph.replay()
        self.replays += 1
        transformer.launches += launches
        return out.clone()
