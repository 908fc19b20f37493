"""Schedule metrics in the reference's on-disk format: the ``metrics`` block of a schedule JSON that the NSGA-II
driver and the plotting / selection scripts read (``by_inference_step`` MACs, ``total_macs``, ``latency``).

Reference: ecad/benchmark/compute_macs.py:170-303 (calflops over one cached forward per step, batch 2, written as
``metrics.by_inference_step["%03d"] = {"flops", "macs"}`` + ``total_*``) and ecad/benchmark/compute_latency.py:20-85
(``metrics.latency = {avg, batch_size, num_samples, warmup_steps, gpu, warmups, latencies}``, ms per image, merged into
the existing block).  Here the MACs come from the decision trace and the analytic model of ecad_b200/macs.py - which
reproduces every ``macs`` value the reference recorded (tests/test_schedule_golden.py) - so no profiler pass is
needed; ``flops`` is calflops' own count (MACs*2 + elementwise) and is NOT reproduced: the key is written only when
the caller passes it through from an existing block.  Latency is measured with ``generate_images_timed`` exactly like
the reference (CUDA events around one pipeline call, divided by the batch size).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any

import numpy as np

from .macs import FluxShape, PixArtShape, flux_macs_per_step, macs_per_step
from .schedule import CacheSchedule, FluxCacheSchedule, PixArtCacheSchedule, trace_decisions


def executed_trace(schedule: CacheSchedule) -> np.ndarray:
    """``uint8[S][rows][3]`` executed/reused decisions of one generation under ``schedule`` (flag or cache-empty, TGATE
    rule for attn2 when the schedule carries it)."""
    if isinstance(schedule, FluxCacheSchedule):
        return trace_decisions(schedule.dense())
    gates = None
    if isinstance(schedule, PixArtCacheSchedule):
        # what the runtime decides from: the gate_step kwarg of each block's compute_attn_tgate entry
        # (cached_transformer_block.py:393-454); a schedule that only carries the pipeline-level TGATE config (no block
        # entry names the function) falls back to that value for every block
        gates = schedule.block_gate_steps()
        if (gates < 0).all():
            gates = schedule.gate_step()
    return trace_decisions(schedule.to_numpy(), attn2_tgate_gate_step=gates)


def macs_metrics(schedule: CacheSchedule, tokens: int = 256, text_tokens: int | None = None,
                 additional_conditions: bool = False) -> dict[str, Any]:
    """``by_inference_step`` / ``total_macs`` / ``total_macs_T`` of ``schedule`` in the reference's layout."""
    executed = executed_trace(schedule)
    if isinstance(schedule, FluxCacheSchedule):
        shape = FluxShape(tokens=tokens, text_tokens=text_tokens or 512, num_blocks=schedule.num_blocks,
                          num_single_blocks=schedule.num_single_blocks)
        per_step = flux_macs_per_step(executed, shape)
    else:
        pipe = (schedule.top_level_config or {}).get("pipeline") or {}
        gate = (pipe.get("kwargs") or {}).get("gate_step") if pipe.get("name") == "tgate" else None
        shape = PixArtShape(tokens=tokens, text_tokens=text_tokens or 120, additional_conditions=additional_conditions)
        per_step = macs_per_step(executed, shape, tgate_gate_step=gate)
    total = int(per_step.sum())
    return {
        "by_inference_step": {f"{s:03}": {"macs": int(m)} for s, m in enumerate(per_step)},
        "total_macs": total,
        "total_macs_T": total / 1e12,
    }


def latency_metrics(image_generator, prompt_embeds: dict, num_samples: int = 5, warmup_steps: int = 1) -> dict[str, Any]:
    """compute_latency.py:52-73: ``warmup_steps + num_samples`` timed generations of one batch, ms per image.  The
    reference's figure includes the VAE decode and the PIL conversion; a generator built with ``output_type="pt"`` /
    ``"pil"`` times the same span (the default ``"latent"`` stops after the denoising loop) - the block records which."""
    import torch

    times = [float(image_generator.generate_images_timed(prompt_embeds)) for _ in range(warmup_steps + num_samples)]
    warmups, latencies = times[:warmup_steps], times[warmup_steps:]
    return {
        "avg": sum(latencies) / len(latencies),
        "batch_size": int(prompt_embeds["prompt_embeds"].shape[0]),
        "num_samples": num_samples,
        "warmup_steps": warmup_steps,
        "gpu": torch.cuda.get_device_name(0),
        "output_type": getattr(image_generator, "output_type", "latent"),
        "warmups": warmups,
        "latencies": latencies,
    }


def merge_metrics(data: dict[str, Any], new: dict[str, Any]) -> dict[str, Any]:
    """compute_macs.py:227-233 / compute_latency.py:77-80: new keys replace old ones, everything else is kept.  A
    ``flops`` entry already recorded for a step by calflops survives next to the recomputed ``macs``."""
    old = dict(data.get("metrics") or {})
    for k, v in new.items():
        if k == "by_inference_step" and isinstance(old.get(k), dict):
            merged = {}
            for step, entry in v.items():
                e = dict(old[k].get(step) or {})
                e.update(entry)
                merged[step] = e
            old[k] = merged
        else:
            old[k] = v
    data["metrics"] = old
    return data


def annotate_schedule_file(schedule_file: Path | str, tokens: int | None = None, image_generator=None,
                           prompt_embeds: dict | None = None, num_samples: int = 5, warmup_steps: int = 1,
                           recompute_existing: bool = False) -> dict[str, Any]:
    """Fill ``metrics`` of a schedule JSON in place (MACs always; latency when a generator and embeddings are given)."""
    path = Path(schedule_file)
    data = json.loads(path.read_text())
    is_flux = "num_single_blocks" in data["cache_schedule"]
    schedule = (FluxCacheSchedule if is_flux else PixArtCacheSchedule).from_dict(data)
    cfg = data.get("config") or {}
    if tokens is None:
        h, w = cfg.get("height", 256), cfg.get("width", 256)
        tokens = (h // 16) * (w // 16)
    have = data.get("metrics") or {}
    if recompute_existing or "total_macs" not in have:
        merge_metrics(data, macs_metrics(schedule, tokens=tokens))
    if image_generator is not None and prompt_embeds is not None and (recompute_existing or "latency" not in have):
        image_generator.set_schedule(schedule)
        merge_metrics(data, {"latency": latency_metrics(image_generator, prompt_embeds, num_samples, warmup_steps)})
    path.write_text(json.dumps(data, indent=4))
    return data["metrics"]
