"""Analytic work model of the PixArt hot path.

Two conventions, both per *sample* (one CFG image = 2 samples per step):

* ``calflops`` convention - Linear/Conv MACs only, no SDPA, no norms - the one the reference's
  ecad/benchmark/compute_macs.py:255-303 records into ``metrics.by_inference_step`` of the shipped schedule
  JSONs.  Reproducing those numbers from a decision trace is the known-answer test for the compute/reuse
  decisions.
* algorithmic FLOPs - 2 x MACs *including* the SDPA matmuls - the numerator of the tensor roofline that
  bench.py reports.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class PixArtShape:
    tokens: int = 256  # N image tokens = (H/16)*(W/16)
    text_tokens: int = 120  # T (120 alpha, 300 sigma)
    dim: int = 1152  # D
    heads: int = 16
    head_dim: int = 72
    ff_mult: int = 4
    caption_channels: int = 4096
    patch_in: int = 16  # in_channels * patch^2 = 4*2*2
    patch_out: int = 32  # patch^2 * out_channels = 2*2*8
    additional_conditions: bool = False  # 1024-MS alpha only (resolution + aspect-ratio embedders)

    # ---- calflops convention (MACs of nn.Linear / nn.Conv2d only) --------------------------------
    def macs_attn1(self) -> int:
        return 4 * self.tokens * self.dim**2

    def macs_attn2(self) -> int:
        return 2 * self.tokens * self.dim**2 + 2 * self.text_tokens * self.dim**2

    def macs_ff(self) -> int:
        return 2 * self.ff_mult * self.tokens * self.dim**2

    def macs_fixed(self) -> int:
        N, D, T = self.tokens, self.dim, self.text_tokens
        fixed = (
            self.patch_in * N * D  # patch-embed conv
            + 256 * D + D * D  # timestep MLP
            + 6 * D * D  # adaLN-single linear
            + T * (self.caption_channels * D + D * D)  # caption projection
            + self.patch_out * N * D  # proj_out
        )
        if self.additional_conditions:
            fixed += 3 * (256 * 384 + 384 * 384)
        return fixed

    def macs_components(self) -> np.ndarray:
        return np.array([self.macs_attn1(), self.macs_attn2(), self.macs_ff()], dtype=np.int64)

    # ---- algorithmic FLOPs (2*MACs incl. SDPA) ----------------------------------------------------
    def flops_attn1(self) -> int:
        return 2 * (4 * self.tokens * self.dim**2 + 2 * self.tokens**2 * self.dim)

    def flops_attn2(self) -> int:
        N, D, T = self.tokens, self.dim, self.text_tokens
        return 2 * (2 * N * D * D + 2 * T * D * D + 2 * N * T * D)

    def flops_ff(self) -> int:
        return 2 * self.macs_ff()

    def flops_fixed(self) -> int:
        return 2 * self.macs_fixed()

    def flops_components(self) -> np.ndarray:
        return np.array([self.flops_attn1(), self.flops_attn2(), self.flops_ff()], dtype=np.int64)


def macs_per_step(
    executed: np.ndarray, shape: PixArtShape, batch: int = 2, tgate_gate_step: int | None = None
) -> np.ndarray:
    """``macs[step] = batch_s * (fixed + sum_{b,c} executed[s][b][c] * comp_c)``.

    ``batch`` is the calflops input batch (2 = one CFG pair, compute_macs.py:40); under the TGATE pipeline the
    batch drops to 1 from ``gate_step`` on (ecad/pipelines/tgate.py:329-341).
    """
    executed = np.asarray(executed)
    comp = shape.macs_components()
    per = shape.macs_fixed() + (executed.astype(np.int64) * comp[None, None, :]).sum(axis=(1, 2))
    b = np.full(executed.shape[0], batch, dtype=np.int64)
    if tgate_gate_step is not None:
        b[tgate_gate_step:] = batch // 2
    return per * b


def flops_per_image(executed: np.ndarray, shape: PixArtShape, samples_per_image: int = 2) -> int:
    """Algorithmic FLOPs of one generated image (all steps, both CFG samples) under a decision trace."""
    executed = np.asarray(executed)
    comp = shape.flops_components()
    per_step = shape.flops_fixed() + (executed.astype(np.int64) * comp[None, None, :]).sum(axis=(1, 2))
    return int(per_step.sum()) * samples_per_image


# Measured B200 rates behind `b200_seconds_per_image` (profiles/r2_candidate_times.json: the 72 seed-population
# candidates at batch 100, least-squares over executed attn1 / attn2 / ff and reused sub-blocks; rms error 0.7 % of the
# mean candidate time against 2.5 % for a FLOPs-only model).  Algorithmic TFLOP/s of an executed sub-block including
# its LayerNorm / glue, and the effective HBM rate of a reused sub-block (cache read + its share of the stream update).
B200_SUBBLOCK_TFLOPS = (901.0, 967.0, 1063.0)  # attn1, attn2, ff
B200_REUSE_BYTES_PER_S = 2.7e12


def b200_seconds_per_image(executed: np.ndarray, shape: PixArtShape, samples_per_image: int = 2) -> float:
    """Estimated B200 seconds per generated image under a decision trace - the COST the population partition uses
    (ecad_b200.population.partition_lpt): cheap, reuse-dominated candidates run HBM-bound, which analytic FLOPs alone
    under-estimate."""
    executed = np.asarray(executed).astype(np.int64)
    counts = executed.sum(axis=(0, 1))  # executed attn1, attn2, ff sub-blocks over the generation
    reused = int((1 - executed).sum())
    comp = shape.flops_components()
    t = sum(float(counts[c]) * float(comp[c]) / (B200_SUBBLOCK_TFLOPS[c] * 1e12) for c in range(3))
    t += reused * (shape.tokens * shape.dim * 2) / B200_REUSE_BYTES_PER_S
    t += executed.shape[0] * shape.flops_fixed() / (B200_SUBBLOCK_TFLOPS[2] * 1e12)
    return t * samples_per_image


@dataclass(frozen=True)
class FluxShape:
    """FLUX.1-dev MAC model in the reference's calflops convention."""

    tokens: int = 256  # image tokens N = (H/16)*(W/16)
    text_tokens: int = 512
    dim: int = 3072
    num_blocks: int = 19
    num_single_blocks: int = 38

    def macs_components(self) -> np.ndarray:
        """int64[NB + NS][3] per-sample MACs of each cacheable component, rows in FluxCacheSchedule.dense() order."""
        N, T, D = self.tokens, self.text_tokens, self.dim
        S = N + T
        full = [4 * S * D * D, 8 * N * D * D, 8 * T * D * D]
        single = [3 * S * D * D, 4 * S * D * D, 5 * S * D * D]
        return np.array([full] * self.num_blocks + [single] * self.num_single_blocks, dtype=np.int64)

    def flops_components(self) -> np.ndarray:
        """Algorithmic FLOPs per sample = 2 * MACs + the SDPA matmuls (2 * S^2 * D MACs per joint attention), the
        convention SURVEY.md section 8d uses for every roofline figure."""
        S, D = self.tokens + self.text_tokens, self.dim
        f = 2 * self.macs_components()
        f[:, 0] += 4 * S * S * D
        return f

    def flops_always(self) -> int:
        return 2 * self.macs_always()

    def macs_always(self) -> int:
        """Modulation linears that run even when every component is reused + the embedders / output head."""
        N, T, D = self.tokens, self.text_tokens, self.dim
        per_block = 12 * D * D * self.num_blocks + 3 * D * D * self.num_single_blocks
        fixed = 64 * N * D + T * 4096 * D + 2 * (256 * D + D * D) + (768 * D + D * D) + 2 * D * D + 64 * N * D
        return per_block + fixed


def flux_macs_per_step(executed: np.ndarray, shape: FluxShape, batch: int = 2) -> np.ndarray:
    executed = np.asarray(executed)
    comp = shape.macs_components()
    per = shape.macs_always() + (executed.astype(np.int64) * comp[None]).sum(axis=(1, 2))
    return per * batch
