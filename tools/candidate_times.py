"""Seconds per generation of each of the 72 seed-population candidates at batch 100 (config 2), next to the features
an LPT cost model can use: executed attn1 / attn2 / ff sub-blocks and reused sub-blocks over the 20 steps.

    python tools/candidate_times.py [batch] > gpurun_out/candidate_times.json

The fit (least squares, printed at the end and stored under "fit") is what ecad_b200.macs.schedule_cost uses to
partition candidates over ranks (ecad_b200.population.partition_lpt): analytic FLOPs alone under-estimate the cheap,
reuse-dominated candidates, which run HBM-bound."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200.image_generator import B200PixArtAlphaImageGenerator  # noqa: E402
from ecad_b200.macs import PixArtShape, flops_per_image  # noqa: E402
from ecad_b200.schedule import load_packed_schedules, schedule_from_packed, trace_decisions  # noqa: E402
from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rows = load_packed_schedules(ROOT / "tests" / "golden" / "pixart_schedules.json.gz")
cands = sorted((r for r in rows if "population_initialization/pixart_alpha_256x256/gen_000/candidates/" in r["path"]),
               key=lambda r: r["path"])
assert len(cands) == 72
sd = random_init_state_dict(PixArtConfig(), 0)
gen = B200PixArtAlphaImageGenerator(cache_schedule=schedule_from_packed(cands[0]), state_dict=sd)
emb = {k: v.cuda() for k, v in synthetic_prompt_embeddings(B, seed=1).items()}
shape = PixArtShape()
for r in cands[:3]:  # warm-up
    gen.set_schedule(schedule_from_packed(r))
    gen.generate_images(emb)
torch.cuda.synchronize()
out = []
for r in cands:
    sched = schedule_from_packed(r)
    ex = trace_decisions(sched.to_numpy()).astype(np.int64)
    gen.set_schedule(sched)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    gen.generate_images(emb)
    b.record()
    torch.cuda.synchronize()
    out.append({"path": r["path"], "seconds": a.elapsed_time(b) * 1e-3, "executed": ex.sum(axis=(0, 1)).tolist(),
                "reused": int((1 - ex).sum()), "tflop_per_image": flops_per_image(ex, shape) / 1e12})
X = np.array([o["executed"] + [o["reused"], 1.0] for o in out], dtype=np.float64)
y = np.array([o["seconds"] for o in out])
coef, *_ = np.linalg.lstsq(X, y, rcond=None)
pred = X @ coef
flops_only = np.array([o["tflop_per_image"] for o in out])
k = float((flops_only @ y) / (flops_only @ flops_only))
print(json.dumps({
    "batch": B, "candidates": out,
    "fit": {"features": ["executed attn1", "executed attn2", "executed ff", "reused sub-blocks", "constant"],
            "seconds_per_unit": coef.tolist(),
            "max_rel_error": float(np.abs(pred - y).max() / y.mean()), "rms_rel_error": float(np.sqrt(((pred - y) ** 2).mean()) / y.mean()),
            "flops_only_max_rel_error": float(np.abs(k * flops_only - y).max() / y.mean()),
            "flops_only_rms_rel_error": float(np.sqrt(((k * flops_only - y) ** 2).mean()) / y.mean())}}))
