"""BASELINE config 5 (FLUX.1-dev 1024x1024, batch 4; a parity-test configuration, not the bench line):

    python tools/flux_bench.py dense [batch] [px]   # one dense forward (all 171 components executed)
    python tools/flux_bench.py gen [batch] [px] [schedule-suffix]
                                              # 20-step generations under a shipped flux schedule (default
                                              # schedules_in_paper/flux_256_to_1024/fast_256_to_1024.json)

Random-init FLUX.1-dev weights drawn in HBM (12 B parameters), synthetic T5/CLIP embeddings.  Prints ms, algorithmic
TFLOP/s (ecad_b200/macs.py FluxShape: 2*MACs + SDPA) against the measured sustained bf16 peak, and the in-situ
CUDA-event split by kernel class.
"""
import gzip
import json
import statistics
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402
from ecad_b200.image_generator import B200FluxImageGenerator  # noqa: E402
from ecad_b200.macs import FluxShape  # noqa: E402
from ecad_b200.schedule import FluxCacheSchedule, trace_decisions  # noqa: E402
from ecad_b200.weights import FluxConfig  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "dense"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
px = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
suffix = sys.argv[4] if len(sys.argv) > 4 else "schedules_in_paper/flux_256_to_1024/fast_256_to_1024.json"
cfg = FluxConfig()
N, T = (px // 16) ** 2, 512
shape = FluxShape(tokens=N)
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
peak = float(peaks.get("bf16_tflops_sustained", 1400.0))

if mode == "dense":
    steps, flags = 1, np.ones((1, 57, 3), bool)
    name = "dense"
else:
    rows = json.loads(gzip.open(ROOT / "tests" / "golden" / "flux_schedules.json.gz").read())["rows"]
    r = [r for r in rows if r["path"].endswith(suffix)][0]
    steps = r["S"]
    flags = np.unpackbits(np.frombuffer(bytes.fromhex(r["bits"]), np.uint8))[: steps * 57 * 3].reshape(steps, 57, 3) \
        .astype(bool)
    name = r["path"]
sched = FluxCacheSchedule.from_numpy(flags, steps, 19, 38, name, top_level_config={"height": px, "width": px})
gen = B200FluxImageGenerator(cache_schedule=sched, model_config=cfg, weights_on_device=True)
g = torch.Generator().manual_seed(1)
emb = {"prompt_embeds": (torch.randn(B, T, 4096, generator=g) * 0.2).cuda(),
       "pooled_prompt_embeds": (torch.randn(B, 768, generator=g) * 0.2).cuda()}
executed = trace_decisions(flags)
flops = B * int((executed.astype(np.int64) * shape.flops_components()[None]).sum() + steps * shape.flops_always())

gen.generate_images(emb)  # warm-up (weights, workspace, TMA descriptors)
torch.cuda.synchronize()
times = []
for _ in range(2 if mode == "gen" else 4):
    times.append(gen.generate_images_timed(emb) * B)
_lib.profile_start()
gen.generate_images(emb)
prof = _lib.profile_stop()
ms = statistics.mean(times)
tf = flops / ms / 1e9
tr = gen.diffusion_pipeline.transformer
out = {"config": "c5", "mode": mode, "schedule": name, "batch": B, "px": px, "img_tokens": N, "steps": steps,
       "ms": ms, "ms_all": times, "images_per_s": B / (ms / 1e3) if mode == "gen" else None,
       "algorithmic_tflop": flops / 1e12, "tflops": tf, "frac_of_sustained_bf16": tf / peak,
       "executed_fraction": float(executed.mean()),
       "hbm_gb": torch.cuda.max_memory_allocated() / 1e9,
       "in_situ": {k: {"launches": v["launches"], "ms": round(v["total_ms"], 3),
                       "tflops": (v["flops"] / v["total_ms"] / 1e9) if v["total_ms"] > 0 and v["flops"] else None}
                   for k, v in prof.items()}}
print(json.dumps(out))
