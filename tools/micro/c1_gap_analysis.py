"""Where do the 28-33 ms of a batch-1 `ours_fast` generation go?  Sum of per-kernel CUDA-event durations (in-situ
profiler) versus the wall time of the generation, eager and graph-replayed."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from ecad_b200 import _lib  # noqa: E402
from ecad_b200.image_generator import B200PixArtAlphaImageGenerator  # noqa: E402
from ecad_b200.schedule import schedule_from_packed  # noqa: E402
from ecad_b200.weights import synthetic_prompt_embeddings  # noqa: E402
from golden_util import row_by_path  # noqa: E402

import contextlib, io  # noqa: E402

row = row_by_path("schedules_in_paper/pixart_alpha_256/ours_fast.json")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
emb = {k: v.cuda() for k, v in synthetic_prompt_embeddings(batch).items()}
out = {}
with contextlib.redirect_stdout(io.StringIO()):
    gen = B200PixArtAlphaImageGenerator(cache_schedule=schedule_from_packed(row))
    for _ in range(3):
        gen.generate_images(emb)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    gen.generate_images(emb)
    b.record()
    torch.cuda.synchronize()
    out["eager_ms"] = a.elapsed_time(b)
    _lib.profile_start()
    gen.generate_images(emb)
    prof = _lib.profile_stop()
out["kernel_ms_sum"] = sum(v["total_ms"] for v in prof.values())
out["launches"] = sum(v["launches"] for v in prof.values())
out["by_class"] = {k: {"launches": v["launches"], "ms": round(v["total_ms"], 3),
                       "avg_us": round(1e3 * v["total_ms"] / max(v["launches"], 1), 2)} for k, v in prof.items()}
print(json.dumps(out))
