"""BASELINE config 1 latency: PixArt-alpha 256x256, ours_fast, batch 1 (2 CFG samples), 20 steps - ms per image through
the ImageGenerator API (CUDA events, after 2 warm-up generations).  The small-batch configuration is launch-bound."""
import json
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200.image_generator import B200PixArtAlphaImageGenerator  # noqa: E402
from ecad_b200.schedule import load_packed_schedules, schedule_from_packed  # noqa: E402
from ecad_b200.weights import synthetic_prompt_embeddings  # noqa: E402

rows = load_packed_schedules(ROOT / "tests" / "golden" / "pixart_schedules.json.gz")
row = [r for r in rows if r["path"] == "schedules_in_paper/pixart_alpha_256/ours_fast.json"][0]
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
gen = B200PixArtAlphaImageGenerator(cache_schedule=schedule_from_packed(row), use_cuda_graph=graph)
emb = {k: v.cuda() for k, v in synthetic_prompt_embeddings(batch).items()}
import contextlib
import io

with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(2):
        gen.generate_images(emb)
    torch.cuda.synchronize()
    times = []
    l0 = gen.diffusion_pipeline.transformer.launches
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gen.generate_images(emb)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
launches = (gen.diffusion_pipeline.transformer.launches - l0) / 5
print(json.dumps({"config": "c1", "batch": batch, "cuda_graph": graph, "ms_per_generation": statistics.mean(times),
                  "ms_per_image": statistics.mean(times) / batch, "launches_per_generation": launches}))
