"""Experiment: evaluate TWO candidate schedules concurrently on one GPU (two resident models, two CUDA streams, two host
threads) versus one after the other.  Units of the population evaluation are independent, so kernels of one unit can fill
the tails / co-run with the memory-bound kernels of the other.

    python tools/dual_stream_bench.py [prompts_per_unit=100] [units=6]
"""
import json
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from ecad_b200.image_generator import B200PixArtAlphaImageGenerator  # noqa: E402
from ecad_b200.schedule import schedule_from_packed  # noqa: E402
from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings  # noqa: E402
from golden_util import rows  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
U = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cands = [r for r in rows() if "population_initialization/pixart_alpha_256x256/gen_000/candidates/" in r["path"]][:U]
sd = random_init_state_dict(PixArtConfig(), 0)
emb = {k: v.cuda() for k, v in synthetic_prompt_embeddings(B, seed=1).items()}
import contextlib, io  # noqa: E402


def make():
    g = B200PixArtAlphaImageGenerator(cache_schedule=schedule_from_packed(cands[0]), state_dict=sd)
    g.create_diffusion_pipeline()
    return g


def run_units(gen, idx, stream):
    with torch.cuda.stream(stream), contextlib.redirect_stdout(io.StringIO()):
        for i in idx:
            gen.set_schedule(schedule_from_packed(cands[i]))
            gen.generate_images(emb)


gens = [make(), make()]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for g, s in zip(gens, streams):
    run_units(g, [0], s)
torch.cuda.synchronize()

res = {}
for mode in ("sequential", "concurrent", "sequential", "concurrent"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if mode == "sequential":
        run_units(gens[0], list(range(U)), streams[0])
    else:
        th = [threading.Thread(target=run_units, args=(gens[k], list(range(k, U, 2)), streams[k])) for k in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res.setdefault(mode, []).append(U * B / dt)
print(json.dumps({"prompts_per_unit": B, "units": U, "images_per_s": res}))
