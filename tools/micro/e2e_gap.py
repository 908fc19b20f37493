"""Diagnose the gap between the device-resident and the host-buffer (e2e) population paths."""
import contextlib
import io
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from ecad_b200.image_generator import B200PixArtAlphaImageGenerator  # noqa: E402
from ecad_b200.population import PopulationEvaluator  # noqa: E402
from ecad_b200.schedule import schedule_from_packed  # noqa: E402
from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings  # noqa: E402
from golden_util import rows  # noqa: E402

B, U = 100, 3
cands = [r for r in rows() if "population_initialization/pixart_alpha_256x256/gen_000/candidates/" in r["path"]][3:3 + U]
sd = random_init_state_dict(PixArtConfig(), 0)
gen = B200PixArtAlphaImageGenerator(cache_schedule=schedule_from_packed(cands[0]), state_dict=sd)
emb_host = {k: v.pin_memory() for k, v in synthetic_prompt_embeddings(B, seed=1).items()}
emb_dev = {k: v.cuda() for k, v in emb_host.items()}
ev = PopulationEvaluator(0, 1, torch.device("cuda:0"))


def run_unit(i, emb):
    gen.set_schedule(schedule_from_packed(cands[i]))
    return gen.generate_images(emb)[0]


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return a.elapsed_time(b), 1e3 * t_host, 1e3 * (time.perf_counter() - t0)


with contextlib.redirect_stdout(io.StringIO()):
    for i in range(U):
        run_unit(i, emb_dev)
    res = {}
    for name in ("device", "from_host", "device", "from_host", "h2d_only"):
        if name == "device":
            r = timed(lambda: [run_unit(i, emb_dev) for i in range(U)])
        elif name == "from_host":
            r = timed(lambda: ev.run_from_host(range(U), run_unit, emb_host))
        else:
            r = timed(lambda: [{k: v.to("cuda", non_blocking=True) for k, v in emb_host.items()} for _ in range(U)])
        res.setdefault(name, []).append(tuple(round(x, 1) for x in r))
print("(gpu ms between events, host ms to enqueue, wall ms):", res)
