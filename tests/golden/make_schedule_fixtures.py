"""Generate tests/golden/pixart_schedules.json.gz from the reference's shipped schedule JSONs.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_schedule_fixtures.py

What is recorded per PixArt schedule file (reference: /root/reference/schedules/**.json):
  * path        relative path under schedules/
  * name, S (num_inference_steps), NB (num_blocks)
  * bits        hex of np.packbits(flags[S][NB][3]) in the reference's genome order
                (ecad/schedulers/cache_scheduler/pixart_cache_schedule.py:15-27: [step][block][attn1,attn2,ff])
  * custom      {"attn": name, "gate_step": g} when every (step, block) carries the same
                custom_compute_attn entry (ecad/types.py:59-64), else null
  * config      the top-level `config` object verbatim (ecad/types.py:43-47)
  * tokens      image tokens N the metrics were measured at (256, or 4096 for the *1024* families).
                NOTE: population_initialization/pixart_alpha_256x256 carries a stale 1024-MS config while
                its metrics are 256-token numbers - tokens is 256 there.
  * macs        metrics.by_inference_step[*].macs (ecad/benchmark/compute_macs.py:255-303), or null
  * total_macs  metrics.total_macs, or null
  * attributes  cache_schedule.attributes verbatim

These are the reference's only known-answer vectors for the compute/reuse decisions.
"""
import gzip
import json
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/schedules")
OUT = Path(__file__).parent / "pixart_schedules.json.gz"
COMPONENTS = ["attn1", "attn2", "ff"]


def pack(path: Path):
    d = json.loads(path.read_text())
    cs = d["cache_schedule"]
    S, NB = cs["num_inference_steps"], cs["num_blocks"]
    flags = np.zeros((S, NB, 3), dtype=np.bool_)
    customs = set()
    for step, blocks in cs["schedule"].items():
        for b, comp in blocks.items():
            for i, c in enumerate(COMPONENTS):
                flags[int(step), int(b), i] = comp[c]
            cc = comp.get("custom_compute_attn")
            customs.add(json.dumps(cc, sort_keys=True) if cc else "")
            assert "custom_compute_ff" not in comp, path
    custom = None
    if customs != {""}:
        assert len(customs) == 1, (path, customs)
        cc = json.loads(next(iter(customs)))
        custom = {"attn": cc["name"], "gate_step": cc.get("kwargs", {}).get("gate_step")}
    rel = str(path.relative_to(REF))
    tokens = 4096 if ("gen_tgate_1024" in rel or "gen_default_1024x1024" in rel) else 256
    metrics = d.get("metrics") or {}
    by_step = metrics.get("by_inference_step")
    macs = None
    if by_step is not None:
        macs = [by_step[f"{s:03}"]["macs"] for s in range(S)]
    return {
        "path": rel,
        "name": cs["name"],
        "S": S,
        "NB": NB,
        "bits": np.packbits(flags.reshape(-1)).tobytes().hex(),
        "custom": custom,
        "config": d.get("config"),
        "tokens": tokens,
        "macs": macs,
        "total_macs": metrics.get("total_macs"),
        "latency_ms_a6000": (metrics.get("latency") or {}).get("avg"),
        "attributes": cs.get("attributes"),
    }


def main():
    rows = []
    for p in sorted(REF.rglob("*.json")):
        if "flux" in str(p.relative_to(REF)):
            continue
        rows.append(pack(p))
    payload = json.dumps({"source": "AniAggarwal/ecad schedules/", "rows": rows}, separators=(",", ":"))
    with gzip.GzipFile(OUT, "wb", mtime=0) as f:
        f.write(payload.encode())
    n_metrics = sum(r["macs"] is not None for r in rows)
    print(f"{len(rows)} PixArt schedules, {n_metrics} with per-step MACs -> {OUT} ({OUT.stat().st_size} bytes)")


if __name__ == "__main__":
    sys.exit(main())
