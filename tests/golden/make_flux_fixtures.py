"""Generate tests/golden/flux_schedules.json.gz from the reference's shipped FLUX schedule JSONs
(/root/reference/schedules/**; build container only).

Per file: path, name, S, NB (19 double blocks), NS (38 single blocks), bits = hex(np.packbits(flags[S][NB+NS][3])) with
rows = double blocks 0..18 then single_0..single_37 and columns in the reference's component order
(full_attn, full_ff, full_ff_context) / (single_attn, single_proj_mlp, single_proj_out)
(ecad/schedulers/cache_scheduler/flux_cache_schedule.py:51-90), config, tokens (256 or 4096 image tokens, from
config.height or the 256 default), per-step MACs and total (ecad/benchmark/compute_macs.py:255-303), attributes.
"""
import gzip
import json
from pathlib import Path

import numpy as np

REF = Path("/root/reference/schedules")
OUT = Path(__file__).parent / "flux_schedules.json.gz"
FULL = ["full_attn", "full_ff", "full_ff_context"]
SINGLE = ["single_attn", "single_proj_mlp", "single_proj_out"]


def pack(path: Path):
    d = json.loads(path.read_text())
    cs = d["cache_schedule"]
    S, NB, NS = cs["num_inference_steps"], cs["num_blocks"], cs["num_single_blocks"]
    flags = np.zeros((S, NB + NS, 3), dtype=np.bool_)
    for step, blocks in cs["schedule"].items():
        if int(step) >= S:  # a few files carry more step entries than num_inference_steps; the reference only
            continue        # ever indexes steps < num_inference_steps (compute_macs.py:279-293)
        for b, comp in blocks.items():
            if b.startswith("single_"):
                r, names = NB + int(b[len("single_"):]), SINGLE
            else:
                r, names = int(b), FULL
            for i, c in enumerate(names):
                flags[int(step), r, i] = comp[c]
    cfg = d.get("config") or {}
    h = cfg.get("height", 256)
    metrics = d.get("metrics") or {}
    by_step = metrics.get("by_inference_step")
    return {
        "path": str(path.relative_to(REF)), "name": cs["name"], "S": S, "NB": NB, "NS": NS,
        "bits": np.packbits(flags.reshape(-1)).tobytes().hex(), "config": cfg,
        "tokens": (h // 16) ** 2,
        "macs": [by_step[f"{s:03}"]["macs"] for s in range(S)] if by_step else None,
        "total_macs": metrics.get("total_macs"), "attributes": cs.get("attributes"),
        "latency_ms_a6000": (metrics.get("latency") or {}).get("avg"),
    }


def main():
    rows = [pack(p) for p in sorted(REF.rglob("*.json")) if "flux" in str(p.relative_to(REF))]
    payload = json.dumps({"source": "AniAggarwal/ecad schedules/ (FLUX)", "rows": rows}, separators=(",", ":"))
    with gzip.GzipFile(OUT, "wb", mtime=0) as f:
        f.write(payload.encode())
    print(f"{len(rows)} FLUX schedules, {sum(r['macs'] is not None for r in rows)} with MACs -> {OUT} "
          f"({OUT.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
