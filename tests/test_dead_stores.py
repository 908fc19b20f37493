"""Safety of the dead-cache-store elimination, simulated on the host over the reference's shipped schedules: with the
look-ahead rule of ecad_b200.schedule.*_dead_store_mask no reuse ever reads a slot whose store was skipped, the TGATE
averaging always finds its attn2 slots written, and the rule really removes stores."""
import gzip
import json
from pathlib import Path

import numpy as np

from ecad_b200.schedule import (FluxCacheSchedule, flux_dead_store_mask, pixart_dead_store_mask, trace_decisions)
from golden_util import flags_of, rows, schedule_of


def _gate(row):
    pipe = (row["config"] or {}).get("pipeline") or {}
    return pipe.get("kwargs", {}).get("gate_step") if pipe.get("name") == "tgate" else None


def test_pixart_dead_stores_never_starve_a_reuse():
    checked = skipped = executed_total = 0
    all_rows = rows()
    sample = all_rows[::7] + [r for r in all_rows if _gate(r) is not None][::9]
    for r in sample:
        sched = schedule_of(r)
        gate = _gate(r)
        ex = trace_decisions(flags_of(r), gate).astype(bool)
        S, NB = r["S"], r["NB"]
        written = np.zeros((NB, 3), bool)
        for s in range(S):
            keep = list(range(NB)) if (gate is not None and s == gate - 1) else []
            dead = pixart_dead_store_mask(sched, s, ex[s], keep).astype(bool)
            assert not (dead & ~ex[s]).any()                       # only executed sub-blocks are ever marked
            assert written[~ex[s]].all(), (r["path"], s)           # every reuse reads a slot that holds data
            written = np.where(ex[s], ~dead, written)
            if keep:
                assert written[:, 1].all(), (r["path"], s)         # TGATE averages the attn2 caches after this step
            skipped += int(dead.sum())
            executed_total += int(ex[s].sum())
        assert pixart_dead_store_mask(sched, S - 1, ex[S - 1])[ex[S - 1]].all() or gate is not None
        checked += 1
    assert checked > 200 and skipped > 0.3 * executed_total, (checked, skipped, executed_total)


def test_flux_dead_stores_never_starve_a_reuse():
    data = json.loads(gzip.open(Path(__file__).parent / "golden" / "flux_schedules.json.gz").read())["rows"]
    skipped = 0
    for r in data[::4]:
        n = r["S"] * (r["NB"] + r["NS"]) * 3
        flags = np.unpackbits(np.frombuffer(bytes.fromhex(r["bits"]), np.uint8))[:n].reshape(r["S"], -1, 3).astype(bool)
        sched = FluxCacheSchedule.from_numpy(flags, r["S"], r["NB"], r["NS"], r["name"])
        ex = trace_decisions(flags).astype(bool)
        written = np.zeros(ex.shape[1:], bool)
        for s in range(r["S"]):
            dead = flux_dead_store_mask(sched, s, ex[s]).astype(bool)
            assert not dead[r["NB"]:, :2].any()                    # produced in place: never skipped
            assert written[~ex[s]].all(), (r["path"], s)
            written = np.where(ex[s], ~dead, written)
            skipped += int(dead.sum())
    assert skipped > 0
