"""FLUX decision layer: schedule object + decision trace against the per-step MACs the reference recorded in its
shipped FLUX schedules (38 files; same known-answer logic as tests/test_schedule_golden.py).  The FLUX CUDA blocks
are not built yet; this pins the bookkeeping they will run under."""
import gzip
import json
from pathlib import Path

import numpy as np
import pytest

from ecad_b200.macs import FluxShape, flux_macs_per_step
from ecad_b200.schedule import FluxCacheSchedule, trace_decisions

GOLDEN = Path(__file__).resolve().parent / "golden" / "flux_schedules.json.gz"


@pytest.fixture(scope="module")
def rows():
    with gzip.open(GOLDEN, "rb") as f:
        return json.loads(f.read())["rows"]


def _flags(r):
    n = r["S"] * (r["NB"] + r["NS"]) * 3
    return np.unpackbits(np.frombuffer(bytes.fromhex(r["bits"]), np.uint8))[:n].reshape(r["S"], r["NB"] + r["NS"], 3) \
        .astype(bool)


def test_all_flux_golden_mac_vectors_bit_exact(rows):
    n = 0
    bad = []
    for r in rows:
        if r["macs"] is None:
            continue
        n += 1
        ex = trace_decisions(_flags(r))
        m = flux_macs_per_step(ex, FluxShape(tokens=r["tokens"]))
        if list(map(int, m)) != r["macs"] or int(m.sum()) != r["total_macs"]:
            bad.append(r["path"])
    assert n == 38
    assert not bad, bad


def test_flux_schedule_object_round_trip(rows, tmp_path):
    r = [r for r in rows if r["path"].endswith("flux_256/ours_fast.json")][0]
    s = FluxCacheSchedule.from_numpy(_flags(r), r["S"], r["NB"], r["NS"], r["name"], r["config"])
    assert s.block_keys()[:2] == ["0", "1"] and s.block_keys()[-1] == "single_37"
    assert s.get_recompute("0", "full_attn") in (True, False)
    with pytest.raises(ValueError):
        s.get_recompute("0", "attn1")
    p = tmp_path / "f.json"
    s.to_json(p)
    d = json.loads(p.read_text())
    assert d["cache_schedule"]["num_single_blocks"] == 38
    s2 = FluxCacheSchedule.from_json(p)
    assert np.array_equal(s2.dense(), _flags(r))
    # genome: per step, 19x3 double-block flags then 38x3 single-block flags
    g = s2.to_numpy()
    assert g.shape == (20 * (19 * 3 + 38 * 3),)
    assert np.array_equal(g.reshape(20, 57, 3), _flags(r))
    with pytest.raises(NotImplementedError):
        s2.to_numpy(flatten=False)
    with pytest.raises(ValueError):
        FluxCacheSchedule(19, 20, "x", {})
