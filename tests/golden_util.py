"""Access to the committed golden fixtures (tests/golden/)."""
from pathlib import Path

import numpy as np

from ecad_b200.schedule import load_packed_schedules, schedule_from_packed

GOLDEN = Path(__file__).resolve().parent / "golden"
_rows = None


def rows():
    global _rows
    if _rows is None:
        _rows = load_packed_schedules(GOLDEN / "pixart_schedules.json.gz")
    return _rows


def row_by_path(suffix: str):
    hits = [r for r in rows() if r["path"].endswith(suffix)]
    assert len(hits) == 1, (suffix, len(hits))
    return hits[0]


def flags_of(row) -> np.ndarray:
    S, NB = row["S"], row["NB"]
    bits = np.unpackbits(np.frombuffer(bytes.fromhex(row["bits"]), np.uint8))[: S * NB * 3]
    return bits.reshape(S, NB, 3).astype(bool)


def schedule_of(row):
    return schedule_from_packed(row)
