"""ecad_b200/metrics.py writes the reference's ``metrics`` block (compute_macs.py / compute_latency.py format); the MAC
part is checked against the values the reference itself recorded in its shipped schedule files."""
import json

import numpy as np

from ecad_b200.metrics import annotate_schedule_file, executed_trace, macs_metrics, merge_metrics
from ecad_b200.schedule import FluxCacheSchedule
from golden_util import row_by_path, schedule_of


def test_macs_metrics_match_reference_recorded_values():
    for path, tokens in (("schedules_in_paper/pixart_alpha_256/ours_fast.json", 256),
                         ("alpha_cache_schedules/gen_tgate/tgate_m_010_sp_001_fi_001_warmup_002.json", 256)):
        r = row_by_path(path)
        m = macs_metrics(schedule_of(r), tokens=tokens)
        assert [m["by_inference_step"][f"{s:03}"]["macs"] for s in range(r["S"])] == r["macs"]
        assert m["total_macs"] == r["total_macs"] and abs(m["total_macs_T"] - r["total_macs"] / 1e12) < 1e-12
    assert macs_metrics(schedule_of(row_by_path("pixart_alpha_256/ours_fast.json")))["total_macs"] == 2_134_989_471_744


def test_flux_macs_metrics_match_reference_recorded_values():
    import gzip
    from pathlib import Path

    rows = json.loads(gzip.open(Path(__file__).parent / "golden" / "flux_schedules.json.gz").read())["rows"]
    r = [r for r in rows if r["path"].endswith("flux_256/ours_fast.json")][0]
    n = r["S"] * (r["NB"] + r["NS"]) * 3
    flags = np.unpackbits(np.frombuffer(bytes.fromhex(r["bits"]), np.uint8))[:n].reshape(r["S"], -1, 3).astype(bool)
    sched = FluxCacheSchedule.from_numpy(flags, r["S"], r["NB"], r["NS"], r["name"])
    m = macs_metrics(sched, tokens=r["tokens"])
    assert [v["macs"] for v in m["by_inference_step"].values()] == r["macs"]
    assert executed_trace(sched)[0].all()


def test_annotate_file_merges_like_the_reference(tmp_path):
    r = row_by_path("schedules_in_paper/pixart_alpha_256/ours_fast.json")
    sched = schedule_of(r)
    data = sched.to_dict()
    data["metrics"] = {"by_inference_step": {"000": {"flops": 571581041148}}, "latency": {"avg": 84.09, "gpu": "NVIDIA RTX A6000"},
                       "custom_note": "kept"}
    f = tmp_path / "ours_fast.json"
    f.write_text(json.dumps(data))
    m = annotate_schedule_file(f)
    back = json.loads(f.read_text())
    assert back["metrics"] == m and back["cache_schedule"] == data["cache_schedule"]
    assert m["by_inference_step"]["000"] == {"flops": 571581041148, "macs": r["macs"][0]}  # calflops' flops survive
    assert m["latency"]["gpu"] == "NVIDIA RTX A6000" and m["custom_note"] == "kept"       # untouched without a generator
    assert m["total_macs"] == r["total_macs"]
    # a second pass without recompute_existing leaves the file alone
    assert annotate_schedule_file(f) == m
    assert merge_metrics({"metrics": {"a": 1}}, {"b": 2})["metrics"] == {"a": 1, "b": 2}
