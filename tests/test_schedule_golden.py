"""Known-answer test for the compute/reuse DECISIONS.

The reference's `compute_macs.py` ran the real cached forward under every shipped schedule and recorded per-step MACs
in the schedule JSON.  `tests/golden/pixart_schedules.json.gz` holds those vectors (generator:
tests/golden/make_schedule_fixtures.py).  A per-step MAC count pins which sub-blocks were executed, including the
cache-empty fallback and the TGATE rule.  Bit-exact integers.
"""
import numpy as np
import pytest

from ecad_b200.macs import PixArtShape, flops_per_image, macs_per_step
from ecad_b200.schedule import trace_decisions
from golden_util import flags_of, row_by_path, rows, schedule_of


def _gate(row):
    pipe = (row["config"] or {}).get("pipeline") or {}
    return pipe.get("kwargs", {}).get("gate_step") if pipe.get("name") == "tgate" else None


def test_fixture_inventory():
    rs = rows()
    assert len(rs) == 1488
    assert sum(r["macs"] is not None for r in rs) == 1387
    assert sum((r["config"] or {}).get("pipeline", {}).get("name") == "tgate" for r in rs if r["macs"]) == 224


def test_all_golden_mac_vectors_bit_exact():
    bad = []
    n = 0
    for r in rows():
        if r["macs"] is None:
            continue
        n += 1
        g = _gate(r)
        shape = PixArtShape(tokens=r["tokens"], additional_conditions=(r["tokens"] == 4096))
        ex = trace_decisions(flags_of(r), g)
        m = macs_per_step(ex, shape, 2, g)
        if list(map(int, m)) != r["macs"] or int(m.sum()) != r["total_macs"]:
            bad.append(r["path"])
    assert n == 1387
    assert not bad, bad[:5]


def test_headline_numbers():
    r = row_by_path("schedules_in_paper/pixart_alpha_256/ours_fast.json")
    ex = trace_decisions(flags_of(r))
    # 221 attn1 + 160 attn2 + 218 ff executed of 560 each
    assert ex.sum(axis=(0, 1)).tolist() == [221, 160, 218]
    assert r["macs"][0] == 285_689_806_848  # step 0 is a full step despite False flags (cache-empty fallback)
    assert r["macs"][3] == 2 * PixArtShape().macs_fixed() == 1_498_447_872  # everything reused
    assert abs(flops_per_image(ex, PixArtShape()) / 1e12 - 4.4488) < 1e-4
    dense = trace_decisions(np.ones((20, 28, 3), bool))
    assert abs(flops_per_image(dense, PixArtShape()) / 1e12 - 11.924) < 1e-3


def test_cache_empty_fallback_in_trace():
    r = row_by_path("schedules_in_paper/pixart_alpha_256/ours_fast.json")
    flags = flags_of(r)
    assert not flags[0].all()  # the schedule asks for reuse at step 0 ...
    assert trace_decisions(flags)[0].all()  # ... but nothing is cached yet, so everything runs


def test_tgate_trace_rule():
    r = row_by_path("alpha_cache_schedules/gen_tgate/tgate_m_010_sp_001_fi_001_warmup_002.json")
    g = _gate(r)
    assert g == 10
    ex = trace_decisions(flags_of(r), g)
    assert ex[g:, :, 1].sum() == 0  # attn2 never runs from the gate step on
    sched = schedule_of(r)
    assert sched.gate_step() == 10
    assert sched.attn_kinds(0)[0] == ("compute_attn_tgate", {"gate_step": 10})


def test_product_registry_rules_match_trace():
    """The per-sub-block registry policies (what the transformer evaluates each step) reproduce trace_decisions."""
    from ecad_b200.registry import ComputeAttnRegistry, ComputeFFRegistry, DecisionContext

    for path in ["schedules_in_paper/pixart_alpha_256/ours_faster.json",
                 "alpha_cache_schedules/gen_tgate/tgate_m_015_sp_001_fi_001_warmup_002.json"]:
        r = row_by_path(path)
        sched = schedule_of(r)
        g = _gate(r)
        want = trace_decisions(flags_of(r), g)
        has = np.zeros((28, 3), bool)
        for s in range(20):
            got = np.zeros((28, 3), np.uint8)
            for b in range(28):
                acfg = sched.get_custom_compute_attn(str(b))
                fn = ComputeAttnRegistry.get(acfg.get("name"), False)
                for c, comp in enumerate(("attn1", "attn2")):
                    got[b, c] = fn(DecisionContext(b, comp, sched.get_recompute(str(b), comp), not has[b, c], s,
                                                   dict(acfg.get("kwargs", {}))))
                got[b, 2] = ComputeFFRegistry.get(None, False)(
                    DecisionContext(b, "ff", sched.get_recompute(str(b), "ff"), not has[b, 2], s, {}))
            assert np.array_equal(got, want[s]), (path, s)
            has |= got.astype(bool)
            sched.per_step_callback(s, 0)
