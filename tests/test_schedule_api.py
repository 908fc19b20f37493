"""Host-side schedule objects: same behaviour as the reference's CacheSchedule / PixArtCacheSchedule
(ecad/schedulers/cache_scheduler/cache_schedule.py, pixart_cache_schedule.py)."""
import json

import numpy as np
import pytest

from ecad_b200.registry import ComputeAttnRegistry, ComputeFFRegistry, ImageGeneratorRegistry
from ecad_b200.schedule import PixArtCacheSchedule, trace_decisions
from golden_util import flags_of, row_by_path, schedule_of


def test_json_round_trip_reference_format(tmp_path):
    r = row_by_path("schedules_in_paper/pixart_alpha_256/ours_fast.json")
    s = schedule_of(r)
    p = tmp_path / "s.json"
    s.to_json(p)
    d = json.loads(p.read_text())
    # cache_schedule.py:75-89: zero-padded step keys, block keys "0".."27", siblings config/metrics
    assert set(d) == {"cache_schedule", "config", "metrics"}
    assert list(d["cache_schedule"]["schedule"])[:2] == ["000", "001"]
    assert set(d["cache_schedule"]["schedule"]["000"]["0"]) == {"attn1", "attn2", "ff"}
    s2 = PixArtCacheSchedule.from_json(p)
    assert np.array_equal(s2.to_numpy(), flags_of(r))
    assert s2.name == r["name"] and s2.num_blocks == 28 and s2.num_inference_steps == 20
    assert s2.attributes == r["attributes"]


def test_genome_order_and_inverse():
    rng = np.random.default_rng(0)
    g = rng.integers(0, 2, 20 * 28 * 3).astype(bool)
    s = PixArtCacheSchedule.from_numpy(g)
    flat = s.to_numpy(flatten=True)
    assert flat.shape == (1680,) and np.array_equal(flat, g)
    # [step][block][attn1, attn2, ff] C-order (pixart_cache_schedule.py:15-27)
    assert s.schedule[3]["5"]["attn2"] == bool(g[(3 * 28 + 5) * 3 + 1])
    assert s.flags.flags.writeable is False


def test_step_counter_and_lookup_errors():
    s = PixArtCacheSchedule.default()
    assert s.curr_step == 0
    s.per_step_callback(0, 999)
    assert s.curr_step == 1
    s.per_step_callback(7, 500)
    assert s.curr_step == 8
    assert s.get_recompute("3", "ff") is True
    with pytest.raises(ValueError):
        s.get_recompute("3", "attn3")
    with pytest.raises(KeyError):
        s.get_recompute("28", "ff")
    s.per_step_callback(19, 50)
    with pytest.raises(KeyError):  # step 20 does not exist: the generator must reset before the next image
        s.get_recompute("0", "ff")
    s.reset_step()
    assert s.curr_step == 0


def test_from_dict_without_cache_schedule_raises_keyerror():
    with pytest.raises(KeyError):
        PixArtCacheSchedule.from_dict({"dit_schedule": {}})


def test_custom_compute_lookup():
    r = row_by_path("alpha_cache_schedules/gen_tgate/tgate_m_010_sp_001_fi_001_warmup_002.json")
    s = schedule_of(r)
    assert s.get_custom_compute_attn("0") == {"name": "compute_attn_tgate", "kwargs": {"gate_step": 10}}
    assert s.get_custom_compute_ff("0") == {}
    plain = PixArtCacheSchedule.default()
    assert plain.get_custom_compute_attn("0") == {}


def test_registries_mirror_reference_lookup_rules():
    # custom_attn_ff.py:22-35: lower-cased names; unknown -> default unless none_if_not_found
    assert ComputeAttnRegistry.get("COMPUTE_ATTN_TGATE").__name__ == "compute_attn_tgate"
    assert ComputeAttnRegistry.get(None).__name__ == "compute_attn_cached"
    assert ComputeAttnRegistry.get("nope").__name__ == "compute_attn_cached"
    assert ComputeAttnRegistry.get("nope", True) is None
    assert ComputeFFRegistry.get(None).__name__ == "compute_ff_cached"
    import ecad_b200.image_generator  # noqa: F401  (registers the generator)
    assert "b200_pixart_alpha" in ImageGeneratorRegistry.registry
    with pytest.raises(ValueError):
        ImageGeneratorRegistry.get("missing")


def test_trace_resets_between_generations():
    flags = np.zeros((4, 2, 3), bool)
    ex = trace_decisions(flags)
    assert ex[0].all() and not ex[1:].any()
