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
    # load_image_generator.py:23-40: unknown name -> the default's class, else None; the helper functions raise
    from ecad_b200.registry import get_image_generator_type, get_image_generator_type_from_config
    assert ImageGeneratorRegistry.get("missing") is None
    alpha = ImageGeneratorRegistry.get("b200_pixart_alpha")
    assert ImageGeneratorRegistry.get("missing", "b200_pixart_alpha") is alpha
    assert ImageGeneratorRegistry.get("missing", "also_missing") is None
    # the reference's own class names select the B200 generators (a schedule JSON's config.image_generator)
    assert ImageGeneratorRegistry.get("PixArtAlphaImageGenerator") is alpha
    sigma = ImageGeneratorRegistry.get("PixArtSigmaImageGenerator")
    assert sigma.text_tokens == 300 and sigma.default_pipeline_name == "pixart_sigma" and issubclass(sigma, alpha.__mro__[1])
    assert ImageGeneratorRegistry.get("FluxImageGenerator") is ImageGeneratorRegistry.get("b200_flux")
    with pytest.raises(ValueError):
        get_image_generator_type("missing")  # default "PixArtImageGenerator" is not registered (as in the reference)
    assert get_image_generator_type("missing", "b200_pixart_sigma") is sigma
    assert get_image_generator_type_from_config({"image_generator": "PixArtSigmaImageGenerator"}) is sigma
    with pytest.raises(ValueError):
        get_image_generator_type_from_config({})


def test_block_gate_steps_drive_the_trace():
    """The TGATE attn2 rule is taken per block from custom_compute_attn kwargs (what the runtime decides from), not
    from one pipeline-level value: two blocks with different gate steps, one block without TGATE."""
    S, NB = 6, 3
    flags = np.ones((S, NB, 3), bool)
    sched = PixArtCacheSchedule.from_numpy(flags, S, NB, "mixed")
    for s in range(S):
        sched.schedule[s]["0"]["custom_compute_attn"] = {"name": "compute_attn_tgate", "kwargs": {"gate_step": 2}}
        sched.schedule[s]["1"]["custom_compute_attn"] = {"name": "Compute_Attn_TGATE", "kwargs": {"gate_step": 4}}
    gates = sched.block_gate_steps()
    assert gates.shape == (S, NB) and (gates[:, 0] == 2).all() and (gates[:, 1] == 4).all() and (gates[:, 2] == -1).all()
    assert sched.gate_step() is None  # no pipeline-level entry
    from ecad_b200.metrics import executed_trace
    ex = executed_trace(sched)
    assert ex[:2, 0, 1].all() and not ex[2:, 0, 1].any()
    assert ex[:4, 1, 1].all() and not ex[4:, 1, 1].any()
    assert ex[:, 2, 1].all() and ex[:, :, 0].all() and ex[:, :, 2].all()
    sched.schedule[0]["0"]["custom_compute_attn"] = {"name": "compute_attn_tgate", "kwargs": {}}
    with pytest.raises(ValueError):
        sched.block_gate_steps()


def test_content_key_tracks_content_not_identity():
    a = PixArtCacheSchedule.from_numpy(np.ones((4, 2, 3), bool), 4, 2, "x")
    b = PixArtCacheSchedule.from_numpy(np.ones((4, 2, 3), bool), 4, 2, "y")
    assert a.content_key() == b.content_key()
    f = np.ones((4, 2, 3), bool)
    f[2, 1, 0] = False
    c = PixArtCacheSchedule.from_numpy(f, 4, 2, "x")
    assert c.content_key() != a.content_key()
    d = PixArtCacheSchedule.from_numpy(np.ones((4, 2, 3), bool), 4, 2, "x",
                                       top_level_config={"pipeline": {"name": "tgate", "kwargs": {"gate_step": 2}}})
    assert d.content_key() != a.content_key()


def test_trace_resets_between_generations():
    flags = np.zeros((4, 2, 3), bool)
    ex = trace_decisions(flags)
    assert ex[0].all() and not ex[1:].any()
