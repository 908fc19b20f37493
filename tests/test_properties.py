"""Property tests of the host logic on RANDOM genomes (the NSGA-II search hands the path arbitrary flag arrays, not
only the shipped schedules the golden tests cover): genome <-> schedule round trips
(ecad/genetic/pixart_population_io_manager.py:213-240), the `flag or cache is None` decision rule
(ecad/transformer_blocks/cached_transformer_block.py:340-347,367-373), safety of the dead-cache-store elimination,
and the longest-processing-time-first partition of candidates over ranks."""
import numpy as np
from hypothesis import given, settings, strategies as st

from ecad_b200.population import partition_lpt, partition_round_robin
from ecad_b200.schedule import (FluxCacheSchedule, PixArtCacheSchedule, flux_dead_store_mask, pixart_dead_store_mask,
                                trace_decisions)


@st.composite
def genomes(draw, max_steps=6, max_blocks=5):
    s = draw(st.integers(1, max_steps))
    nb = draw(st.integers(1, max_blocks))
    bits = draw(st.lists(st.booleans(), min_size=s * nb * 3, max_size=s * nb * 3))
    return np.asarray(bits, dtype=bool).reshape(s, nb, 3)


@settings(max_examples=60, deadline=None)
@given(genomes())
def test_genome_round_trips(flags):
    s, nb, _ = flags.shape
    sched = PixArtCacheSchedule.from_numpy(flags.reshape(-1), s, nb)
    assert np.array_equal(sched.to_numpy(), flags)
    again = PixArtCacheSchedule.from_dict(sched.to_dict())
    assert np.array_equal(again.to_numpy(), flags)
    assert again.content_key() == sched.content_key()
    flipped = flags.copy()
    flipped[-1, -1, -1] ^= True
    assert PixArtCacheSchedule.from_numpy(flipped, s, nb).content_key() != sched.content_key()


@settings(max_examples=60, deadline=None)
@given(genomes())
def test_decision_rule(flags):
    ex = trace_decisions(flags).astype(bool)
    assert ex[0].all()                      # caches start empty: the first pass executes everything
    assert (ex | ~flags).all()              # a raised flag always executes
    assert np.array_equal(ex[1:], flags[1:])  # ... and after the first pass the flag alone decides
    # idempotent: the executed trace is its own trace (what `metrics.by_inference_step` is computed from)
    assert np.array_equal(trace_decisions(ex), ex.astype(np.uint8))


@settings(max_examples=60, deadline=None)
@given(genomes())
def test_pixart_dead_stores_are_safe_on_random_genomes(flags):
    s, nb, _ = flags.shape
    sched = PixArtCacheSchedule.from_numpy(flags, s, nb)
    ex = trace_decisions(flags).astype(bool)
    written = np.zeros((nb, 3), bool)
    for step in range(s):
        dead = pixart_dead_store_mask(sched, step, ex[step]).astype(bool)
        assert not (dead & ~ex[step]).any()
        assert written[~ex[step]].all()     # every reuse reads a slot whose latest store was kept
        written = np.where(ex[step], ~dead, written)
        if step + 1 < s:                    # the rule is exact, not only safe: kept <=> the next step reuses it
            assert np.array_equal(~dead[ex[step]], ~flags[step + 1][ex[step]])
    assert pixart_dead_store_mask(sched, s - 1, ex[s - 1]).astype(bool)[ex[s - 1]].all()


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 4), st.integers(1, 3), st.integers(1, 4), st.data())
def test_flux_dead_stores_are_safe_on_random_genomes(s, nb, ns, data):
    bits = data.draw(st.lists(st.booleans(), min_size=s * (nb + ns) * 3, max_size=s * (nb + ns) * 3))
    flags = np.asarray(bits, dtype=bool).reshape(s, nb + ns, 3)
    sched = FluxCacheSchedule.from_numpy(flags, s, nb, ns)
    assert np.array_equal(sched.dense(), flags)
    ex = trace_decisions(flags).astype(bool)
    written = np.zeros((nb + ns, 3), bool)
    for step in range(s):
        dead = flux_dead_store_mask(sched, step, ex[step]).astype(bool)
        assert not (dead & ~ex[step]).any() and not dead[nb:, :2].any()
        assert written[~ex[step]].all()
        written = np.where(ex[step], ~dead, written)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.floats(0.01, 100.0, allow_nan=False), min_size=0, max_size=40), st.integers(1, 8))
def test_lpt_partition(costs, world):
    parts = partition_lpt(costs, world)
    assert len(parts) == world
    flat = sorted(i for p in parts for i in p)
    assert flat == list(range(len(costs)))                 # every candidate exactly once
    assert all(p == sorted(p) for p in parts)
    assert parts == partition_lpt(costs, world)            # deterministic (every rank computes the same split)
    if costs:
        loads = [sum(costs[i] for i in p) for p in parts]
        # Graham's list-scheduling bound: the last unit placed on the fullest rank went to the then-emptiest one
        assert max(loads) <= sum(costs) / world + max(costs) + 1e-9
        rr = partition_round_robin(len(costs), world)
        assert sorted(i for p in rr for i in p) == flat
