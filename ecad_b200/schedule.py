"""Cache schedules: which sub-block of which transformer block is recomputed at which denoising step.

Host-side mirror of the reference's schedule objects for the hot path (same names, argument meaning, JSON
format and error behaviour), re-laid-out for the B200 step executor: the flags live in one dense
``bool[S][NB][3]`` array so that a whole step's decision row is a single 84-byte slice handed to the C-ABI
(`ecadk_pixart_blocks`), instead of three dict look-ups per sub-block.

Reference interfaces mirrored (all under /root/reference/):
  * ``CacheSchedule``            ecad/schedulers/cache_scheduler/cache_schedule.py:18-112
  * ``PixArtCacheSchedule``      ecad/schedulers/cache_scheduler/pixart_cache_schedule.py:9-37
  * schedule JSON TypedDicts     ecad/types.py:43-64
  * genome <-> schedule dict     ecad/genetic/pixart_population_io_manager.py:213-240
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Sequence, Iterable

import numpy as np

PIXART_COMPONENTS = ("attn1", "attn2", "ff")

# Names registered by the reference's compute registries (ecad/transformer_blocks/custom_attn_ff.py:52-59,
# ecad/transformer_blocks/cached_transformer_block.py:326,362,393).
ATTN_CACHED = "compute_attn_cached"
ATTN_TGATE = "compute_attn_tgate"
FF_CACHED = "compute_ff_cached"


class CacheSchedule:
    """``schedule[step][block][component] -> bool`` plus the step counter the pipeline callback advances.

    Mirrors ecad/schedulers/cache_scheduler/cache_schedule.py:18-112.  ``schedule`` accepts the same
    dict-of-dict-of-dict the reference JSON holds (step keys may be ``"000"`` strings or ints).
    """

    components: tuple[str, ...] = ()

    def __init__(
        self,
        num_blocks: int,
        num_inference_steps: int,
        name: str,
        schedule: dict[int | str, dict[str, dict[str, Any]]],
        top_level_config: dict[str, Any] | None = None,
        attributes: dict[str, Any] | None = None,
        metrics: dict[str, Any] | None = None,
        **kwargs: Any,
    ) -> None:
        self.num_blocks = int(num_blocks)
        self.num_inference_steps = int(num_inference_steps)
        self.name = name
        self.metrics = metrics if metrics is not None else {}
        self.attributes = attributes if attributes is not None else {}
        self.top_level_config = top_level_config if top_level_config is not None else {}
        # step keys are cast to int on load (cache_schedule.py:38-41)
        self.schedule: dict[int, dict[str, dict[str, Any]]] = {int(s): blocks for s, blocks in schedule.items()}
        self._last_step = -1
        self._flags: np.ndarray | None = None

    # ---- step counter (cache_schedule.py:58-66) -------------------------------------------------
    def reset_step(self) -> None:
        self._last_step = -1

    @property
    def curr_step(self) -> int:
        return self._last_step + 1

    def per_step_callback(self, step: int, timestep: Any = None, **kwargs: Any) -> None:
        self._last_step = int(step)

    # ---- lookups ---------------------------------------------------------------------------------
    def block_keys(self) -> list[str]:
        raise NotImplementedError

    def get_recompute(self, block_num: str, component: str) -> bool:
        """cache_schedule.py:68-73 - ValueError on an unknown component, KeyError on an unknown step/block."""
        if component not in self.components:
            raise ValueError(f"Invalid component {component}. Must be one of {list(self.components)}.")
        return bool(self.schedule[self.curr_step][block_num][component])

    # ---- dense view ------------------------------------------------------------------------------
    def to_numpy(self, flatten: bool = False) -> np.ndarray:
        raise NotImplementedError

    @property
    def flags(self) -> np.ndarray:
        """Dense read-only ``bool[S][NB][C]`` view, built once (the executor slices one row per step)."""
        if self._flags is None:
            arr = self.to_numpy()
            arr.setflags(write=False)
            self._flags = arr
        return self._flags

    def content_key(self) -> str:
        """Digest of everything that decides what a generation executes (flags, custom decision functions and their
        kwargs, the pipeline/config block).  Two schedules with the same key replay the same recorded work, whatever
        their names or object identities are - the key of the whole-generation CUDA-graph cache."""
        import hashlib

        h = hashlib.sha1()
        h.update(type(self).__name__.encode())
        h.update(repr((self.num_blocks, self.num_inference_steps, getattr(self, "num_single_blocks", None))).encode())
        h.update(json.dumps(self.schedule, sort_keys=True, default=str).encode())
        h.update(json.dumps(self.top_level_config, sort_keys=True, default=str).encode())
        return h.hexdigest()

    # ---- JSON (cache_schedule.py:75-112) ----------------------------------------------------------
    def to_dict(self) -> dict[str, Any]:
        return {
            "cache_schedule": {
                "num_blocks": self.num_blocks,
                "num_inference_steps": self.num_inference_steps,
                "name": self.name,
                "attributes": self.attributes,
                "schedule": {f"{step:03}": blocks for step, blocks in self.schedule.items()},
            },
            "config": self.top_level_config,
            "metrics": self.metrics,
        }

    def to_json(self, file_path: Path) -> None:
        with Path(file_path).open("w") as f:
            json.dump(self.to_dict(), f, indent=4)

    @classmethod
    def from_dict(cls, data: dict[str, Any]) -> "CacheSchedule":
        data = dict(data)
        top_level_config = data.pop("config", None)
        metrics = data.pop("metrics", None)
        # a file without a "cache_schedule" key raises KeyError, which the image generator catches to fall
        # back to the default schedule (ecad/image_generators/image_generator.py:119-125)
        return cls(**data["cache_schedule"], top_level_config=top_level_config, metrics=metrics)

    @classmethod
    def from_json(cls, file_path: Path) -> "CacheSchedule":
        with Path(file_path).open("r") as f:
            return cls.from_dict(json.load(f))


class PixArtCacheSchedule(CacheSchedule):
    """PixArt-alpha/sigma schedule: 28 blocks x (attn1, attn2, ff).  pixart_cache_schedule.py:9-37."""

    components = PIXART_COMPONENTS

    def block_keys(self) -> list[str]:
        return [str(b) for b in range(self.num_blocks)]

    def to_numpy(self, flatten: bool = False) -> np.ndarray:
        arr = np.zeros((self.num_inference_steps, self.num_blocks, 3), dtype=np.bool_)
        for step, blocks in self.schedule.items():
            for block_num, comp in blocks.items():
                b = int(block_num)
                arr[step, b, 0] = comp["attn1"]
                arr[step, b, 1] = comp["attn2"]
                arr[step, b, 2] = comp["ff"]
        return arr.reshape(-1) if flatten else arr

    def get_custom_compute_attn(self, block_num: str) -> dict[str, Any]:
        return self.schedule[self.curr_step][block_num].get("custom_compute_attn", {})

    def get_custom_compute_ff(self, block_num: str) -> dict[str, Any]:
        return self.schedule[self.curr_step][block_num].get("custom_compute_ff", {})

    # ---- constructors the GA side uses ------------------------------------------------------------
    @classmethod
    def from_numpy(
        cls,
        flags: np.ndarray | Iterable[bool],
        num_inference_steps: int = 20,
        num_blocks: int = 28,
        name: str = "from_numpy",
        custom_compute_attn: dict[str, Any] | None = None,
        top_level_config: dict[str, Any] | None = None,
        attributes: dict[str, Any] | None = None,
        metrics: dict[str, Any] | None = None,
    ) -> "PixArtCacheSchedule":
        """Genome -> schedule; inverse of ``to_numpy`` (pixart_population_io_manager.py:213-240)."""
        arr = np.asarray(flags).astype(np.bool_).reshape(num_inference_steps, num_blocks, 3)
        sched: dict[int, dict[str, dict[str, Any]]] = {}
        for s in range(num_inference_steps):
            blocks: dict[str, dict[str, Any]] = {}
            for b in range(num_blocks):
                entry: dict[str, Any] = {
                    "attn1": bool(arr[s, b, 0]),
                    "attn2": bool(arr[s, b, 1]),
                    "ff": bool(arr[s, b, 2]),
                }
                if custom_compute_attn:
                    entry["custom_compute_attn"] = json.loads(json.dumps(custom_compute_attn))
                blocks[str(b)] = entry
            sched[s] = blocks
        return cls(num_blocks, num_inference_steps, name, sched, top_level_config, attributes, metrics)

    @classmethod
    def default(cls, num_inference_steps: int = 20, num_blocks: int = 28) -> "PixArtCacheSchedule":
        """All-true schedule == the uncached model (schedules/alpha_cache_schedules/gen_default/default.json)."""
        return cls.from_numpy(
            np.ones((num_inference_steps, num_blocks, 3), np.bool_), num_inference_steps, num_blocks, "default"
        )

    # ---- per-step rows for the executor -----------------------------------------------------------
    def attn_kinds(self, step: int) -> list[tuple[str, dict[str, Any]]]:
        """Per block: (registry name lower-cased or default, kwargs) for `custom_compute_attn` at `step`."""
        out = []
        for b in self.block_keys():
            cfg = self.schedule[step][b].get("custom_compute_attn", {}) or {}
            name = cfg.get("name")
            out.append(((name or ATTN_CACHED).lower(), dict(cfg.get("kwargs", {}) or {})))
        return out

    def gate_step(self) -> int | None:
        """TGATE gate step of the PIPELINE (``config.pipeline``, ecad/pipelines/tgate.py:329-341): from this step on
        the loop drops the CFG pair.  The per-block decision rule is `block_gate_steps`."""
        pipe = (self.top_level_config or {}).get("pipeline") or {}
        if pipe.get("name") == "tgate":
            g = (pipe.get("kwargs") or {}).get("gate_step")
            if g is not None:
                return int(g)
        return None

    def block_gate_steps(self) -> np.ndarray:
        """``int32[S][NB]``: the ``gate_step`` kwarg of every (step, block) whose ``custom_compute_attn`` resolves to
        ``compute_attn_tgate`` (cached_transformer_block.py:393-454), -1 elsewhere - what the runtime decides from.
        A block entry that names TGATE without a ``gate_step`` raises like the reference (:438-440)."""
        out = np.full((self.num_inference_steps, self.num_blocks), -1, dtype=np.int32)
        for step, blocks in self.schedule.items():
            for key, entry in blocks.items():
                cfg = entry.get("custom_compute_attn") or {}
                if (cfg.get("name") or "").lower() == ATTN_TGATE:
                    g = (cfg.get("kwargs") or {}).get("gate_step")
                    if g is None:
                        raise ValueError("gate_step must be provided as a kwarg to commpute_attn_tgate.")
                    out[step, int(key)] = int(g)
        return out


def trace_decisions(
    flags: np.ndarray,
    attn2_tgate_gate_step: int | np.ndarray | None = None,
) -> np.ndarray:
    """Executed/reused decision of every (step, block, component) over one generation.

    Rule (ecad/transformer_blocks/cached_transformer_block.py:340-347,367-373): a sub-block is executed iff
    ``flag or cache_is_None``; caches start empty (reset at the end of the previous generation,
    ecad/image_generators/image_generator.py:197-202) and are filled by the first pass through a sub-block.
    TGATE attn2 (``compute_attn_tgate``, :393-454): the flag rule applies while ``curr_step <= gate_step-1``;
    from ``gate_step`` on attn2 is never executed (the reference asserts the cache exists).
    ``attn2_tgate_gate_step``: one gate step for every block, or ``int[S][NB]`` per (step, block) with -1 = the block
    decides through ``compute_attn_cached`` (`PixArtCacheSchedule.block_gate_steps`).

    Returns ``uint8[S][NB][3]``: 1 = execute, 0 = reuse the cached tensor.
    """
    flags = np.asarray(flags, dtype=np.bool_)
    S, NB, C = flags.shape
    if attn2_tgate_gate_step is None:
        gates = np.full((S, NB), -1, dtype=np.int64)
    elif np.ndim(attn2_tgate_gate_step) == 0:
        gates = np.full((S, NB), int(attn2_tgate_gate_step), dtype=np.int64)
    else:
        gates = np.asarray(attn2_tgate_gate_step, dtype=np.int64).reshape(S, NB)
    executed = np.zeros((S, NB, C), dtype=np.uint8)
    have_cache = np.zeros((NB, C), dtype=np.bool_)
    for s in range(S):
        run = flags[s] | ~have_cache
        gated = (gates[s] >= 0) & (s >= gates[s])
        if gated.any():
            if not have_cache[gated, 1].all():
                raise AssertionError("Cross-Attention must be cached at gate step for TGATE.")
            run[gated, 1] = False
        executed[s] = run
        have_cache |= run
    return executed


def load_packed_schedules(path: Path | str) -> list[dict[str, Any]]:
    """Rows of a packed schedule collection (format: tests/golden/make_schedule_fixtures.py)."""
    import gzip

    with gzip.open(path, "rb") as f:
        return json.loads(f.read())["rows"]


def schedule_from_packed(row: dict[str, Any]) -> PixArtCacheSchedule:
    """Rebuild a PixArtCacheSchedule (reference JSON semantics) from one packed row."""
    S, NB = row["S"], row["NB"]
    bits = np.unpackbits(np.frombuffer(bytes.fromhex(row["bits"]), np.uint8))[: S * NB * 3]
    custom = None
    if row.get("custom"):
        custom = {"name": row["custom"]["attn"], "kwargs": {"gate_step": row["custom"]["gate_step"]}}
    return PixArtCacheSchedule.from_numpy(
        bits.reshape(S, NB, 3).astype(np.bool_), S, NB, row["name"], custom_compute_attn=custom,
        top_level_config=row.get("config") or {}, attributes=row.get("attributes") or {},
    )


FLUX_FULL_COMPONENTS = ("full_attn", "full_ff", "full_ff_context")
FLUX_SINGLE_COMPONENTS = ("single_attn", "single_proj_mlp", "single_proj_out")


class FluxCacheSchedule(CacheSchedule):
    """FLUX.1 schedule: 19 double-stream blocks x (full_attn, full_ff, full_ff_context) and 38 single-stream blocks x
    (single_attn, single_proj_mlp, single_proj_out); block keys "0".."18" and "single_0".."single_37".
    Mirrors ecad/schedulers/cache_scheduler/flux_cache_schedule.py:13-90 (same constructor contract: a missing
    ``num_single_blocks`` raises ValueError; ``to_numpy`` only supports ``flatten=True``)."""

    components = FLUX_SINGLE_COMPONENTS + FLUX_FULL_COMPONENTS

    def __init__(self, num_blocks, num_inference_steps, name, schedule, top_level_config=None, attributes=None,
                 metrics=None, num_single_blocks: int | None = None, **kwargs: Any) -> None:
        super().__init__(num_blocks, num_inference_steps, name, schedule, top_level_config, attributes, metrics)
        if num_single_blocks is None:
            raise ValueError("num_single_blocks must be provided for FluxCacheSchedule")
        self.num_single_blocks = int(num_single_blocks)

    def block_keys(self) -> list[str]:
        return [str(b) for b in range(self.num_blocks)] + [f"single_{b}" for b in range(self.num_single_blocks)]

    def to_dict(self) -> dict[str, Any]:
        data = super().to_dict()
        data["cache_schedule"]["num_single_blocks"] = self.num_single_blocks
        return data

    def dense(self) -> np.ndarray:
        """``bool[S][NB + NS][3]``: double blocks first, then single blocks, reference component order per kind."""
        arr = np.zeros((self.num_inference_steps, self.num_blocks + self.num_single_blocks, 3), dtype=np.bool_)
        for step, blocks in self.schedule.items():
            for key, comp in blocks.items():
                if key.startswith("single_"):
                    r, names = self.num_blocks + int(key[len("single_"):]), FLUX_SINGLE_COMPONENTS
                else:
                    r, names = int(key), FLUX_FULL_COMPONENTS
                for i, c in enumerate(names):
                    arr[step, r, i] = comp[c]
        return arr

    def to_numpy(self, flatten: bool = True) -> np.ndarray:
        """Genome: per step all double-block flags then all single-block flags (flux_cache_schedule.py:62-90)."""
        if not flatten:
            raise NotImplementedError("FluxCacheSchedule only supports flatten=True")
        return self.dense().reshape(-1)

    @classmethod
    def from_numpy(cls, flags, num_inference_steps: int = 20, num_blocks: int = 19, num_single_blocks: int = 38,
                   name: str = "from_numpy", top_level_config=None, attributes=None, metrics=None):
        arr = np.asarray(flags).astype(np.bool_).reshape(num_inference_steps, num_blocks + num_single_blocks, 3)
        sched: dict[int, dict[str, dict[str, bool]]] = {}
        for s in range(num_inference_steps):
            blocks: dict[str, dict[str, bool]] = {}
            for b in range(num_blocks):
                blocks[str(b)] = {c: bool(arr[s, b, i]) for i, c in enumerate(FLUX_FULL_COMPONENTS)}
            for b in range(num_single_blocks):
                blocks[f"single_{b}"] = {c: bool(arr[s, num_blocks + b, i])
                                         for i, c in enumerate(FLUX_SINGLE_COMPONENTS)}
            sched[s] = blocks
        return cls(num_blocks, num_inference_steps, name, sched, top_level_config, attributes, metrics,
                   num_single_blocks=num_single_blocks)


# ---- dead cache stores ------------------------------------------------------------------------------------------
def pixart_dead_store_mask(schedule: "PixArtCacheSchedule", step: int, executed: np.ndarray,
                           keep_attn2_blocks: Sequence[int] = ()) -> np.ndarray:
    """``uint8[NB][3]``: 1 where an EXECUTED sub-block's cache store at ``step`` is dead - the slot is overwritten or
    dropped before anything reads it - so the GEMM epilogue may skip the reference's ``self.cached_* = out``
    (cached_transformer_block.py:357-358,388-389).  Rule (one step of look-ahead, conservative):

      * last step of the generation: every slot is dropped by the reset callback (image_generator.py:193-202);
      * otherwise the next step's flag recomputes the sub-block AND both steps decide through the default functions
        (``compute_attn_cached`` / ``compute_ff_cached``): then ``flag or cache is None`` executes it whatever the
        cache holds.  Custom / TGATE decision functions keep their stores (their next decision is not a pure flag),
        and so do the blocks whose attn2 cache is averaged after this forward (``keep_attn2_blocks``).
    """
    from .registry import ComputeAttnRegistry, ComputeFFRegistry

    nb = schedule.num_blocks
    dead = np.zeros((nb, 3), dtype=np.uint8)
    last = step >= schedule.num_inference_steps - 1
    nxt = None if last else schedule.schedule.get(step + 1)
    if not last and nxt is None:
        return dead
    default_attn, default_ff = ComputeAttnRegistry.default(), ComputeFFRegistry.default()

    def is_default(entry) -> tuple[bool, bool]:
        an = (entry.get("custom_compute_attn") or {}).get("name")
        fn = (entry.get("custom_compute_ff") or {}).get("name")
        a = ComputeAttnRegistry.get(an, False)
        f = ComputeFFRegistry.get(fn, False)
        # a user-registered tensor function may read any slot of its block whatever the flags say
        tensor = ComputeAttnRegistry.get_tensor(an) is not None or ComputeFFRegistry.get_tensor(fn) is not None
        return a is default_attn and not tensor, f is default_ff and not tensor

    row = schedule.schedule[step]
    keep = set(keep_attn2_blocks)
    for b in range(nb):
        cur_a, cur_f = is_default(row[str(b)])
        if last:
            nxt_a, nxt_f, flags = True, True, (True, True, True)
        else:
            e = nxt[str(b)]
            nxt_a, nxt_f = is_default(e)
            flags = (bool(e["attn1"]), bool(e["attn2"]), bool(e["ff"]))
        ok = (cur_a and nxt_a, cur_a and nxt_a and b not in keep, cur_f and nxt_f)
        for c in range(3):
            dead[b, c] = bool(executed[b, c]) and ok[c] and flags[c]
    return dead


def flux_dead_store_mask(schedule: "FluxCacheSchedule", step: int, executed: np.ndarray) -> np.ndarray:
    """FLUX counterpart (``uint8[NB + NS][3]``, FluxCacheSchedule.dense() order).  single_attn / single_proj_mlp are
    produced straight into their cache slots by the kernels and are never skipped."""
    rows = schedule.num_blocks + schedule.num_single_blocks
    dead = np.zeros((rows, 3), dtype=np.uint8)
    last = step >= schedule.num_inference_steps - 1
    if last:
        nxt_flags = np.ones((rows, 3), dtype=np.bool_)
    else:
        nxt = schedule.schedule.get(step + 1)
        if nxt is None:
            return dead
        nxt_flags = np.zeros((rows, 3), dtype=np.bool_)
        for r in range(rows):
            if r < schedule.num_blocks:
                e = nxt[str(r)]
                nxt_flags[r] = [e[c] for c in FLUX_FULL_COMPONENTS]
            else:
                e = nxt[f"single_{r - schedule.num_blocks}"]
                nxt_flags[r] = [e[c] for c in FLUX_SINGLE_COMPONENTS]
    dead[:] = np.asarray(executed).astype(np.bool_) & nxt_flags
    dead[schedule.num_blocks:, 0:2] = 0
    return dead
