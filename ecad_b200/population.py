"""Population evaluation sharded over the GPUs of one box (BASELINE config 2).

The NSGA-II driver of the reference evaluates a population of candidate schedules by shelling out to
``ecad/benchmark/generate_images.py`` once per generation, which rebuilds the pipeline and reloads all weights for
every candidate (/root/reference/ecad/genetic/train_nsga2_single_gpu.py:131-158,198-224;
ecad/benchmark/generate_images.py:48-63).  Here one resident model per GPU serves every candidate assigned to it.

A unit of work is (candidate schedule, prompt batch); units share nothing but read-only weights, so the path shards
with NO data-path collective.  One process per GPU (torchrun); the only communication is the final gather of latents
(16 KB / image) and per-candidate metrics to the search driver on rank 0 - NCCL over NVLink on GPUs, gloo in the
CPU tests.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist


def partition_lpt(costs: Sequence[float], world_size: int) -> list[list[int]]:
    """Longest-processing-time-first assignment of units to ranks (deterministic; ties by index).

    ``costs`` = analytic FLOPs of each candidate (ecad_b200.macs.flops_per_image) - cheap schedules and expensive
    ones differ by >5x, so a round-robin split would leave GPUs idle."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * world_size
    parts: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += float(costs[i])
    return [sorted(p) for p in parts]


def partition_round_robin(n_units: int, world_size: int) -> list[list[int]]:
    return [list(range(r, n_units, world_size)) for r in range(world_size)]


class PopulationEvaluator:
    """Runs ``run_unit(i) -> Tensor`` for the units of this rank and gathers all results on every rank.

    ``run_unit`` returns a tensor of a fixed shape (e.g. final latents ``[B, 4, h, w]``) on ``device``.
    """

    def __init__(self, rank: int = 0, world_size: int = 1, device: torch.device | str = "cpu",
                 group: dist.ProcessGroup | None = None):
        self.rank, self.world_size = rank, world_size
        self.device = torch.device(device)
        self.group = group
        # persistent staging buffers of run_from_host: page-locking host memory (cudaHostAlloc) and cudaMalloc are slow
        # and can stall the launching thread behind the running GPU work, so they must not happen per unit
        self._pinned_pool: dict[tuple, list[torch.Tensor]] = {}
        self._dev_inputs: dict[tuple, list[dict]] = {}
        if world_size > 1 and not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised for world_size > 1")

    def evaluate(self, n_units: int, run_unit: Callable[[int], torch.Tensor], costs: Sequence[float] | None = None,
                 gather: bool = True) -> dict:
        parts = (partition_lpt(costs, self.world_size) if costs is not None
                 else partition_round_robin(n_units, self.world_size))
        mine = parts[self.rank]
        timed = self.device.type == "cuda"
        if timed:  # device-side clock of this rank's own work and of the whole call (incl. the gather)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        local = [run_unit(i) for i in mine]
        if timed:
            ev[1].record()
        out = {"assignment": parts, "local_indices": mine, "local": local}
        if gather:
            out["results"] = self.gather(local, parts, n_units)
        if timed:
            ev[2].record()
            torch.cuda.synchronize(self.device)
            out["busy_s"] = ev[0].elapsed_time(ev[1]) * 1e-3
            out["total_s"] = ev[0].elapsed_time(ev[2]) * 1e-3
        if costs is not None:
            loads = [sum(float(costs[i]) for i in p) for p in parts]
            out["planned_load"] = loads
            # what the cost model predicts for this split: mean load / max load (1.0 = perfectly balanced)
            out["planned_efficiency"] = (sum(loads) / len(loads)) / max(loads) if max(loads) > 0 else 1.0
        return out

    def run_from_host(self, unit_indices: Sequence[int], run_unit: Callable[[int, dict], torch.Tensor],
                      host_inputs: Callable[[int], dict] | dict, keep_device_results: bool = True) -> dict:
        """Runs this rank's units with their inputs coming from (pinned) HOST memory and their results going back to
        pinned host memory, software-pipelined on a copy stream: the host->device copy of unit k+1 and the
        device->host copy of unit k-1 overlap the generation of unit k (two device input buffers).

        ``host_inputs`` is a dict of host tensors used for every unit (the ECAD search evaluates every candidate on
        the same prompt set) or a callable ``i -> dict``.  ``run_unit(i, device_inputs)`` returns the unit's result
        on the device.  Returns ``{"host": [pinned tensors], "device": [device tensors]}`` in unit order.  The pinned
        result buffers and the two device input buffers belong to the evaluator and are REUSED by the next call:
        consume (or copy) the host results before calling again."""
        if self.device.type != "cuda":
            raise RuntimeError("run_from_host pipelines CUDA copies; use evaluate() on CPU")
        get = host_inputs if callable(host_inputs) else (lambda _i: host_inputs)
        cur = torch.cuda.current_stream(self.device)
        copy = torch.cuda.Stream(self.device)
        copy.wait_stream(cur)  # the staging buffers are reused: earlier work on this stream may still read them
        units = list(unit_indices)
        sig = tuple((n, tuple(t.shape), t.dtype) for n, t in get(units[0]).items()) if units else ()
        bufs: list[dict | None] = self._dev_inputs.setdefault(sig, [None, None])
        taken: dict[tuple, int] = {}
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        host_out: list[torch.Tensor] = []
        dev_out: list[torch.Tensor] = []

        def upload(k: int) -> None:
            slot = k & 1
            src = get(units[k])
            with torch.cuda.stream(copy):
                if k >= 2:
                    copy.wait_event(consumed[slot])  # unit k-2 has finished reading this buffer
                if bufs[slot] is None:
                    bufs[slot] = {n: torch.empty(t.shape, dtype=t.dtype, device=self.device) for n, t in src.items()}
                for n, t in src.items():
                    bufs[slot][n].copy_(t, non_blocking=True)
                ready[slot].record(copy)

        if units:
            upload(0)
        for k, i in enumerate(units):
            slot = k & 1
            if k + 1 < len(units):
                upload(k + 1)
            cur.wait_event(ready[slot])
            res = run_unit(i, bufs[slot])
            consumed[slot].record(cur)
            done = torch.cuda.Event()
            done.record(cur)
            pkey = (tuple(res.shape), res.dtype)
            pool = self._pinned_pool.setdefault(pkey, [])
            j = taken.get(pkey, 0)
            taken[pkey] = j + 1
            if j >= len(pool):
                pool.append(torch.empty(res.shape, dtype=res.dtype, pin_memory=True))
            pinned = pool[j]
            with torch.cuda.stream(copy):
                copy.wait_event(done)
                pinned.copy_(res, non_blocking=True)
            res.record_stream(copy)
            host_out.append(pinned)
            if keep_device_results:
                dev_out.append(res)
        copy.synchronize()
        return {"host": host_out, "device": dev_out}

    def gather(self, local: list[torch.Tensor], parts: list[list[int]], n_units: int) -> list[torch.Tensor | None]:
        """All ranks receive every unit's result, ordered by unit index."""
        if self.world_size == 1:
            res: list[torch.Tensor | None] = [None] * n_units
            for i, t in zip(parts[0], local):
                res[i] = t
            return res
        max_units = max(len(p) for p in parts)
        shape_t = torch.zeros(8, dtype=torch.int64, device=self.device)
        if local:
            shp = list(local[0].shape)
            shape_t[0] = len(shp)
            shape_t[1:1 + len(shp)] = torch.tensor(shp, dtype=torch.int64)
        # ranks with no unit learn the unit shape from the others
        dist.all_reduce(shape_t, op=dist.ReduceOp.MAX, group=self.group)
        shp = [int(v) for v in shape_t[1:1 + int(shape_t[0])]]
        dtype = local[0].dtype if local else torch.float32
        buf = torch.zeros(max_units, *shp, dtype=dtype, device=self.device)
        for j, t in enumerate(local):
            buf[j].copy_(t)
        allbuf = torch.empty(self.world_size, max_units, *shp, dtype=dtype, device=self.device)
        dist.all_gather_into_tensor(allbuf.view(-1, *shp), buf, group=self.group)
        res = [None] * n_units
        for r, p in enumerate(parts):
            for j, i in enumerate(p):
                res[i] = allbuf[r, j]
        return res
