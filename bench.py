#!/usr/bin/env python
"""bench.py - PixArt-alpha 256x256 cached images/s on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" = one candidate cache schedule evaluated on one batch of 100 prompts: 20 DPM-Solver++ steps, CFG on
(200 samples per forward), random-init PixArt-alpha XL/2 weights, synthetic T5 embeddings -> 100 final latents.
This is BASELINE config 2 (NSGA-II population eval: 72 schedules x 100 prompts); with N GPUs every rank evaluates its
own candidates (weak scaling, no data-path collective) and the final latents are gathered over NCCL.

    python bench.py                                   # N=1
    torchrun --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference                  # the CPU oracle (the reference cannot run without CUDA +
                                                      # diffusers) timed on the host cores, same metric/config

Keys beyond the base contract: `roofline` (dominant kernel = the tcgen05 GEMM: algorithmic FLOPs of every GEMM launch
of the timed region / their CUDA-event durations, recorded on the launch stream inside libecad_b200; plus the same
kernel timed alone and the whole-step tensor fraction), `cpu_baseline` (oracle on the host cores, bounded sample),
`e2e` (through the ImageGenerator API with pinned host inputs and a device->host read of the latents),
`gpu_launches`, `clocks`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

GOLDEN = ROOT / "tests" / "golden" / "pixart_schedules.json.gz"
CANDIDATE_DIR = "population_initialization/pixart_alpha_256x256/gen_000/candidates/"
HEADLINE = "schedules_in_paper/pixart_alpha_256/ours_fast.json"
METRIC = "PixArt-alpha 256x256 20-step cached images/s (NSGA-II population eval, 100 prompts per candidate)"
PROMPTS_PER_STEP = 100
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant GEMM instance (FF up-projection,
# M=51200 N=4608 K=1152) from the committed `ncu --set full` capture profiles/r1_kernels_ncu_full.csv
NCU_TRAFFIC_BYTES_PER_LAUNCH = int((128.69 + 427.45) * 1e6)


def load_candidates():
    """The reference's shipped 72-candidate seed population (packed in tests/golden) + the paper's ours_fast."""
    from ecad_b200.schedule import load_packed_schedules, schedule_from_packed

    rows = load_packed_schedules(GOLDEN)
    cands = sorted((r for r in rows if r["path"].startswith(CANDIDATE_DIR)), key=lambda r: r["path"])
    head = [r for r in rows if r["path"] == HEADLINE]
    assert len(cands) == 72 and len(head) == 1
    return head[0], cands, schedule_from_packed


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_oracle_images_per_s(step_rows, warmup: int):
    """The CPU oracle (fp32, all host threads) on a bounded sample of the SAME workload as the GPU arm: for each step
    the same candidate schedule, 20 DPM steps, CFG on - but ONE prompt (2 samples per forward) instead of 100."""
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings
    from oracle.pixart_oracle import OracleConfig, OracleSchedule, PixArtOracle, generate_latents

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = random_init_state_dict(PixArtConfig(), 0)
    emb = synthetic_prompt_embeddings(1, seed=1)
    times = []
    for it, row in enumerate(step_rows):
        S, NB = row["S"], row["NB"]
        flags = np.unpackbits(np.frombuffer(bytes.fromhex(row["bits"]), np.uint8))[: S * NB * 3].reshape(S, NB, 3)
        model = PixArtOracle(sd, OracleConfig(), OracleSchedule.from_flags(flags.astype(bool)))
        noise = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(it))
        t0 = time.perf_counter()
        generate_latents(model, emb["prompt_embeds"], emb["prompt_attention_mask"], emb["negative_prompt_embeds"],
                         emb["negative_prompt_attention_mask"], noise, S)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, cores


def workload_text(fixed_schedule: bool, prompts: int) -> str:
    return ("PixArt-alpha XL/2 256x256, 20 DPM-Solver++ steps, CFG 4.5, "
            + ("ours_fast schedule" if fixed_schedule else
               "candidates of the reference's gen_000 seed population (candidate i at step i)")
            + f", {prompts} prompts per step (200 samples/forward), random-init weights, synthetic T5")


def run_reference(args):
    """--impl reference: the reference's CPU path.  The reference itself refuses to start without CUDA
    (pixart_image_generator.py:55-56) and needs diffusers (absent, no network), so this arm times the fp32 CPU
    restatement (oracle/), all host threads, one image (batch 1) per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    head_row, cand_rows, _ = load_candidates()
    n = args.warmup + args.steps
    step_rows = [head_row if args.fixed_schedule else cand_rows[i % len(cand_rows)] for i in range(n)]
    times, cores = cpu_oracle_images_per_s(step_rows, args.warmup)
    total = sum(times)
    value = len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.fixed_schedule, args.prompts),
                   "bounded_sample": "1 prompt (2 samples/forward) per step instead of 100, same candidate schedules",
                   "kind": "port (oracle/pixart_oracle.py): the reference needs CUDA + diffusers and cannot run here"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} steps x (1 image, 20 denoising steps, the step's candidate schedule)"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def time_dominant_gemm(device, peaks, iters=20):
    """The FF up-projection (+bias +GELU) of one batch-100 forward, [51200,1152] x [4608,1152]^T, timed alone with
    CUDA events on the launching stream; inputs (118 MB) + output (472 MB) exceed L2 (126 MB)."""
    from ecad_b200 import _lib

    M, N, K = 200 * 256, 4608, 1152
    g = torch.Generator(device=device).manual_seed(0)
    a = torch.randn(M, K, device=device, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=device, generator=g) / K**0.5).to(torch.bfloat16)
    b = torch.randn(N, device=device, generator=g)
    out = torch.empty(M, N, device=device, dtype=torch.bfloat16)
    for _ in range(3):
        _lib.gemm_bias(a, w, b, out, gelu=True)
    torch.cuda.synchronize()
    st = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        st[i].record()
        _lib.gemm_bias(a, w, b, out, gelu=True)
        en[i].record()
    torch.cuda.synchronize()
    ms = statistics.mean(s.elapsed_time(e) for s, e in zip(st, en))
    flops = 2.0 * M * N * K
    achieved = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "gemm_bf16_kernel<256,EPI_BIAS_GELU> M=51200 N=4608 K=1152",
            "achieved": achieved, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_burst"],
            "peak_source": f"{peaks['src']} burst (kernel timed alone)", "ms_per_launch": ms,
            "flops_per_launch": flops, "traffic": None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--prompts", type=int, default=PROMPTS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fixed-schedule", action="store_true", help="every step runs ours_fast instead of a candidate")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    from ecad_b200 import _lib
    from ecad_b200.image_generator import B200PixArtAlphaImageGenerator
    from ecad_b200.macs import PixArtShape, flops_per_image
    from ecad_b200.population import PopulationEvaluator
    from ecad_b200.schedule import trace_decisions
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks = measured_peaks()

    head_row, cand_rows, from_packed = load_candidates()
    W, K, B = args.warmup, args.steps, args.prompts
    total_steps = W + K
    # Weak scaling with EQUAL work per rank: step i evaluates candidate i of the population on every rank, each rank on
    # its own chunk of prompts (units = (candidate schedule, prompt chunk)).
    def row_for(step_idx):
        if args.fixed_schedule:
            return head_row
        return cand_rows[step_idx % len(cand_rows)]

    sd = random_init_state_dict(PixArtConfig(), 0)
    gen = B200PixArtAlphaImageGenerator(cache_schedule=from_packed(head_row), start_seed=1234 + rank, state_dict=sd,
                                        device=f"cuda:{local_rank}")
    gen.create_diffusion_pipeline()
    tr = gen.diffusion_pipeline.transformer
    emb_host = {k: v.pin_memory() for k, v in synthetic_prompt_embeddings(B, seed=1 + rank).items()}
    emb_dev = {k: v.to(device) for k, v in emb_host.items()}
    shape = PixArtShape()
    evaluator = PopulationEvaluator(rank, world, device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(step_idx, emb):
        gen.set_schedule(from_packed(row_for(step_idx)))
        return gen.generate_images(emb, images_per_prompt=1)[0]

    def run_region(first, count, emb, through_host, profile=False):
        """`count` steps starting at schedule index `first`; returns (seconds, launches, flops, last latents)."""
        flops = 0
        for i in range(first, first + count):
            r = row_for(i)
            S_, NB_ = r["S"], r["NB"]
            fl = np.unpackbits(np.frombuffer(bytes.fromhex(r["bits"]), np.uint8))[: S_ * NB_ * 3].reshape(S_, NB_, 3)
            flops += flops_per_image(trace_decisions(fl.astype(bool)), shape) * B
        barrier()
        l0 = tr.launches
        if profile:
            _lib.profile_start()  # CUDA-event pair around every kernel of the library, on its launch stream
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        host_out = None
        results = []
        if through_host:
            # the repo's public population API: every step's inputs come from pinned host memory and its latents go
            # back to pinned host memory, with the copies pipelined on a side stream (ecad_b200/population.py)
            out = evaluator.run_from_host(range(first, first + count), one_step, emb_host)
            results, host_out = out["device"], out["host"][-1]
        else:
            for i in range(first, first + count):
                results.append(one_step(i, emb))
        # the search driver needs every candidate's latents: gather over NCCL (no-op at N=1)
        parts = [[r_ * count + j for j in range(count)] for r_ in range(world)]
        evaluator.gather(results, parts, world * count)
        ev1.record()
        barrier()
        prof = _lib.profile_stop() if profile else None
        secs = ev0.elapsed_time(ev1) * 1e-3
        if world > 1:
            t = torch.tensor([secs], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t)
            f = torch.tensor([float(flops)], device=device, dtype=torch.float64)
            dist.all_reduce(f, op=dist.ReduceOp.SUM)
            flops = float(f)
        return secs, tr.launches - l0, flops, host_out, prof

    # warm-up (W >= 3 steps): allocations, descriptor cache, clocks
    for i in range(W):
        one_step(i, emb_dev)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    secs, launches, flops, _, prof = run_region(W, K, emb_dev, through_host=False, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    run_region(W, K, None, through_host=True)  # untimed: allocates the evaluator's pinned / device staging buffers
    sampler_e2e = ClockSampler(local_rank)
    if rank == 0:
        sampler_e2e.start()
    secs_e2e, _, _, host_out, _ = run_region(W, K, None, through_host=True)
    clocks_e2e = sampler_e2e.stop() if rank == 0 else None

    images = world * K * B
    value = images / secs
    e2e_value = images / secs_e2e
    h2d = sum(v.numel() * v.element_size() for v in emb_host.values())
    d2h = B * 4 * 32 * 32 * 4

    line = None
    if rank == 0:
        # dominant kernel = the tcgen05 GEMM (all epilogues): achieved = algorithmic 2*M*N*K of every GEMM launch in
        # the timed region / the sum of their CUDA-event durations (events recorded on the launch stream inside
        # libecad_b200); the kernel runs inside a long step, so the peak is the measured SUSTAINED bf16 figure.
        g = prof["gemm"]
        achieved = g["flops"] / (g["total_ms"] * 1e-3) / 1e12
        alone = time_dominant_gemm(device, peaks)
        roof = {
            "bound": "tensor", "kernel": "gemm2_bf16_kernel<BN,EPI> (all GEMM launches of the timed region)",
            "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
            "frac": achieved / peaks["tf_sustained"],
            "peak_source": f"{peaks['src']} sustained bf16 (kernel timed inside a long step)",
            "launches": g["launches"], "avg_ms_per_launch": g["total_ms"] / max(g["launches"], 1),
            "flops_per_launch": g["flops"] / max(g["launches"], 1),
            "share_of_step": g["total_ms"] * 1e-3 / secs,
            "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
            "timed_alone": alone,
            "other_kernels": {
                "attention": {"launches": prof["attention"]["launches"], "ms": prof["attention"]["total_ms"],
                              "tflops": prof["attention"]["flops"] / max(prof["attention"]["total_ms"], 1e-9) / 1e9},
                "glue_residual_ln": {"launches": prof["glue"]["launches"], "ms": prof["glue"]["total_ms"],
                                     "hbm_gbs": prof["glue"]["bytes"] / max(prof["glue"]["total_ms"], 1e-9) / 1e6,
                                     "hbm_frac": prof["glue"]["bytes"] / max(prof["glue"]["total_ms"], 1e-9) / 1e6
                                     / peaks["hbm"]},
                "other": {"launches": prof["other"]["launches"], "ms": prof["other"]["total_ms"]},
            },
        }
        roof["step_tflops"] = flops / secs / 1e12 / world
        roof["step_frac_of_sustained"] = roof["step_tflops"] / peaks["tf_sustained"]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            times, cores = cpu_oracle_images_per_s([row_for(W), row_for(W)], warmup=1)
            cpu = {"value": len(times) / sum(times), "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": "1 image (1 prompt, 2 CFG samples/forward), 20 steps, the first timed step's candidate "
                             "schedule, after 1 warm-up image"}
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": workload_text(args.fixed_schedule, B),
                "images_per_step": B, "l2": "working set (9.9 GB of caches + activations) >> 126 MB L2",
                "algorithmic_tflop_per_image": flops / images / 1e12,
            },
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "sm_mhz": clocks_e2e["sm_mhz"] if clocks_e2e else None},
            "gpu_launches": launches, "clocks": clocks,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)


# The host mirrors the reference's `print("WARNING: No cached ... found. Recomputing.")` on stdout; the bench contract is
# ONE JSON line on stdout, so everything else goes to stderr.
_REAL_STDOUT = sys.stdout

if __name__ == "__main__":
    sys.stdout = sys.stderr
    main()
