"""Batch-1 GEMMs (M = 512): GPU time per launch with the weights COLD (a different weight matrix per launch, 48 of
them cycling: > L2) versus HOT (the same matrix every launch), measured as CUDA-graph replays of 48 PDL-chained
launches so that host launch cost does not count.  Answers: how much of a small GEMM is the HBM latency of its weight
stream (what an L2 prefetch of the next kernel's weights could remove)?"""
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
M, D = 512, 1152
NW = 48
ws = torch.empty(16 << 20, dtype=torch.uint8, device=dev)


def bench(name, n, k, residual, split):
    a = torch.randn(M, k, device=dev, generator=g).to(torch.bfloat16)
    w = [(torch.randn(n, k, device=dev, generator=g) / math.sqrt(k)).to(torch.bfloat16) for _ in range(NW)]
    bias = torch.randn(n, device=dev, generator=g)
    x = torch.randn(M, n, device=dev, generator=g)
    cache = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
    table = torch.zeros(n, device=dev)
    _lib.set_splitk_workspace(ws if split else None)

    def launch(i):
        if residual:
            _lib.gemm_gated_residual(a, w[i], bias, x, cache, 256, gate_table=table)
        else:
            _lib.gemm_bias(a, w[i], bias, out)

    res = {}
    for mode in ("cold", "hot"):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for i in range(3):
                launch(i)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for i in range(NW):
                    launch(i if mode == "cold" else 0)
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        for _ in range(10):
            gr.replay()
        en.record()
        torch.cuda.synchronize()
        res[mode] = st.elapsed_time(en) * 1e3 / (10 * NW)
    print(f"{name:44s} split={int(split)}  cold {res['cold']:6.2f} us   hot {res['hot']:6.2f} us", flush=True)
    _lib.set_splitk_workspace(None)


for split in (False, True):
    bench("QKV      [512,1152]x[3456,1152] plain", 3456, 1152, False, split)
    bench("FF1-like [512,1152]x[4608,1152] plain", 4608, 1152, False, split)
    bench("out-proj [512,1152]x[1152,1152] residual", 1152, 1152, True, split)
    bench("FF2      [512,4608]x[1152,4608] residual", 1152, 4608, True, split)
