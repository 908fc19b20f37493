"""Which part of the gated-residual epilogue costs the time?  The out-projection GEMM (M = 51200, N = K = 1152) with
side outputs switched off one by one (null pointers are part of the ABI: dead cache stores, no bf16 shadow, no gate):
    python tools/micro/epi2_sensitivity.py [samples]"""
import math
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402

S, N, D = int(sys.argv[1]) if len(sys.argv) > 1 else 200, 256, 1152
M = S * N
g = torch.Generator(device="cuda").manual_seed(0)
bf = torch.bfloat16


def timed(fn, iters=6, reps=4):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    return statistics.mean(ts) * 1e3


for K in (D, 4 * D):
    a = (torch.randn(M, K, device="cuda", generator=g)).to(bf)
    w = (torch.randn(D, K, device="cuda", generator=g) / math.sqrt(K)).to(bf)
    b = torch.randn(D, device="cuda", generator=g)
    x = torch.randn(M, D, device="cuda", generator=g)
    cache = torch.empty(M, D, device="cuda", dtype=bf)
    xb = torch.empty(M, D, device="cuda", dtype=bf)
    out = torch.empty(M, D, device="cuda", dtype=bf)
    table = torch.randn(D, device="cuda", generator=g) * 0.1
    temb = torch.randn(S, D, device="cuda", generator=g) * 0.1
    print(f"--- K = {K}")
    print(f"plain bias -> bf16 out (2 B/elem written)            {timed(lambda: _lib.gemm_bias(a, w, b, out)):8.1f} us")
    for name, kw in [
        ("x rmw only (8 B/elem)", dict(cache=None)),
        ("x rmw + gate", dict(cache=None, gate_table=table, gate_temb=temb, temb_stride=D)),
        ("x rmw + cache (10 B/elem)", dict(cache=cache)),
        ("x rmw + cache + gate", dict(cache=cache, gate_table=table, gate_temb=temb, temb_stride=D)),
        ("x rmw + cache + xb + gate (12 B/elem)", dict(cache=cache, xb=xb, gate_table=table, gate_temb=temb, temb_stride=D)),
    ]:
        c = kw.pop("cache")
        us = timed(lambda: _lib.gemm_gated_residual(a, w, b, x, c, N, **kw))
        print(f"{name:52s} {us:8.1f} us")
