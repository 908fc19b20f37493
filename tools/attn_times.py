"""Attention kernels alone at the shapes of configs 2 / 3 / 4 (CUDA events, 3 launches per event pair)."""
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402

H, HP = 16, 80
for name, S, nq, nk, bias in [("c2 self", 200, 256, 256, False), ("c2 cross", 200, 256, 128, True),
                              ("c3 self", 32, 1024, 1024, False), ("c3 cross", 32, 1024, 128, True),
                              ("c4 self", 16, 4096, 4096, False), ("c4 cross", 16, 4096, 384, True)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.zeros(S, H, nq, HP, device="cuda", dtype=torch.bfloat16)
    k = torch.zeros(S, H, nk, HP, device="cuda", dtype=torch.bfloat16)
    v = torch.zeros(S, H, nk, HP, device="cuda", dtype=torch.bfloat16)
    for t in (q, k, v):
        t[..., :72] = torch.randn(t.shape[:-1] + (72,), device="cuda", generator=g).to(torch.bfloat16)
    b = None
    if bias:
        b = torch.zeros(S, nk, device="cuda")
        b[:, nk - 8:] = float("-inf")
    out = torch.empty(S, nq, H * 72, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        _lib.attention(q, k, v, b, out, S, H, nq, nk)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            _lib.attention(q, k, v, b, out, S, H, nq, nk)
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) / 3)
    ms = statistics.mean(ts)
    fl = 4.0 * S * H * nq * nk * 72
    print(f"{name:9s} S={S:3d} Nq={nq:4d} Nk={nk:4d}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s (real d=72)  "
          f"{fl/ms/1e9*80/72:7.1f} incl. padding", flush=True)
    # row-major operands (the executor's path): Q (+ K, V for self-attention) gathered from the [M, 3456] projection output
    qkv = torch.randn(S * nq, 3 * H * 72, device="cuda", generator=g).to(torch.bfloat16)
    if bias:
        run_rm = lambda: _lib.attention_ex(qkv, 3 * H * 72, k, v, 0, b, out, S, H, nq, nk)  # noqa: E731
    else:
        run_rm = lambda: _lib.attention_ex(qkv, 3 * H * 72, qkv[:, H * 72:], qkv[:, 2 * H * 72:], 3 * H * 72, None, out,  # noqa: E731
                                           S, H, nq, nk)
    for _ in range(3):
        run_rm()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            run_rm()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) / 3)
    print(f"{'':9s} row-major operands:        {statistics.mean(ts)*1e3:8.1f} us", flush=True)

# FLUX joint attention (head_dim 128): config 5 shape and the 256x256 shape
for name, S, Hh, n in [("c5 joint", 4, 24, 4608), ("flux256", 16, 24, 768)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(S, Hh, n, 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
    out = torch.empty(S, n, Hh * 128, device="cuda", dtype=torch.bfloat16)
    lib = _lib.load()

    def run():
        _lib.check(lib.ecadk_attention_d128(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), Hh * 128, None, 0,
                                            S, Hh, n, n, _lib.stream_ptr()), "attention_d128")

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            run()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) / 3)
    ms = statistics.mean(ts)
    fl = 4.0 * S * Hh * n * n * 128
    print(f"{name:9s} S={S:3d} N={n:4d} d=128: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s")
