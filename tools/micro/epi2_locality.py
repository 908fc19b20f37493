"""Does the gated-residual GEMM epilogue run at a higher HBM rate when a tile's rows are contiguous in memory?
Same M*N elements of residual stream: (M=51200, N=1152: a 256-column tile touches 1 KB of every 4.6 KB row) versus
(M=230400, N=256: a tile's 128 x 256 fp32 block is one contiguous 128 KB range)."""
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402

K = 1152
g = torch.Generator(device="cuda").manual_seed(0)
for M, N in ((51200, 1152), (230400, 256), (115200, 512)):
    tokens, samples = 256, M // 256
    a = (torch.randn(M, K, device="cuda", generator=g)).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    cache = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    xb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for with_side in (True, False):
        def run():
            _lib.gemm_gated_residual(a, w, b, x, cache if with_side else None, tokens, xb=xb if with_side else None)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        byts = M * K * 2 + M * N * 8 + (M * N * 4 if with_side else 0)
        print(f"M={M:6d} N={N:4d} cache+xb={with_side!s:5s}: {ms * 1e3:7.1f} us  {byts / ms / 1e6:7.0f} GB/s  "
              f"{2 * M * N * K / ms / 1e9:6.0f} TFLOP/s")
