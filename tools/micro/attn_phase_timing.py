"""Phase timing of attn_flash_kernel from an instrumented build (nvcc ... -DECADK_ATTN_TIMING -o
tools/micro/libecad_b200_timing.so ecad_b200/csrc/capi.cu):   ECAD_B200_LIB=tools/micro/libecad_b200_timing.so python
tools/micro/attn_phase_timing.py [d128|d72]
Prints, per key block, the average clocks the MMA-issuing thread and one softmax warp of each query tile spend in each
phase (averaged over the CTAs)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402

lib = _lib.load()
mode = sys.argv[1] if len(sys.argv) > 1 else "d128"
g = torch.Generator(device="cuda").manual_seed(0)
if mode == "d128":
    S, H, n = 4, 24, 4608
    q, k, v = (torch.randn(S, H, n, 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
    out = torch.empty(S, n, H * 128, device="cuda", dtype=torch.bfloat16)

    def run():
        _lib.check(lib.ecadk_attention_d128(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), H * 128, None, 0,
                                            S, H, n, n, _lib.stream_ptr()))
elif mode in ("c2self", "c2cross"):
    S, H, n = 200, 16, 256
    nk = 256 if mode == "c2self" else 128
    q = torch.zeros(S, H, n, 80, device="cuda", dtype=torch.bfloat16)
    k, v = (torch.zeros(S, H, nk, 80, device="cuda", dtype=torch.bfloat16) for _ in range(2))
    for t in (q, k, v):
        t[..., :72] = torch.randn(t.shape[:-1] + (72,), device="cuda", generator=g).to(torch.bfloat16)
    b = None
    if nk == 128:
        b = torch.zeros(S, nk, device="cuda")
        b[:, 120:] = float("-inf")
    out = torch.empty(S, n, H * 72, device="cuda", dtype=torch.bfloat16)

    def run():
        _lib.attention(q, k, v, b, out, S, H, n, nk)
else:
    S, H, n = 16, 16, 4096
    q, k, v = (torch.zeros(S, H, n, 80, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    for t in (q, k, v):
        t[..., :72] = torch.randn(S, H, n, 72, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty(S, n, H * 72, device="cuda", dtype=torch.bfloat16)

    def run():
        _lib.attention(q, k, v, None, out, S, H, n, n)

for _ in range(3):
    run()
torch.cuda.synchronize()
buf = (C.c_uint * (148 * 32))()
lib.ecadk_debug_attn_timing.argtypes = [C.POINTER(C.c_uint)]
_lib.check(lib.ecadk_debug_attn_timing(buf))
a = np.frombuffer(buf, dtype=np.uint32).reshape(148, 32).astype(np.float64)
mma, s0, s1 = a[:, 0:8], a[:, 8:16], a[:, 16:24]
if mode.startswith("c2"):
    items = mma[:, 7].mean()
    print(f"{mode}: items per CTA {items:.1f}; clocks per item (mean over CTAs); whole MMA loop {a[:, 24].mean() / items:.0f}")
    for i, nm in enumerate(["wait K + Q0", "wait O_0 read out", "issue QK0", "wait P1 + issue PV1",
                            "wait Q1 + O_1 read out + QK1", "wait V + P0", "issue PV0"]):
        print(f"  MMA thread   {nm:26s} {mma[:, i].mean() / items:8.1f}")
    for tile, arr in ((0, s0), (1, s1)):
        for i, nm in enumerate(["wait S", "softmax", "wait O", "O out of TMEM", "wait staging tile", "staging writes + fence"]):
            print(f"  softmax t{tile}   {nm:26s} {arr[:, i].mean() / items:8.1f}")
    sys.exit(0)
nb = mma[:, 7].mean()
print(f"{mode}: key blocks per CTA {nb:.0f}; clocks per key block (mean over CTAs)")
names_m = ["wait K", "issue QK0", "wait P1 + issue PV1", "issue QK1", "wait V", "wait P0", "TOTAL loop"]
for i, nm in enumerate(names_m):
    print(f"  MMA thread   {nm:22s} {mma[:, i].mean() / nb:8.1f}")
names_s = ["wait S", "TMEM load", "max + exchange", "exp + P store issue", "st wait + fence + arrive", "-", "TOTAL loop"]
for tile, arr in ((0, s0), (1, s1)):
    for i, nm in enumerate(names_s):
        if nm != "-":
            print(f"  softmax t{tile}   {nm:22s} {arr[:, i].mean() / nb:8.1f}")
