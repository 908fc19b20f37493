"""Phase timing of attn_flash2_kernel (head dim 72) from an instrumented build:
    nvcc <build.py flags> -DECADK_ATTN_TIMING -o tools/micro/libecad_b200_timing.so ecad_b200/csrc/capi.cu
    ECAD_B200_LIB=tools/micro/libecad_b200_timing.so python tools/micro/attn_phase_timing2.py [c4|c3]
Prints, per key block, the average clocks the MMA-issuing thread and three softmax warps (tile 0 half 0, tile 0 half 1,
tile 1 half 1) spend in each phase, averaged over the CTAs."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402

lib = _lib.load()
mode = sys.argv[1] if len(sys.argv) > 1 else "c4"
S, H, n = (16, 16, 4096) if mode == "c4" else (32, 16, 1024)
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.zeros(S, H, n, 80, device="cuda", dtype=torch.bfloat16) for _ in range(3))
for t in (q, k, v):
    t[..., :72] = torch.randn(S, H, n, 72, device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty(S, n, H * 72, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    _lib.attention(q, k, v, None, out, S, H, n, n)
torch.cuda.synchronize()
buf = (C.c_uint * (148 * 32))()
lib.ecadk_debug_attn_timing.argtypes = [C.POINTER(C.c_uint)]
_lib.check(lib.ecadk_debug_attn_timing(buf))
a = np.frombuffer(buf, dtype=np.uint32).reshape(148, 32).astype(np.float64)
nb = a[:, 15].mean()
print(f"{mode}: key blocks per CTA {nb:.0f}; clocks per key block (mean over CTAs)")
for i, nm in enumerate(["wait s_free", "  of which wait K (+Q)", "wait K + issue QK", "wait V", "wait P", "issue PV"]):
    print(f"  MMA thread tile 0 {nm:28s} {a[:, i].mean() / nb:8.1f}")
for lab, off in (("softmax t0 half0", 8), ("softmax t0 half1", 24), ("softmax t1 half1", 16)):
    for i, nm in enumerate(["wait S", "TMEM load", "s_free + max + exchange", "(rescale) + exp", "wait pv_done",
                            "P store + fences + arrive", "TOTAL loop"]):
        print(f"  {lab}  {nm:28s} {a[:, off + i].mean() / nb:8.1f}")
