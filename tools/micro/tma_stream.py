"""HBM -> shared-memory streaming rate of TMA for the attention kernels' access patterns (instrumented build):
ECAD_B200_LIB=tools/micro/libecad_b200_timing.so python tools/micro/tma_stream.py"""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.ecadk_debug_tma_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
tensors = 3 * 3200  # Q, K, V of 3200 (sample, head) items: 393 MB
buf = torch.randn(tensors * 256 * 80, device="cuda").to(torch.bfloat16)
for mode, name in ((0, "64 + 16 column boxes (Q, K)"), (1, "five 16-column boxes (V)"), (2, "contiguous 128-byte rows")):
    ms = C.c_float()
    _lib.check(lib.ecadk_debug_tma_stream(buf.data_ptr(), tensors, mode, C.byref(ms)))
    gb = tensors * 256 * 160 / 1e9
    print(f"{name:34s} {ms.value * 1e3:8.1f} us   {gb / ms.value * 1e3 / 1e3:6.2f} TB/s")
