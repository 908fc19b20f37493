"""Phase clocks of the LAST joint-attention launch of a real FLUX.1-dev forward (config 5) - instrumented build only:
ECAD_B200_LIB=tools/micro/libecad_b200_timing.so python tools/micro/flux_attn_in_situ.py"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402
from ecad_b200.flux_pipeline import latent_image_ids  # noqa: E402
from ecad_b200.flux_transformer import B200FluxTransformer2D  # noqa: E402
from ecad_b200.transformer import SequentialDiTScheduler  # noqa: E402
from ecad_b200.weights import FluxConfig  # noqa: E402

lib = _lib.load()
model = B200FluxTransformer2D.from_random_init(SequentialDiTScheduler(1), None, FluxConfig(), seed=0, on_device=True)
B, N, T = 4, 4096, 512
g = torch.Generator(device="cuda").manual_seed(1)
lat = torch.randn(B, N, 64, device="cuda", generator=g)
emb = torch.randn(B, T, 4096, device="cuda", generator=g) * 0.2
pooled = torch.randn(B, 768, device="cuda", generator=g) * 0.2
ids, tids = latent_image_ids(B, 64, 64), torch.zeros(B, T, 3)
t, guid = torch.full((B,), 0.7, device="cuda"), torch.full((B,), 3.5, device="cuda")
for _ in range(2):
    model.reset_cache()
    model(lat, emb, pooled, t, ids, tids, guid, return_dict=False)
torch.cuda.synchronize()
buf = (C.c_uint * (148 * 32))()
lib.ecadk_debug_attn_timing.argtypes = [C.POINTER(C.c_uint)]
_lib.check(lib.ecadk_debug_attn_timing(buf))
a = np.frombuffer(buf, dtype=np.uint32).reshape(148, 32).astype(np.float64)
mma, s0 = a[:, 0:8], a[:, 8:16]
nb = mma[:, 7].mean()
print(f"in situ (last single-stream block): key blocks per CTA {nb:.0f}; clocks per key block")
for i, nm in enumerate(["wait K", "issue QK0", "wait P1 + issue PV1", "issue QK1", "wait V", "wait P0", "TOTAL loop"]):
    print(f"  MMA thread   {nm:22s} {mma[:, i].mean() / nb:8.1f}")
for i, nm in enumerate(["wait S", "TMEM load", "max + exchange", "exp + P store issue", "st wait + fence + arrive"]):
    print(f"  softmax t0   {nm:22s} {s0[:, i].mean() / nb:8.1f}")
q = model._ws["q"].float()
print("q rms", float(q.pow(2).mean().sqrt()), "k rms", float(model._ws["k"].float().pow(2).mean().sqrt()),
      "v rms", float(model._ws["v"].float().pow(2).mean().sqrt()))
