"""VAE decode throughput: B200VaeDecoder.decode on `batch` latents of `hw` x `hw` (CUDA events, after 2 warm-ups), with
the in-library profiler's split by kernel class."""
import json
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402
from ecad_b200.vae import B200VaeDecoder, VaeConfig, random_init_vae_state_dict  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 100
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dec = B200VaeDecoder(random_init_vae_state_dict(VaeConfig(), 0))
if len(sys.argv) > 3 and sys.argv[3] == "nofuse":  # Upsample2D as upsample kernel + 3x3 convolution
    dec.fused_upsample = False
lat = torch.randn(batch, 4, hw, hw, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
for _ in range(2):
    dec.decode(lat)
torch.cuda.synchronize()
times = []
l0 = dec.launches
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    dec.decode(lat)
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
launches = (dec.launches - l0) // 5
_lib.profile_start()
dec.decode(lat)
prof = _lib.profile_stop()
ms = statistics.mean(times)
flops = B200VaeDecoder.flops(batch, hw, hw)
print(json.dumps({"batch": batch, "latent": hw, "image": 8 * hw, "ms_per_decode": ms, "images_per_s": batch / ms * 1e3,
                  "algorithmic_tflop": flops / 1e12, "tflops": flops / ms / 1e9, "launches": launches,
                  "by_class": {k: {"launches": v["launches"], "ms": round(v["total_ms"], 3)} for k, v in prof.items()}}))
