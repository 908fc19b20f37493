"""Dense single-forward timing for BASELINE configs 3 and 4 (parity-test configs, not bench lines):

    python tools/forward_bench.py c3   # PixArt-alpha 512x512, batch 16 (32 samples, N = 1024), no cache
    python tools/forward_bench.py c4   # PixArt-sigma 1024x1024, batch 8 (16 samples, N = 4096, T = 300), no cache
    python tools/forward_bench.py c2   # PixArt-alpha 256x256, batch 100 (200 samples, N = 256), no cache

Prints ms per forward and the algorithmic TFLOP/s (SURVEY.md section 8d formulas) against the measured bf16 peak.
"""
import json
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200.macs import PixArtShape  # noqa: E402
from ecad_b200.schedule import PixArtCacheSchedule  # noqa: E402
from ecad_b200.transformer import B200PixArtTransformer2D, SequentialDiTScheduler  # noqa: E402
from ecad_b200.weights import PixArtConfig, random_init_state_dict  # noqa: E402

CFGS = {
    "c2": dict(sample_size=32, batch=100, text=120),
    "c3": dict(sample_size=64, batch=16, text=120),
    "c4": dict(sample_size=128, batch=8, text=300),
}
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
c = CFGS[name]
cfg = PixArtConfig(sample_size=c["sample_size"], use_additional_conditions=False)
sd = random_init_state_dict(cfg, 0)
tr = B200PixArtTransformer2D(sd, cfg, SequentialDiTScheduler(20), PixArtCacheSchedule.default())
S = 2 * c["batch"]
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(S, 4, c["sample_size"], c["sample_size"], device="cuda", generator=g)
enc = torch.randn(S, c["text"], 4096, device="cuda", generator=g) * 0.2
mask = torch.ones(S, c["text"], device="cuda", dtype=torch.int64)
mask[:, c["text"] // 2:] = 0
ts = torch.tensor([500], device="cuda").expand(S)


def fwd():
    tr.reset_cache()
    return tr(x, encoder_hidden_states=enc, encoder_attention_mask=mask, timestep=ts,
              added_cond_kwargs={"resolution": None, "aspect_ratio": None}, return_dict=False)[0]


for _ in range(3):
    fwd()
torch.cuda.synchronize()
times = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fwd()
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
ms = statistics.mean(times)
N = (c["sample_size"] // 2) ** 2
shape = PixArtShape(tokens=N, text_tokens=c["text"])
flops = S * (28 * int(shape.flops_components().sum()) + shape.flops_fixed())
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {
    "bf16_tflops_sustained": 1400.0}
tf = flops / ms / 1e9
print(json.dumps({"config": name, "samples": S, "tokens": N, "text_tokens": c["text"], "ms_per_forward": ms,
                  "algorithmic_tflop_per_forward": flops / 1e12, "tflops": tf,
                  "frac_of_sustained_bf16": tf / peaks["bf16_tflops_sustained"]}))
