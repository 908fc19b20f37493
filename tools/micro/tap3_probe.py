"""Probe of the TAP3 convolution path: error against F.conv2d per tap pattern."""
import math, sys
from pathlib import Path
import torch
import torch.nn.functional as F
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib

b, h, w, cin, cout = 24, 64, 64, 128, 128
g = torch.Generator().manual_seed(0)
x = torch.randn(b, cin, h, w, generator=g)
xb = torch.zeros(b, h + 2, w + 2, cin, dtype=torch.bfloat16, device="cuda")
xb[:, 1:-1, 1:-1] = x.permute(0, 2, 3, 1).to("cuda", torch.bfloat16)
xi = xb[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float()
for name, mask in [("all", None), ("kx0", (None, 0)), ("kx1", (None, 1)), ("kx2", (None, 2)), ("center", (1, 1))]:
    wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    if mask is not None:
        m = torch.zeros(3, 3)
        ky, kx = mask
        if ky is None:
            m[:, kx] = 1
        else:
            m[ky, kx] = 1
        wt = wt * m
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).to("cuda", torch.bfloat16).contiguous()
    out = torch.zeros(b, h + 2, w + 2, cout, dtype=torch.bfloat16, device="cuda")
    _lib.conv_nhwc(xb, wp, None, out, h, w, 9)
    torch.cuda.synchronize()
    ref = F.conv2d(xi, wt.to(torch.bfloat16).float().cuda(), None, padding=1)
    got = out[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float()
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"{name:7s} rel err {err:.4f}")
