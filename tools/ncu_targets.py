"""Launches every hot kernel twice at the batch-100 shapes (BASELINE config 2) so that one
`ncu --set full` pass captures each of them once warm.  Run under ncu only; prints nothing to judge."""
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402

S, N, D, H = 200, 256, 1152, 16
M = S * N
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
bf = torch.bfloat16


def rnd(*shape, scale=1.0, dtype=bf):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)


x = torch.randn(M, D, device=dev, generator=g)
h = rnd(M, D)
table = torch.randn(6, D, device=dev, generator=g) * 0.1
temb = torch.randn(S, 6 * D, device=dev, generator=g) * 0.1
cache = [rnd(M, D) for _ in range(3)]
xb = torch.empty(M, D, device=dev, dtype=bf)
q, k, v = (torch.zeros(S, H, N, 80, device=dev, dtype=bf) for _ in range(3))
k2, v2 = (torch.zeros(S, H, 128, 80, device=dev, dtype=bf) for _ in range(2))
bias2 = torch.zeros(S, 128, device=dev)
bias2[:, 120:] = float("-inf")
attn_o = torch.empty(M, D, device=dev, dtype=bf)
ffh = torch.empty(M, 4 * D, device=dev, dtype=bf)
qkv = torch.empty(M, 3 * D, device=dev, dtype=bf)
w_qkv, b_qkv = rnd(3 * D, D, scale=1 / math.sqrt(D)), torch.randn(3 * D, device=dev, generator=g)
w_d, b_d = rnd(D, D, scale=1 / math.sqrt(D)), torch.randn(D, device=dev, generator=g)
w_f1, b_f1 = rnd(4 * D, D, scale=1 / math.sqrt(D)), torch.randn(4 * D, device=dev, generator=g)
w_f2 = rnd(D, 4 * D, scale=1 / math.sqrt(4 * D))
reuse3 = [(cache[0], table[2], temb[:, 2 * D:]), (cache[1], None, None), (cache[2], table[5], temb[:, 5 * D:])]
ln = dict(shift_table=table[0], scale_table=table[1], shift_temb=temb, scale_temb=temb[:, D:], temb_stride=6 * D)

for _ in range(2):
    _lib.residual_ln(x, N, h=h, **ln)
    _lib.residual_ln(x, N, reuse=reuse3, h=h, **ln)
    _lib.gemm_bias(h, w_qkv, b_qkv, qkv)  # executor path: plain row-major q|k|v, gathered by the attention kernel
    _lib.attention_ex(qkv, 3 * D, qkv[:, D:], qkv[:, 2 * D:], 3 * D, None, attn_o, S, H, N, 256)
    _lib.gemm_gated_residual(attn_o, w_d, b_d, x, cache[0], N, xb=xb, gate_table=table[2], gate_temb=temb[:, 2 * D:],
                             temb_stride=6 * D)
    _lib.gemm_bias(xb, w_d, b_d, qkv)
    _lib.attention_ex(qkv, 3 * D, k2, v2, 0, bias2, attn_o, S, H, N, 128)
    _lib.gemm_gated_residual(attn_o, w_d, b_d, x, cache[1], N)
    _lib.gemm_bias(h, w_f1, b_f1, ffh, gelu=True)
    _lib.gemm_gated_residual(ffh, w_f2, b_d, x, cache[2], N, gate_table=table[5], gate_temb=temb[:, 5 * D:],
                             temb_stride=6 * D)
    torch.cuda.synchronize()
