"""Race detector for the persistent tcgen05 kernels: every kernel is launched many times on the same inputs and each
output is compared BITWISE with the first one (a kernel whose barriers leave a window open shows up as a mismatch or a
hang long before a parity test with a bf16 tolerance notices).   python tools/micro/determinism_stress.py [iters]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402

lib = _lib.load()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 400
H, HP = 16, 80
g = torch.Generator(device="cuda").manual_seed(0)
bf = torch.bfloat16


def stress(name, run, out, n):
    run()
    torch.cuda.synchronize()
    ref = out.clone()
    bad, worst = 0, 0.0
    for i in range(n):
        run()
        if i % 8 == 7 or i == n - 1:
            if not torch.equal(out, ref):
                bad += 1
                worst = max(worst, float((out.float() - ref.float()).abs().max()))
    torch.cuda.synchronize()
    print(f"{name:46s} {n:5d} launches  mismatching checks {bad:4d}  max abs diff {worst:.3e}", flush=True)


def attn_case(name, S, nq, nk, bias, n):
    qkv = torch.randn(S * nq, 3 * H * 72, device="cuda", generator=g).to(bf)
    out = torch.empty(S, nq, H * 72, device="cuda", dtype=bf)
    if bias:
        k = torch.zeros(S, H, nk, HP, device="cuda", dtype=bf)
        v = torch.zeros(S, H, nk, HP, device="cuda", dtype=bf)
        for t in (k, v):
            t[..., :72] = torch.randn(S, H, nk, 72, device="cuda", generator=g).to(bf)
        b = torch.zeros(S, nk, device="cuda")
        b[:, nk - 8:] = -10000.0
        stress(name, lambda: _lib.attention_ex(qkv, 3 * H * 72, k, v, 0, b, out, S, H, nq, nk), out, n)
    else:
        stress(name, lambda: _lib.attention_ex(qkv, 3 * H * 72, qkv[:, H * 72:], qkv[:, 2 * H * 72:], 3 * H * 72, None,
                                               out, S, H, nq, nk), out, n)


attn_case("pair2 self   S=200 256x256 (row-major)", 200, 256, 256, False, iters)
attn_case("pair2 cross  S=200 256x128 bias", 200, 256, 128, True, iters)
attn_case("pair2 self   S=8   256x256", 8, 256, 256, False, iters)
attn_case("flash  self  S=32 1024x1024", 32, 1024, 1024, False, iters // 2)
attn_case("flash  cross S=16 4096x384 bias", 16, 4096, 384, True, iters // 2)
attn_case("flash2 self  S=16 4096x4096", 16, 4096, 4096, False, iters // 8)
# FLUX joint attention
S, Hh, n = 4, 24, 4608
q, k, v = (torch.randn(S, Hh, n, 128, device="cuda", generator=g).to(bf) for _ in range(3))
out = torch.empty(S, n, Hh * 128, device="cuda", dtype=bf)
stress("flash d=128 S=4 4608x4608",
       lambda: _lib.check(lib.ecadk_attention_d128(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), Hh * 128, None,
                                                   0, S, Hh, n, n, _lib.stream_ptr()), "attention_d128"), out, iters // 8)
# GEMMs: plain bias, the pair-tile kernel (M = 51200) and the 1-CTA kernel (M = 2048), several shapes
for M, N, K in [(51200, 1152, 1152), (51200, 4608, 1152), (51200, 1152, 4608), (2048, 1152, 1152), (2048, 1152, 4608),
                (512, 1152, 1152), (512, 4608, 1152)]:
    a = torch.randn(M, K, device="cuda", generator=g).to(bf)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).to(bf)
    bias = torch.randn(N, device="cuda", generator=g)
    o = torch.empty(M, N, device="cuda", dtype=bf)
    stress(f"gemm bias M={M} N={N} K={K}", lambda: _lib.gemm_bias(a, w, bias, o, False), o, iters // (4 if M > 4096 else 1))
