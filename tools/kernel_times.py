"""Per-kernel timing at the batch-100 shapes of BASELINE config 2 (CUDA events, kernels timed alone, mean of 10
after 3 warm-ups).  Prints one line per kernel with achieved TFLOP/s or GB/s against MEASURED_PEAKS.json."""
import json
import math
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ecad_b200 import _lib  # noqa: E402

peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {
    "hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
S, N, D, H, T = int(sys.argv[1]) if len(sys.argv) > 1 else 200, 256, 1152, 16, 120
M = S * N
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
bf = torch.bfloat16


def timed(fn, iters=6, reps=4):
    """mean per-launch ms; `reps` back-to-back launches per event pair so host launch gaps do not count"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    return statistics.mean(ts)


def rnd(*shape, scale=1.0, dtype=bf):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)


rows = []


def report(name, ms, flops=None, bytes_=None):
    s = f"{name:46s} {ms*1e3:9.1f} us"
    if flops:
        tf = flops / ms / 1e9
        s += f"  {tf:8.1f} TFLOP/s ({tf / peaks['bf16_tflops']:.2f} of burst)"
    if bytes_:
        gb = bytes_ / ms / 1e6
        s += f"  {gb:8.1f} GB/s ({gb / peaks['hbm_gbs']:.2f} of HBM)"
    print(s, flush=True)
    rows.append({"kernel": name, "us": ms * 1e3, "flops": flops, "bytes": bytes_})


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

w_qkv, b_qkv = rnd(3 * D, D, scale=1 / math.sqrt(D)), torch.randn(3 * D, device=dev, generator=g)
w_d, b_d = rnd(D, D, scale=1 / math.sqrt(D)), torch.randn(D, device=dev, generator=g)
w_f1, b_f1 = rnd(4 * D, D, scale=1 / math.sqrt(D)), torch.randn(4 * D, device=dev, generator=g)
w_f2 = rnd(D, 4 * D, scale=1 / math.sqrt(4 * D))

report("residual_ln LN only (x->h)", timed(lambda: _lib.residual_ln(
    x, N, h=h, shift_table=table[0], scale_table=table[1], shift_temb=temb, scale_temb=temb[:, D:],
    temb_stride=6 * D)), bytes_=M * D * 6)
report("residual_ln 3 reuse + LN", timed(lambda: _lib.residual_ln(
    x, N, reuse=[(cache[0], table[2], temb[:, 2 * D:]), (cache[1], None, None), (cache[2], table[5], temb[:, 5 * D:])],
    h=h, shift_table=table[0], scale_table=table[1], shift_temb=temb, scale_temb=temb[:, D:], temb_stride=6 * D)),
    bytes_=M * D * (4 + 4 + 6 + 2))
report("residual_ln 3 reuse only", timed(lambda: _lib.residual_ln(
    x, N, reuse=[(cache[0], table[2], temb[:, 2 * D:]), (cache[1], None, None), (cache[2], table[5], temb[:, 5 * D:])],
    temb_stride=6 * D)), bytes_=M * D * (4 + 4 + 6))
qkv = torch.empty(M, 3 * D, device=dev, dtype=bf)
report("gemm QKV plain -> row-major [M,1152]x[3456,1152] (executor path)", timed(lambda: _lib.gemm_bias(h, w_qkv, b_qkv, qkv)),
       flops=2.0 * M * 3 * D * D)
report("attention self NK=256, row-major operands (executor path)", timed(lambda: _lib.attention_ex(
    qkv, 3 * D, qkv[:, D:], qkv[:, 2 * D:], 3 * D, None, attn_o, S, H, N, 256)),
    flops=4.0 * S * H * N * N * 72, bytes_=M * D * 2 * 4)
report("  gemm QKV head-major scatter (round-1 path)", timed(lambda: _lib.gemm_headmajor(h, w_qkv, b_qkv, [q, k, v], H, N, N)),
       flops=2.0 * M * 3 * D * D)
report("  attention self NK=256, head-major operands", timed(lambda: _lib.attention(q, k, v, None, attn_o, S, H, N, 256)),
       flops=4.0 * S * H * N * N * 72, bytes_=S * H * N * 80 * 2 * 3 + M * D * 2)
report("gemm out1 gated-residual+cache+xb [M,1152]x[1152,1152]", timed(lambda: _lib.gemm_gated_residual(
    attn_o, w_d, b_d, x, cache[0], N, xb=xb, gate_table=table[2], gate_temb=temb[:, 2 * D:], temb_stride=6 * D)),
    flops=2.0 * M * D * D, bytes_=M * D * (2 + 4 + 4 + 2 + 2))
report("gemm Q2 plain -> first D columns of the row-major buffer", timed(lambda: _lib.gemm_bias(xb, w_d, b_d, qkv)),
       flops=2.0 * M * D * D)
report("attention cross NK=128 bias, row-major Q", timed(lambda: _lib.attention_ex(
    qkv, 3 * D, k2, v2, 0, bias2, attn_o, S, H, N, 128)),
    flops=4.0 * S * H * N * 128 * 72, bytes_=M * D * 2 * 2 + S * H * 256 * 80 * 2)
report("gemm out2 residual+cache (no gate)", timed(lambda: _lib.gemm_gated_residual(
    attn_o, w_d, b_d, x, cache[1], N)), flops=2.0 * M * D * D, bytes_=M * D * (2 + 4 + 4 + 2))
report("gemm FF1 bias+gelu [M,1152]x[4608,1152]", timed(lambda: _lib.gemm_bias(h, w_f1, b_f1, ffh, gelu=True)),
       flops=2.0 * M * 4 * D * D)
report("gemm FF2 gated-residual+cache [M,4608]x[1152,4608]", timed(lambda: _lib.gemm_gated_residual(
    ffh, w_f2, b_d, x, cache[2], N, gate_table=table[5], gate_temb=temb[:, 5 * D:], temb_stride=6 * D)),
    flops=2.0 * M * 4 * D * D)
lib = _lib.load()
lat = torch.randn(S, 4, 32, 32, device=dev, generator=g)
wt = torch.randn(16, D, device=dev, generator=g)
pos = torch.randn(N, D, device=dev, generator=g)
wo = torch.randn(32, D, device=dev, generator=g) / math.sqrt(D)
bo = torch.randn(32, device=dev, generator=g)
out = torch.empty(S, 8, 32, 32, device=dev)
x2 = torch.empty(M, D, device=dev)
report("patch_embed", timed(lambda: _lib.check(lib.ecadk_patch_embed(
    lat.data_ptr(), wt.data_ptr(), b_d.data_ptr(), pos.data_ptr(), x2.data_ptr(), S, 4, 32, 32, D, _lib.stream_ptr()))),
    bytes_=M * D * 4)
w_pad = torch.zeros(128, D, device=dev, dtype=bf)
w_pad[:32] = wo.to(bf)
b_pad = torch.zeros(128, device=dev)
report("final_layer (LN + tcgen05 GEMM + unpatchify)", timed(lambda: _lib.check(lib.ecadk_final_layer(
    x2.data_ptr(), table.data_ptr(), temb.data_ptr(), 0, w_pad.data_ptr(), b_pad.data_ptr(), h.data_ptr(),
    out.data_ptr(), S, 16, 16, D, 8, 1e-6, _lib.stream_ptr()))), bytes_=M * D * 4)
for n_, k_ in [(1152, 1152), (3456, 1152), (4608, 1152), (1152, 4608)]:
    a_ = h if k_ == D else ffh
    w_ = rnd(n_, k_, scale=1 / math.sqrt(k_))
    o_ = torch.empty(M, n_, device=dev, dtype=bf)
    report(f"gemm plain bias [M,{k_}]x[{n_},{k_}]", timed(lambda: _lib.gemm_bias(a_, w_, None, o_)),
           flops=2.0 * M * n_ * k_)
    report(f"  torch.matmul same shape (cuBLAS)", timed(lambda: torch.matmul(a_, w_.t(), out=o_)), flops=2.0 * M * n_ * k_)
total = sum(r["us"] for r in rows[:13] if "reuse" not in r["kernel"] and not r["kernel"].startswith("  ")) + rows[0]["us"]
print(f"sum of one dense block (2 LN + 7 GEMM/attn kernels): {total:.0f} us; x28 = {total * 28 / 1e3:.1f} ms per forward")
out = ROOT / "gpurun_out" / "kernel_times.json"
out.parent.mkdir(exist_ok=True)
out.write_text(json.dumps(rows, indent=1))
