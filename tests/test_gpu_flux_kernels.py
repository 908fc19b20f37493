"""Kernel-level parity of the FLUX building blocks (C ABI) against plain PyTorch fp32 statements of the same ops."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    got, ref = got.float(), ref.float()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-20))


def _bf(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("samples,heads,q_tokens,split", [(1, 4, 768, 512), (2, 3, 512, 0), (1, 24, 768, 512)])
def test_attention_d128(cuda_device, samples, heads, q_tokens, split):
    from ecad_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(q_tokens + heads)
    q = _bf(torch.randn(samples, heads, q_tokens, 128, device="cuda", generator=g) * 1.5)
    k = _bf(torch.randn(samples, heads, q_tokens, 128, device="cuda", generator=g) * 1.5)
    v = _bf(torch.randn(samples, heads, q_tokens, 128, device="cuda", generator=g))
    D = heads * 128
    ld = D + 64  # a pitch larger than the row: the single-stream blocks write into a wider concat buffer
    out = torch.full((samples, q_tokens - split, ld), float("nan"), device="cuda", dtype=torch.bfloat16)
    out_lo = torch.full((samples, max(split, 1), ld), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ecadk_attention_d128(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), ld,
                                        out_lo.data_ptr() if split else None, split, samples, heads, q_tokens,
                                        q_tokens, _lib.stream_ptr()), "attention_d128")
    torch.cuda.synchronize()
    s = torch.einsum("shqd,shkd->shqk", q.float(), k.float()) / math.sqrt(128)
    ref = torch.einsum("shqk,shkd->shqd", torch.softmax(s, -1), v.float()).permute(0, 2, 1, 3).reshape(samples, q_tokens, D)
    got = torch.cat([out_lo[:, :split, :D], out[:, :, :D]], dim=1) if split else out[:, :, :D]
    assert torch.isfinite(got.float()).all()
    assert _rel(got, ref) < 2e-2
    assert float((got.float() - ref).abs().mean() / ref.abs().mean()) < 8e-3
    assert torch.isnan(out[:, :, D:].float()).all()  # nothing written past the head columns


def test_qk_norm_rope(cuda_device):
    from ecad_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, S, split = 2, 3, 40, 12
    q = _bf(torch.randn(B, H, S, 128, device="cuda", generator=g) * 2)
    k = _bf(torch.randn(B, H, S, 128, device="cuda", generator=g) * 2)
    w = [torch.rand(128, device="cuda", generator=g) + 0.5 for _ in range(4)]
    ang = torch.rand(S, 64, device="cuda", generator=g) * 6.28
    cos, sin = torch.cos(ang).contiguous(), torch.sin(ang).contiguous()
    q0, k0 = q.clone(), k.clone()
    _lib.check(lib.ecadk_qk_norm_rope(q.data_ptr(), k.data_ptr(), w[0].data_ptr(), w[1].data_ptr(), w[2].data_ptr(),
                                      w[3].data_ptr(), cos.data_ptr(), sin.data_ptr(), B, H, S, split, 1e-6,
                                      _lib.stream_ptr()), "qk_norm_rope")
    torch.cuda.synchronize()

    def ref(x, w_img, w_txt):
        x = x.float()
        x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)
        wsel = torch.where(torch.arange(S, device="cuda")[:, None] < split, w_txt[None, :], w_img[None, :])
        x = x * wsel[None, None]
        xp = x.reshape(B, H, S, 64, 2)
        o0 = cos[None, None] * xp[..., 0] - sin[None, None] * xp[..., 1]
        o1 = sin[None, None] * xp[..., 0] + cos[None, None] * xp[..., 1]
        return torch.stack([o0, o1], -1).reshape(B, H, S, 128)

    assert _rel(q, ref(q0, w[0], w[2])) < 6e-3
    assert _rel(k, ref(k0, w[1], w[3])) < 6e-3


def test_strided_unary_axpy_and_f32_gemm(cuda_device):
    from ecad_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(9)
    src = _bf(torch.randn(300, 256, device="cuda", generator=g))
    dst = torch.zeros(300, 640, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ecadk_strided_unary(src.data_ptr(), dst[:, 128:].data_ptr(), 300, 256, 256, 640, 1, _lib.stream_ptr()))
    _lib.check(lib.ecadk_strided_unary(src.data_ptr(), dst.data_ptr(), 300, 128, 256, 640, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :128], src[:, :128])
    assert _rel(dst[:, 128:384], torch.nn.functional.gelu(src.float(), approximate="tanh")) < 8e-3
    assert float(dst[:, 384:].abs().max()) == 0.0

    y = torch.randn(1000, device="cuda", generator=g)
    x = torch.randn(1000, device="cuda", generator=g)
    y0 = y.clone()
    _lib.check(lib.ecadk_axpy_f32(y.data_ptr(), x.data_ptr(), -0.25, 1000, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.allclose(y, y0 - 0.25 * x, atol=1e-6)

    a = _bf(torch.randn(700, 64, device="cuda", generator=g))
    w = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    w[:64] = _bf(torch.randn(64, 64, device="cuda", generator=g) / 8)
    b = torch.zeros(128, device="cuda")
    b[:64] = torch.randn(64, device="cuda", generator=g)
    out = torch.full((700, 64), float("nan"), device="cuda")
    _lib.check(lib.ecadk_gemm_bias_f32(a.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), 700, 128, 64, 64, 64,
                                       _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert _rel(out, a.float() @ w[:64].float().T + b[:64]) < 1e-5


def test_headmajor_ex_joint_sequence_and_vector_only_gate(cuda_device):
    from ecad_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, T, N, D = 2, 4, 64, 192, 512
    S = T + N
    outs = [torch.zeros(B, H, S, 128, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    for rows, off in ((T, 0), (N, T)):
        a = _bf(torch.randn(B * rows, D, device="cuda", generator=g))
        w = _bf(torch.randn(3 * D, D, device="cuda", generator=g) / math.sqrt(D))
        bias = torch.randn(3 * D, device="cuda", generator=g)
        _lib.check(lib.ecadk_gemm_bias_headmajor_ex(a.data_ptr(), w.data_ptr(), bias.data_ptr(), outs[0].data_ptr(),
                                                    outs[1].data_ptr(), outs[2].data_ptr(), 3, H, 128, 128, rows, S,
                                                    off, B * rows, D, _lib.stream_ptr()))
        torch.cuda.synchronize()
        ref = (a.float() @ w.float().T + bias).view(B, rows, 3, H, 128)
        for pi in range(3):
            assert _rel(outs[pi][:, :, off:off + rows], ref[:, :, pi].permute(0, 2, 1, 3)) < 1e-2
    # gated residual with a per-sample gate vector and no table (FLUX adaLN-zero gates)
    m = B * N
    a = _bf(torch.randn(m, D, device="cuda", generator=g))
    w = _bf(torch.randn(D, D, device="cuda", generator=g) / math.sqrt(D))
    bias = torch.randn(D, device="cuda", generator=g)
    x = torch.randn(m, D, device="cuda", generator=g)
    x0 = x.clone()
    mod = torch.randn(B, 6 * D, device="cuda", generator=g)
    cache = torch.zeros(m, D, device="cuda", dtype=torch.bfloat16)
    _lib.gemm_gated_residual(a, w, bias, x, cache, N, gate_table=None, gate_temb=mod[:, 2 * D:], temb_stride=6 * D)
    torch.cuda.synchronize()
    o = a.float() @ w.float().T + bias
    ref_x = x0 + mod[:, 2 * D:3 * D].repeat_interleave(N, dim=0) * o
    assert _rel(x, ref_x) < 2e-5 * max(1.0, float(o.abs().max()))
    assert _rel(cache, o) < 1e-2
    # LayerNorm + modulation from the per-sample vectors only (null tables), D = 512
    h = torch.zeros(m, D, device="cuda", dtype=torch.bfloat16)
    _lib.residual_ln(x, N, h=h, shift_table=None, scale_table=None, shift_temb=mod[:, 0:], scale_temb=mod[:, D:],
                     temb_stride=6 * D)
    torch.cuda.synchronize()
    ref_h = torch.nn.functional.layer_norm(x, (D,), eps=1e-6) * (1 + mod[:, D:2 * D].repeat_interleave(N, 0)) + \
        mod[:, :D].repeat_interleave(N, 0)
    assert _rel(h, ref_h) < 6e-3


@pytest.mark.parametrize("m", [300, 20480])  # 1-CTA and 2-CTA (cta_group::2) GEMM kernels
def test_dual_output_gemm_and_split_a_gemm(cuda_device, m):
    from ecad_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(m)
    D, F = 512, 2048
    st = _lib.stream_ptr()
    # proj_mlp: pre-activation into the cache + GELU into the proj_out operand, from one GEMM
    a = _bf(torch.randn(m, D, device="cuda", generator=g))
    w = _bf(torch.randn(F, D, device="cuda", generator=g) / math.sqrt(D))
    b = torch.randn(F, device="cuda", generator=g)
    pre = torch.full((m, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    act = torch.full((m, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ecadk_gemm_bias_dual(a.data_ptr(), w.data_ptr(), b.data_ptr(), pre.data_ptr(), act.data_ptr(), m, F, D,
                                        F, F, st))
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T + b
    assert _rel(pre, ref) < 1e-2 and _rel(act, torch.nn.functional.gelu(ref, approximate="tanh")) < 1e-2
    act2 = torch.zeros_like(act)
    _lib.check(lib.ecadk_gemm_bias_dual(a.data_ptr(), w.data_ptr(), b.data_ptr(), None, act2.data_ptr(), m, F, D, F, F, st))
    torch.cuda.synchronize()
    assert torch.equal(act, act2)  # the pre-activation store is optional (dead cache slot)
    # proj_out over [attn | GELU(mlp)] read from two buffers, gated residual + cache
    tokens, samples = (m // 4, 4) if m % 128 == 0 else (m // 3 // 32 * 32 or 32, 0)
    if samples == 0:
        m2 = 288
        tokens, samples = 96, 3
    else:
        m2 = m
    a1 = _bf(torch.randn(m2, D, device="cuda", generator=g))
    a2 = _bf(torch.randn(m2, F, device="cuda", generator=g))
    wo = _bf(torch.randn(D, D + F, device="cuda", generator=g) / math.sqrt(D + F))
    bo = torch.randn(D, device="cuda", generator=g)
    x = torch.randn(m2, D, device="cuda", generator=g)
    x0 = x.clone()
    gate = torch.randn(samples, 3 * D, device="cuda", generator=g)
    cache = torch.zeros(m2, D, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ecadk_gemm2src_gated_residual_cache(a1.data_ptr(), D, a2.data_ptr(), wo.data_ptr(), bo.data_ptr(),
                                                       x.data_ptr(), cache.data_ptr(), gate[:, 2 * D:].data_ptr(), 3 * D,
                                                       tokens, m2, D, D + F, st))
    torch.cuda.synchronize()
    o = torch.cat([a1, a2], dim=1).float() @ wo.float().T + bo
    assert _rel(cache, o) < 1e-2
    assert _rel(x, x0 + gate[:, 2 * D:].repeat_interleave(tokens, 0) * o) < 2e-5 * max(1.0, float(o.abs().max()))
    rc = lib.ecadk_gemm2src_gated_residual_cache(a1.data_ptr(), 100, a2.data_ptr(), wo.data_ptr(), bo.data_ptr(),
                                                 x.data_ptr(), None, gate.data_ptr(), 3 * D, tokens, m2, D, D + F, st)
    assert rc != 0 and b"k1" in lib.ecadk_last_error()
