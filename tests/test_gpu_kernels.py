"""Kernel-level parity: every C-ABI entry point against a plain PyTorch fp32 statement of the same op.

bf16 kernels: inputs are the SAME bf16-rounded tensors on both sides, the reference accumulates in fp32, so the only
differences are accumulation order and the final bf16 rounding of the output (rel 2^-8).  Tolerances are written
next to each check.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

D, H, HD, HP = 1152, 16, 72, 80


def _rel_err(got, ref):
    got, ref = got.float(), ref.float()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-20))


def _bf(t):
    return t.to(torch.bfloat16)


@pytest.fixture(params=[None, "plan", (128, 2), (128, 3), (128, 4)],
                ids=["plain", "splitk", "splitk128x2", "splitk128x3", "splitk128x4"])
def splitk(request, cuda_device):
    """Run a GEMM test without and with a split-K workspace installed (small problems then take the cluster split-K
    kernel; the large shapes of the same test stay on the ordinary kernels either way), and with the plan forced to
    3- and 4-way splits of 128-wide tiles (taken wherever all clusters are resident)."""
    from ecad_b200 import _lib
    if request.param is None:
        _lib.set_splitk_workspace(None)
        yield None
        return
    ws = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
    _lib.set_splitk_workspace(ws)
    if request.param != "plan":
        _lib.set_splitk_force(*request.param)
    yield request.param
    torch.cuda.synchronize()
    _lib.set_splitk_force(0, 0)
    _lib.set_splitk_workspace(None)


# shapes the split-K planner takes: fewer than half as many 128-wide tiles as SMs
def _splits(m, n):
    return 2 * ((m + 127) // 128) * (n // 128) <= 148


def _check_split_count(splitk, launched, m, n):
    if splitk is None:
        assert launched == 0
    elif splitk == "plan":
        assert launched == (1 if _splits(m, n) else 0)
    elif ((m + 127) // 128) * (n // 128) <= 30:  # forced plan: every cluster of up to 4 is resident for <= 30 tiles
        assert launched == 1


@pytest.mark.parametrize("m,n,k", [(512, 1152, 1152), (384, 4608, 1152), (1000, 2304, 1152), (256, 1152, 4608),
                                   (130, 3456, 1152), (20480, 1152, 1152), (256, 1152, 4096)])
@pytest.mark.parametrize("gelu", [False, True])
def test_gemm_bias(cuda_device, splitk, m, n, k, gelu):
    from ecad_b200 import _lib
    n_split0 = _lib.splitk_launches()
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = _bf(torch.randn(m, k, device="cuda", generator=g))
    w = _bf(torch.randn(n, k, device="cuda", generator=g) / math.sqrt(k))
    bias = torch.randn(n, device="cuda", generator=g)
    out = torch.full((m, n), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.gemm_bias(a, w, bias, out, gelu=gelu)
    torch.cuda.synchronize()
    _check_split_count(splitk, _lib.splitk_launches() - n_split0, m, n)
    ref = a.float() @ w.float().T + bias
    if gelu:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    assert torch.isfinite(out.float()).all()
    # bf16 output rounding 2^-8 relative + tanh.approx (2^-11); measured against the max magnitude
    assert _rel_err(out, ref) < 1.0e-2, _rel_err(out, ref)
    # and tight in the mean
    assert float((out.float() - ref).abs().mean() / ref.abs().mean()) < 4e-3


@pytest.mark.parametrize("samples,tokens,k,gated,with_xb", [(2, 256, 1152, True, True), (3, 256, 4608, True, False),
                                                            (2, 256, 1152, False, False)])
def test_gemm_gated_residual_cache(cuda_device, splitk, samples, tokens, k, gated, with_xb):
    from ecad_b200 import _lib
    n_split0 = _lib.splitk_launches()
    g = torch.Generator(device="cuda").manual_seed(7)
    m = samples * tokens
    a = _bf(torch.randn(m, k, device="cuda", generator=g))
    w = _bf(torch.randn(D, k, device="cuda", generator=g) / math.sqrt(k))
    bias = torch.randn(D, device="cuda", generator=g)
    x = torch.randn(m, D, device="cuda", generator=g)
    x0 = x.clone()
    table = torch.randn(6, D, device="cuda", generator=g)
    temb = torch.randn(samples, 6 * D, device="cuda", generator=g)
    cache = torch.zeros(m, D, device="cuda", dtype=torch.bfloat16)
    xb = torch.zeros(m, D, device="cuda", dtype=torch.bfloat16) if with_xb else None
    _lib.gemm_gated_residual(a, w, bias, x, cache, tokens, xb=xb,
                             gate_table=table[2] if gated else None,
                             gate_temb=temb[:, 2 * D:] if gated else None, temb_stride=6 * D)
    torch.cuda.synchronize()
    _check_split_count(splitk, _lib.splitk_launches() - n_split0, m, D)
    o = a.float() @ w.float().T + bias
    gate = (table[2][None] + temb[:, 2 * D:3 * D]).repeat_interleave(tokens, dim=0) if gated else 1.0
    x_ref = x0 + gate * o
    assert _rel_err(cache, o) < 1e-2
    assert _rel_err(x, x_ref) < 2e-5 * max(1.0, float(o.abs().max()))  # fp32 stream: accumulation-order noise only
    if with_xb:
        assert _rel_err(xb, x_ref) < 1e-2


@pytest.mark.parametrize("parts,samples,tokens,tokens_pad", [(3, 2, 256, 256), (1, 2, 256, 256), (2, 3, 120, 128)])
def test_gemm_headmajor(cuda_device, splitk, parts, samples, tokens, tokens_pad):
    from ecad_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(11)
    m = samples * tokens
    a = _bf(torch.randn(m, D, device="cuda", generator=g))
    w = _bf(torch.randn(parts * D, D, device="cuda", generator=g) / math.sqrt(D))
    bias = torch.randn(parts * D, device="cuda", generator=g)
    outs = [torch.zeros(samples, H, tokens_pad, HP, device="cuda", dtype=torch.bfloat16) for _ in range(parts)]
    _lib.gemm_headmajor(a, w, bias, outs, H, tokens, tokens_pad)
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().T + bias).view(samples, tokens, parts, H, HD)
    for pi in range(parts):
        r = ref[:, :, pi].permute(0, 2, 1, 3)  # [S, H, T, 72]
        got = outs[pi]
        assert _rel_err(got[:, :, :tokens, :HD], r) < 1e-2
        assert float(got[:, :, :, HD:].abs().max()) == 0.0  # padding columns untouched
        if tokens_pad > tokens:
            assert float(got[:, :, tokens:, :].abs().max()) == 0.0  # padding tokens untouched


def _attn_ref(q, k, v, bias):
    s = torch.einsum("shqd,shkd->shqk", q.float(), k.float()) / math.sqrt(HD)
    if bias is not None:
        s = s + bias[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("shqk,shkd->shqd", p, v.float())
    S, Hh, Q, _ = o.shape
    return o.permute(0, 2, 1, 3).reshape(S, Q, Hh * HD)


@pytest.mark.parametrize("samples,nk,use_bias", [(2, 256, False), (3, 128, True), (1, 128, False), (2, 256, True)])
def test_attention(cuda_device, samples, nk, use_bias):
    from ecad_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(5 + nk)
    q_tokens = 256

    def mk(tokens, scale):
        t = torch.zeros(samples, H, tokens, HP, device="cuda", dtype=torch.bfloat16)
        t[..., :HD] = _bf(torch.randn(samples, H, tokens, HD, device="cuda", generator=g) * scale)
        return t

    q, k, v = mk(q_tokens, 2.0), mk(nk, 2.0), mk(nk, 1.0)
    bias = None
    if use_bias:
        bias = torch.zeros(samples, nk, device="cuda")
        valid = [nk - 8, 3, 57][:samples] if nk == 128 else [200, 256][:samples]
        for s, L in enumerate(valid):
            real = min(nk, 120) if nk == 128 else nk
            bias[s, L:real] = -10000.0
            bias[s, real:] = float("-inf")
    out = torch.full((samples, q_tokens, H * HD), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.attention(q, k, v, bias, out, samples, H, q_tokens, nk)
    torch.cuda.synchronize()
    ref = _attn_ref(q[..., :HD], k[..., :HD], v[..., :HD], bias)
    assert torch.isfinite(out.float()).all()
    # P is rounded to bf16 before PV and the output to bf16: ~2^-8 relative
    assert _rel_err(out, ref) < 1.5e-2, _rel_err(out, ref)
    assert float((out.float() - ref).abs().mean() / ref.abs().mean()) < 6e-3


@pytest.mark.parametrize("nk,use_bias", [(256, False), (128, True), (256, True), (128, False)])
def test_attention_many_items_per_cta(cuda_device, nk, use_bias):
    """50 samples x 16 heads = 800 work items on 148 persistent CTAs: five to six items per CTA, so every ring
    (per-tile Q slots, K/V double buffers, staged output tiles, TMEM tiles) wraps several times."""
    from ecad_b200 import _lib
    samples, q_tokens = 50, 256
    g = torch.Generator(device="cuda").manual_seed(77 + nk)

    def mk(tokens, scale):
        t = torch.zeros(samples, H, tokens, HP, device="cuda", dtype=torch.bfloat16)
        t[..., :HD] = _bf(torch.randn(samples, H, tokens, HD, device="cuda", generator=g) * scale)
        return t

    q, k, v = mk(q_tokens, 2.0), mk(nk, 2.0), mk(nk, 1.0)
    bias = None
    if use_bias:
        bias = torch.zeros(samples, nk, device="cuda")
        lens = torch.randint(1, nk - 8, (samples,), generator=torch.Generator().manual_seed(1))
        for s in range(samples):
            bias[s, int(lens[s]):nk - 8] = -10000.0
            bias[s, nk - 8:] = float("-inf")
    out = torch.full((samples, q_tokens, H * HD), float("nan"), device="cuda", dtype=torch.bfloat16)
    for _ in range(2):  # back-to-back launches (programmatic dependent launch between them)
        _lib.attention(q, k, v, bias, out, samples, H, q_tokens, nk)
    torch.cuda.synchronize()
    ref = _attn_ref(q[..., :HD], k[..., :HD], v[..., :HD], bias)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().amax(dim=(1, 2)) / ref.abs().amax(dim=(1, 2))  # per sample
    assert float(err.max()) < 1.5e-2, err


@pytest.mark.parametrize("samples,q_tokens,nk,real_keys", [
    (1, 512, 512, None),     # self-attention, 4 key blocks, two query pairs
    (2, 1024, 1024, None),   # PixArt 512x512 self-attention shape
    (2, 256, 384, 300),      # PixArt-sigma cross-attention: 300 text tokens padded to 384
    (1, 1024, 128, 120),     # cross-attention at 512x512: many queries, one key block
])
def test_attention_streaming(cuda_device, samples, q_tokens, nk, real_keys):
    """Keys streamed in 128-key blocks with an online softmax (lazy rescale) == one-shot softmax."""
    from ecad_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(q_tokens + nk)

    def mk(tokens, scale):
        t = torch.zeros(samples, H, tokens, HP, device="cuda", dtype=torch.bfloat16)
        t[..., :HD] = _bf(torch.randn(samples, H, tokens, HD, device="cuda", generator=g) * scale)
        return t

    # large score spread (scale 3) so the running maximum really moves between blocks and the rescale path runs
    q, k, v = mk(q_tokens, 3.0), mk(nk, 3.0), mk(nk, 1.0)
    bias = None
    if real_keys is not None:
        bias = torch.zeros(samples, nk, device="cuda")
        for s in range(samples):
            bias[s, real_keys - 17 * (s + 1):real_keys] = -10000.0
        bias[:, real_keys:] = float("-inf")
    out = torch.full((samples, q_tokens, H * HD), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.attention(q, k, v, bias, out, samples, H, q_tokens, nk)
    torch.cuda.synchronize()
    ref = _attn_ref(q[..., :HD], k[..., :HD], v[..., :HD], bias)
    assert torch.isfinite(out.float()).all()
    assert _rel_err(out, ref) < 2e-2, _rel_err(out, ref)
    assert float((out.float() - ref).abs().mean() / ref.abs().mean()) < 8e-3


@pytest.mark.parametrize("samples,q_tokens,nk,cross", [(20, 256, 256, False), (20, 256, 128, True), (2, 1024, 1024, False),
                                                         (2, 1024, 384, True)])
def test_attention_rowmajor_operands(cuda_device, samples, q_tokens, nk, cross):
    """ecadk_attention_ex: Q (and for self-attention K, V) read straight from the plain row-major projection output
    [samples*tokens, 3*1152] through 3-D tensor maps - the 72 -> 80 column padding of the shared-memory tiles is the
    maps' out-of-bounds zero fill.  Must equal the head-major path bit for bit (same tiles reach the same MMAs)."""
    from ecad_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(41 + nk + q_tokens)
    D = H * HD
    qkv = _bf(torch.randn(samples * q_tokens, 3 * D, device="cuda", generator=g) * 1.5)

    def headmajor(cols, tokens):  # [S*tokens, D] -> [S, H, tokens, 80] zero padded
        t = torch.zeros(samples, H, tokens, HP, device="cuda", dtype=torch.bfloat16)
        t[..., :HD] = cols.reshape(samples, tokens, H, HD).permute(0, 2, 1, 3)
        return t

    q_hm = headmajor(qkv[:, :D], q_tokens)
    out_rm = torch.full((samples, q_tokens, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    out_hm = torch.full_like(out_rm, float("nan"))
    if cross:
        real = nk - 9
        k_hm = torch.zeros(samples, H, nk, HP, device="cuda", dtype=torch.bfloat16)
        v_hm = torch.zeros_like(k_hm)
        k_hm[:, :, :real, :HD] = _bf(torch.randn(samples, H, real, HD, device="cuda", generator=g) * 1.5)
        v_hm[:, :, :real, :HD] = _bf(torch.randn(samples, H, real, HD, device="cuda", generator=g))
        bias = torch.zeros(samples, nk, device="cuda")
        bias[:, real - 5:real] = -10000.0
        bias[:, real:] = float("-inf")
        _lib.attention_ex(qkv, 3 * D, k_hm, v_hm, 0, bias, out_rm, samples, H, q_tokens, nk)
    else:
        assert nk == q_tokens
        k_hm, v_hm, bias = headmajor(qkv[:, D:2 * D], nk), headmajor(qkv[:, 2 * D:], nk), None
        _lib.attention_ex(qkv, 3 * D, qkv[:, D:], qkv[:, 2 * D:], 3 * D, None, out_rm, samples, H, q_tokens, nk)
    _lib.attention(q_hm, k_hm, v_hm, bias, out_hm, samples, H, q_tokens, nk)
    torch.cuda.synchronize()
    assert torch.isfinite(out_rm.float()).all()
    assert torch.equal(out_rm, out_hm)
    ref = _attn_ref(q_hm[..., :HD], k_hm[..., :HD], v_hm[..., :HD], bias)
    assert _rel_err(out_rm, ref) < 2e-2


def test_attention_all_masked_row_is_uniform(cuda_device):
    """mask of all zeros -> every real key gets -10000 -> softmax is uniform over the REAL keys (reference
    semantics of the additive -10000 bias, pixart_transformer_2d_edited.py:282-289), padding keys excluded."""
    from ecad_b200 import _lib
    samples, nk, q_tokens, T = 1, 128, 128, 120
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.zeros(samples, H, q_tokens, HP, device="cuda", dtype=torch.bfloat16)
    k = torch.zeros(samples, H, nk, HP, device="cuda", dtype=torch.bfloat16)
    v = torch.zeros(samples, H, nk, HP, device="cuda", dtype=torch.bfloat16)
    v[:, :, :T, :HD] = _bf(torch.randn(samples, H, T, HD, device="cuda", generator=g))
    bias = torch.full((samples, nk), -10000.0, device="cuda")
    bias[:, T:] = float("-inf")
    out = torch.empty(samples, q_tokens, H * HD, device="cuda", dtype=torch.bfloat16)
    _lib.attention(q, k, v, bias, out, samples, H, q_tokens, nk)
    torch.cuda.synchronize()
    ref = v[:, :, :T, :HD].float().mean(dim=2)  # [S,H,72]
    ref = ref.reshape(samples, 1, H * HD).expand(samples, q_tokens, H * HD)
    assert _rel_err(out, ref) < 1.5e-2


@pytest.mark.parametrize("n_reuse,do_ln,do_xb", [(0, True, False), (3, True, False), (6, False, False),
                                                  (2, False, True), (1, True, True), (12, True, False),
                                                  (9, False, True)])
def test_residual_ln(cuda_device, n_reuse, do_ln, do_xb):
    from ecad_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(9)
    samples, tokens = 3, 256
    m = samples * tokens
    x = torch.randn(m, D, device="cuda", generator=g) * 3 + 0.5
    x0 = x.clone()
    table = torch.randn(6, D, device="cuda", generator=g) * 0.1
    temb = torch.randn(samples, 6 * D, device="cuda", generator=g) * 0.1
    caches = [_bf(torch.randn(m, D, device="cuda", generator=g)) for _ in range(n_reuse)]
    reuse = []
    x_ref = x0.clone()
    for j, c in enumerate(caches):
        if j % 3 == 1:  # attn2-style: no gate
            reuse.append((c, None, None))
            x_ref = x_ref + c.float()
        else:
            reuse.append((c, table[2], temb[:, 2 * D:]))
            gate = (table[2][None] + temb[:, 2 * D:3 * D]).repeat_interleave(tokens, dim=0)
            x_ref = x_ref + gate * c.float()
    h = torch.zeros(m, D, device="cuda", dtype=torch.bfloat16) if do_ln else None
    xb = torch.zeros(m, D, device="cuda", dtype=torch.bfloat16) if do_xb else None
    _lib.residual_ln(x, tokens, reuse=reuse, xb=xb, h=h, shift_table=table[0], scale_table=table[1],
                     shift_temb=temb[:, 0:], scale_temb=temb[:, D:], temb_stride=6 * D, eps=1e-6)
    torch.cuda.synchronize()
    assert _rel_err(x, x_ref) < 1e-6 if n_reuse else torch.equal(x, x0)
    if do_xb:
        assert torch.equal(xb, _bf(x))
    if do_ln:
        shift = (table[0][None] + temb[:, 0:D]).repeat_interleave(tokens, dim=0)
        scale = (table[1][None] + temb[:, D:2 * D]).repeat_interleave(tokens, dim=0)
        ref = torch.nn.functional.layer_norm(x_ref, (D,), eps=1e-6) * (1 + scale) + shift
        assert _rel_err(h, ref) < 6e-3  # bf16 output rounding


def test_patch_embed_and_final_layer(cuda_device):
    from ecad_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(13)
    S, Cc, Hl, Wl = 3, 4, 32, 32
    lat = torch.randn(S, Cc, Hl, Wl, device="cuda", generator=g)
    w = torch.randn(D, Cc, 2, 2, device="cuda", generator=g) * 0.25
    b = torch.randn(D, device="cuda", generator=g)
    N = (Hl // 2) * (Wl // 2)
    pos = torch.randn(N, D, device="cuda", generator=g)
    x = torch.empty(S * N, D, device="cuda")
    wt = w.reshape(D, Cc * 4).t().contiguous()
    lib = _lib.load()
    _lib.check(lib.ecadk_patch_embed(lat.data_ptr(), wt.data_ptr(), b.data_ptr(), pos.data_ptr(), x.data_ptr(), S, Cc,
                                     Hl, Wl, D, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(lat, w, b, stride=2).flatten(2).transpose(1, 2) + pos[None]
    assert _rel_err(x.view(S, N, D), ref) < 1e-5

    table = torch.randn(2, D, device="cuda", generator=g) * 0.1
    emb = torch.randn(S, D, device="cuda", generator=g) * 0.1
    wo = torch.randn(32, D, device="cuda", generator=g) / math.sqrt(D)
    bo = torch.randn(32, device="cuda", generator=g)
    out = torch.full((S, 8, Hl, Wl), float("nan"), device="cuda")
    w_pad = torch.zeros(128, D, device="cuda", dtype=torch.bfloat16)
    w_pad[:32] = wo.to(torch.bfloat16)
    b_pad = torch.zeros(128, device="cuda")
    b_pad[:32] = bo
    h_scr = torch.empty(S * N, D, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ecadk_final_layer(x.data_ptr(), table.data_ptr(), emb.data_ptr(), D, w_pad.data_ptr(),
                                     b_pad.data_ptr(), h_scr.data_ptr(), out.data_ptr(), S, Hl // 2, Wl // 2, D, 8,
                                     1e-6, _lib.stream_ptr()))
    torch.cuda.synchronize()
    shift, scale = (table[None] + emb[:, None]).chunk(2, dim=1)
    hsd = torch.nn.functional.layer_norm(x.view(S, N, D), (D,), eps=1e-6) * (1 + scale) + shift
    hsd = hsd @ wo.T + bo
    hsd = hsd.reshape(-1, Hl // 2, Wl // 2, 2, 2, 8)
    ref_out = torch.einsum("nhwpqc->nchpwq", hsd).reshape(-1, 8, Hl, Wl)
    assert torch.isfinite(out).all()  # every output element written
    # bf16 operands (h and W rounded to 8 mantissa bits), fp32 accumulation over 1152 terms
    assert _rel_err(out, ref_out) < 6e-3


def test_timestep_path_and_small_ops(cuda_device):
    from ecad_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(17)
    S = 5
    t = torch.tensor([999.0, 949.0, 500.0, 50.0, 0.0], device="cuda")
    proj = torch.empty(S, 256, device="cuda")
    _lib.check(lib.ecadk_timestep_sinusoid(t.data_ptr(), proj.data_ptr(), S, 256, _lib.stream_ptr()))
    f = torch.exp(-math.log(10000) * torch.arange(128, device="cuda", dtype=torch.float32) / 128)
    ang = t[:, None] * f[None]
    ref = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)
    torch.cuda.synchronize()
    assert float((proj - ref).abs().max()) < 2e-4  # fp32 sin/cos of arguments up to 999 rad

    w = torch.randn(300, 256, device="cuda", generator=g) / 16
    b = torch.randn(300, device="cuda", generator=g)
    y = torch.zeros(S, 400, device="cuda")
    _lib.check(lib.ecadk_small_linear(proj.data_ptr(), 256, w.data_ptr(), b.data_ptr(), y.data_ptr(), S, 256, 300, 400, 100,
                                      1, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref_y = torch.nn.functional.silu(proj) @ w.T + b
    assert _rel_err(y[:, 100:], ref_y) < 1e-5
    assert float(y[:, :100].abs().max()) == 0.0

    src = torch.randn(1024, 36, device="cuda", generator=g)
    dst = torch.empty(1024, 36, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.ecadk_cast_f32_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(dst, src.to(torch.bfloat16))

    mask = (torch.rand(S, 120, device="cuda", generator=g) > 0.5).float()
    bias = torch.empty(S, 128, device="cuda")
    _lib.check(lib.ecadk_mask_bias(mask.data_ptr(), bias.data_ptr(), S, 120, 128, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(bias[:, :120], (1 - mask) * -10000.0)
    assert torch.isinf(bias[:, 120:]).all() and (bias[:, 120:] < 0).all()


def test_bad_arguments_return_errors_not_crashes(cuda_device):
    from ecad_b200 import _lib
    a = torch.zeros(128, 100, device="cuda", dtype=torch.bfloat16)  # K=100 not a multiple of 64
    w = torch.zeros(128, 100, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(128, 128, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="multiple of 64"):
        _lib.gemm_bias(a, w, None, out)
    q = torch.zeros(1, 16, 128, 80, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="n_keys"):
        _lib.attention(q, q, q, None, out, 1, 16, 128, 64)
