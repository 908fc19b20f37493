"""ctypes binding of libecad_b200.so (C ABI: include/ecad_b200.h).

The product path has NO fallback: if the shared library is missing or the device is not sm_100 the calls raise.
PyTorch is used only for device memory and the stream handle.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "libecad_b200.so"
ABI_VERSION = 2  # include/ecad_b200.h: ECADK_ABI_VERSION (2: EcadkBlocksArgs.self_bias)
MAX_REUSE = 12
HEAD_DIM = 72
HEAD_PAD = 80

EXPORTED_SYMBOLS = [
    "ecadk_abi_version",
    "ecadk_last_error",
    "ecadk_device_check",
    "ecadk_residual_ln",
    "ecadk_patch_embed",
    "ecadk_patch_embed_padded",
    "ecadk_timestep_sinusoid",
    "ecadk_small_linear",
    "ecadk_cast_f32_bf16",
    "ecadk_mask_bias",
    "ecadk_average_halves",
    "ecadk_final_layer",
    "ecadk_final_layer_padded",
    "ecadk_cfg_dpm_step",
    "ecadk_set_splitk_workspace",
    "ecadk_splitk_launches",
    "ecadk_set_splitk_force",
    "ecadk_gemm_bias",
    "ecadk_gemm_bias_gated_residual_cache",
    "ecadk_gemm_bias_headmajor",
    "ecadk_attention",
    "ecadk_attention_ex",
    "ecadk_attention_d128",
    "ecadk_qk_norm_rope",
    "ecadk_qk_norm_rope_batched",
    "ecadk_strided_unary",
    "ecadk_axpy_f32",
    "ecadk_gemm_bias_f32",
    "ecadk_gemm_bias_headmajor_ex",
    "ecadk_gemm2src_gated_residual_cache",
    "ecadk_gemm_bias_dual",
    "ecadk_create",
    "ecadk_destroy",
    "ecadk_pixart_blocks",
    "ecadk_pixart_blocks_range",
    "ecadk_pixart_text_kv",
    "ecadk_silu_f32_bf16",
    "ecadk_flux_create",
    "ecadk_flux_destroy",
    "ecadk_flux_blocks",
    "ecadk_conv_nhwc",
    "ecadk_conv_up2x_nhwc",
    "ecadk_groupnorm_nhwc",
    "ecadk_groupnorm_scratch_bytes",
    "ecadk_upsample2x_nhwc",
    "ecadk_softmax_rows",
    "ecadk_vae_prepare_latents",
    "ecadk_vae_add_tokens",
    "ecadk_vae_finish",
    "ecadk_profile_start",
    "ecadk_profile_stop",
    "ecadk_set_nvtx",
    "ecadk_nvtx_ranges",
]


class EcadkReuse(C.Structure):
    _fields_ = [("cache", C.c_void_p), ("gate_table", C.c_void_p), ("gate_temb", C.c_void_p)]


class EcadkResidualLnArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("xb", C.c_void_p),
        ("h", C.c_void_p),
        ("rows", C.c_int),
        ("tokens", C.c_int),
        ("dim", C.c_int),
        ("n_reuse", C.c_int),
        ("reuse", EcadkReuse * MAX_REUSE),
        ("shift_table", C.c_void_p),
        ("scale_table", C.c_void_p),
        ("shift_temb", C.c_void_p),
        ("scale_temb", C.c_void_p),
        ("temb_stride", C.c_int),
        ("eps", C.c_float),
    ]


class EcadkModelDesc(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int),
        ("dim", C.c_int),
        ("heads", C.c_int),
        ("ff_dim", C.c_int),
        ("norm_eps", C.c_float),
    ]


class EcadkBlockWeights(C.Structure):
    _fields_ = [
        (n, C.c_void_p)
        for n in (
            "w_qkv1", "b_qkv1", "w_out1", "b_out1", "w_q2", "b_q2", "w_kv2", "b_kv2", "w_out2", "b_out2",
            "w_ff1", "b_ff1", "w_ff2", "b_ff2", "scale_shift_table",
        )
    ]


class EcadkBlocksArgs(C.Structure):
    _fields_ = [
        ("samples", C.c_int),
        ("tokens", C.c_int),
        ("text_pad", C.c_int),
        ("x", C.c_void_p),
        ("xb", C.c_void_p),
        ("h", C.c_void_p),
        ("q", C.c_void_p),
        ("k", C.c_void_p),
        ("v", C.c_void_p),
        ("attn_o", C.c_void_p),
        ("ffh", C.c_void_p),
        ("temb6", C.c_void_p),
        ("temb_stride", C.c_int),
        ("text_bias", C.c_void_p),
        ("k2", C.POINTER(C.c_void_p)),
        ("v2", C.POINTER(C.c_void_p)),
        ("cache", C.POINTER(C.c_void_p)),
        ("cache_dead", C.POINTER(C.c_uint8)),
        ("qkv", C.c_void_p),
        ("self_bias", C.c_void_p),
    ]


class EcadkFluxDesc(C.Structure):
    _fields_ = [("num_layers", C.c_int), ("num_single_layers", C.c_int), ("dim", C.c_int), ("heads", C.c_int),
                ("eps", C.c_float)]


class EcadkFluxDoubleWeights(C.Structure):
    _fields_ = [
        (n, C.c_void_p)
        for n in (
            "w_qkv", "b_qkv", "w_qkv_ctx", "b_qkv_ctx", "w_out", "b_out", "w_out_ctx", "b_out_ctx", "w_ff1", "b_ff1",
            "w_ff2", "b_ff2", "w_ff1_ctx", "b_ff1_ctx", "w_ff2_ctx", "b_ff2_ctx", "norm_q", "norm_k", "norm_added_q",
            "norm_added_k",
        )
    ]


class EcadkFluxSingleWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w_qkv", "b_qkv", "w_mlp", "b_mlp", "w_out", "b_out", "norm_q", "norm_k")]


class EcadkFluxArgs(C.Structure):
    _fields_ = (
        [("samples", C.c_int), ("img_tokens", C.c_int), ("txt_tokens", C.c_int)]
        + [(n, C.c_void_p) for n in ("x_img", "x_txt", "x_cat", "h_img", "h_txt", "h_cat", "q", "k", "v", "attn_img",
                                     "attn_txt", "ffh", "cat", "mod")]
        + [("mod_stride", C.c_int), ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
           ("cache_double", C.POINTER(C.c_void_p)), ("cache_single", C.POINTER(C.c_void_p)),
           ("cache_dead", C.POINTER(C.c_uint8)), ("rope_sample_stride", C.c_int)]
    )


class EcadkProfileRecord(C.Structure):
    _fields_ = [("launches", C.c_longlong), ("total_ms", C.c_double), ("flops", C.c_double), ("bytes", C.c_double)]


PROF_CLASSES = ("gemm", "attention", "glue", "other")

_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """dlopen the in-tree library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("ECAD_B200_LIB", _LIB_PATH))  # instrumented / A-B builds of the same ABI
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the ECAD B200 hot path)"
        )
    lib = C.CDLL(os.fspath(path))
    lib.ecadk_abi_version.restype = C.c_int
    lib.ecadk_last_error.restype = C.c_char_p
    i, f, p, sz = C.c_int, C.c_float, C.c_void_p, C.c_size_t
    sigs = {
        "ecadk_device_check": [i],
        "ecadk_residual_ln": [C.POINTER(EcadkResidualLnArgs), p],
        "ecadk_patch_embed": [p, p, p, p, p, i, i, i, i, i, p],
        "ecadk_patch_embed_padded": [p, p, p, p, p, i, i, i, i, i, i, p],
        "ecadk_timestep_sinusoid": [p, p, i, i, p],
        "ecadk_small_linear": [p, i, p, p, p, i, i, i, i, i, i, i, p],
        "ecadk_cast_f32_bf16": [p, p, sz, p],
        "ecadk_mask_bias": [p, p, i, i, i, p],
        "ecadk_average_halves": [p, sz, p],
        "ecadk_final_layer": [p, p, p, i, p, p, p, p, i, i, i, i, i, f, p],
        "ecadk_final_layer_padded": [p, p, p, i, p, p, p, p, i, i, i, i, i, i, f, p],
        "ecadk_cfg_dpm_step": [p, p, p, i, i, i, i, f, f, f, f, f, f, p],
        "ecadk_set_splitk_workspace": [p, sz],
        "ecadk_splitk_launches": [],
        "ecadk_set_splitk_force": [i, i],
        "ecadk_gemm_bias": [p, p, p, p, i, i, i, i, i, p],
        "ecadk_gemm_bias_gated_residual_cache": [p, p, p, p, p, p, p, p, i, i, i, i, i, p],
        "ecadk_gemm_bias_headmajor": [p, p, p, p, p, p, i, i, i, i, i, i, p],
        "ecadk_attention": [p, p, p, p, p, i, i, i, i, p],
        "ecadk_attention_ex": [p, i, p, p, i, p, p, i, i, i, i, p],
        "ecadk_attention_d128": [p, p, p, p, i, p, i, i, i, i, i, p],
        "ecadk_qk_norm_rope": [p, p, p, p, p, p, p, p, i, i, i, i, f, p],
        "ecadk_qk_norm_rope_batched": [p, p, p, p, p, p, p, p, i, i, i, i, i, f, p],
        "ecadk_strided_unary": [p, p, i, i, i, i, i, p],
        "ecadk_axpy_f32": [p, p, f, sz, p],
        "ecadk_gemm_bias_f32": [p, p, p, p, i, i, i, i, i, p],
        "ecadk_gemm_bias_headmajor_ex": [p, p, p, p, p, p, i, i, i, i, i, i, i, i, i, p],
        "ecadk_gemm2src_gated_residual_cache": [p, i, p, p, p, p, p, p, i, i, i, i, i, p],
        "ecadk_gemm_bias_dual": [p, p, p, p, p, i, i, i, i, i, p],
        "ecadk_create": [i, C.POINTER(EcadkModelDesc), C.POINTER(EcadkBlockWeights), C.POINTER(p)],
        "ecadk_destroy": [p],
        "ecadk_pixart_blocks": [p, C.POINTER(EcadkBlocksArgs), C.POINTER(C.c_uint8), C.POINTER(i), p],
        "ecadk_pixart_blocks_range": [p, C.POINTER(EcadkBlocksArgs), C.POINTER(C.c_uint8), i, i, C.POINTER(i), p],
        "ecadk_pixart_text_kv": [p, p, i, i, i, C.POINTER(p), C.POINTER(p), C.POINTER(i), p],
        "ecadk_silu_f32_bf16": [p, p, sz, p],
        "ecadk_flux_create": [i, C.POINTER(EcadkFluxDesc), C.POINTER(EcadkFluxDoubleWeights),
                              C.POINTER(EcadkFluxSingleWeights), C.POINTER(p)],
        "ecadk_flux_destroy": [p],
        "ecadk_flux_blocks": [p, C.POINTER(EcadkFluxArgs), C.POINTER(C.c_uint8), C.POINTER(i), p],
        "ecadk_conv_nhwc": [p, p, p, p, p, i, i, i, i, i, i, i, i, p],
        "ecadk_conv_up2x_nhwc": [p, p, p, p, i, i, i, i, i, p],
        "ecadk_groupnorm_nhwc": [p, p, p, p, p, i, i, i, i, i, f, i, i, p],
        "ecadk_groupnorm_scratch_bytes": [i, i, i, i],
        "ecadk_upsample2x_nhwc": [p, p, i, i, i, i, p],
        "ecadk_softmax_rows": [p, p, i, i, i, f, p],
        "ecadk_vae_prepare_latents": [p, p, p, f, f, p, i, i, i, i, p],
        "ecadk_vae_add_tokens": [p, p, i, p, i, i, i, i, p],
        "ecadk_vae_finish": [p, p, i, i, i, i, p],
        "ecadk_profile_start": [],
        "ecadk_profile_stop": [C.POINTER(EcadkProfileRecord)],
        "ecadk_set_nvtx": [i],
        "ecadk_nvtx_ranges": [],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.ecadk_splitk_launches.restype = C.c_longlong
    lib.ecadk_groupnorm_scratch_bytes.restype = C.c_size_t
    lib.ecadk_nvtx_ranges.restype = C.c_longlong
    if lib.ecadk_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libecad_b200.so ABI version {lib.ecadk_abi_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ecadk_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libecad_b200 {what} failed (code {rc}): {msg}")


def ptr(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


# ---------------------------------------------------------------------------------------------------
# thin tensor-level wrappers (shape checking happens in C; these only pass pointers)
# ---------------------------------------------------------------------------------------------------
def gemm_bias(a, w, bias, out, gelu=False):
    m, k = a.shape
    n = w.shape[0]
    check(load().ecadk_gemm_bias(ptr(a), ptr(w), ptr(bias), ptr(out), m, n, k, out.stride(0), int(gelu), stream_ptr()),
          "gemm_bias")
    return out


def gemm_gated_residual(a, w, bias, x, cache, tokens, xb=None, gate_table=None, gate_temb=None, temb_stride=0):
    m, k = a.shape
    n = w.shape[0]
    check(
        load().ecadk_gemm_bias_gated_residual_cache(
            ptr(a), ptr(w), ptr(bias), ptr(x), ptr(xb), ptr(cache), ptr(gate_table), ptr(gate_temb), temb_stride,
            tokens, m, n, k, stream_ptr()),
        "gemm_bias_gated_residual_cache")


def gemm_headmajor(a, w, bias, outs, heads, tokens, tokens_pad):
    m, k = a.shape
    o = list(outs) + [None] * (3 - len(outs))
    check(
        load().ecadk_gemm_bias_headmajor(ptr(a), ptr(w), ptr(bias), ptr(o[0]), ptr(o[1]), ptr(o[2]), len(outs), heads,
                                         tokens, tokens_pad, m, k, stream_ptr()),
        "gemm_bias_headmajor")


def attention(q, k, v, bias, out, samples, heads, q_tokens, n_keys):
    check(load().ecadk_attention(ptr(q), ptr(k), ptr(v), ptr(bias), ptr(out), samples, heads, q_tokens, n_keys,
                                 stream_ptr()), "attention")
    return out


def conv_nhwc(x, w, bias, out, h, w_, taps, residual=None, out_cols=None):
    """x bordered NHWC bf16 [B, h+2, w_+2, c_in]; w bf16 [c_out, taps*c_in]; out [B, h+2, w_+2, out_ld]."""
    batch, c_in = x.shape[0], x.shape[-1]
    c_out = w.shape[0]
    check(load().ecadk_conv_nhwc(ptr(x), ptr(w), ptr(bias), ptr(residual), ptr(out), batch, h, w_, c_in, c_out,
                                 out.shape[-1], c_out if out_cols is None else out_cols, taps, stream_ptr()), "conv_nhwc")
    return out


def conv_up2x_nhwc(x, w4, bias, out, h, w_):
    """nearest-2x upsample + 3x3 conv as four 4-tap convolutions of x; w4 [4, c_out, 4*c_in] (see the header)."""
    check(load().ecadk_conv_up2x_nhwc(ptr(x), ptr(w4), ptr(bias), ptr(out), x.shape[0], h, w_, x.shape[-1], w4.shape[1],
                                      stream_ptr()), "conv_up2x_nhwc")
    return out


def groupnorm_nhwc(x, gamma, beta, out, scratch, h, w_, groups=32, eps=1e-6, silu=True, unpadded_out=False):
    """unpadded_out: False -> bordered NHWC out; True -> tokens [B, out.shape[1], C] (out.shape[1] >= h*w_)."""
    batch, c = x.shape[0], x.shape[-1]
    check(load().ecadk_groupnorm_nhwc(ptr(x), ptr(gamma), ptr(beta), ptr(out), ptr(scratch), batch, h, w_, c, groups,
                                      eps, int(silu), out.shape[1] if unpadded_out else 0, stream_ptr()), "groupnorm_nhwc")
    return out


def groupnorm_scratch_bytes(batch, h, w_, groups=32) -> int:
    return int(load().ecadk_groupnorm_scratch_bytes(batch, h, w_, groups))


def upsample2x_nhwc(x, out, h, w_):
    check(load().ecadk_upsample2x_nhwc(ptr(x), ptr(out), x.shape[0], h, w_, x.shape[-1], stream_ptr()), "upsample2x_nhwc")
    return out


def softmax_rows(scores, probs, scale, valid_cols=None):
    rows, cols = scores.shape
    check(load().ecadk_softmax_rows(ptr(scores), ptr(probs), rows, cols, cols if valid_cols is None else valid_cols,
                                    scale, stream_ptr()), "softmax_rows")
    return probs


def vae_prepare_latents(z, pq_w, pq_b, inv_scaling, out, shift=0.0):
    batch, cl, h, w_ = z.shape
    check(load().ecadk_vae_prepare_latents(ptr(z), ptr(pq_w), ptr(pq_b), inv_scaling, shift, ptr(out), batch, cl, h, w_,
                                           stream_ptr()), "vae_prepare_latents")
    return out


def vae_add_tokens(x, tokens, out, h, w_):
    """tokens [B, T, C] with T >= h*w_ (rows past h*w_ are padding)."""
    check(load().ecadk_vae_add_tokens(ptr(x), ptr(tokens), tokens.shape[1], ptr(out), x.shape[0], h, w_, x.shape[-1],
                                      stream_ptr()), "vae_add_tokens")
    return out


def vae_finish(y, image, h, w_, denormalize=False):
    check(load().ecadk_vae_finish(ptr(y), ptr(image), image.shape[0], h, w_, int(denormalize), stream_ptr()), "vae_finish")
    return image


def set_splitk_workspace(buf):
    """Install ``buf`` (a CUDA tensor, or None to remove it) as the split-K workspace of this thread's stand-alone GEMM
    calls (see ecadk_set_splitk_workspace).  The caller keeps the tensor alive while it is installed."""
    if buf is None:
        check(load().ecadk_set_splitk_workspace(None, 0), "set_splitk_workspace")
    else:
        check(load().ecadk_set_splitk_workspace(ptr(buf), buf.numel() * buf.element_size()), "set_splitk_workspace")


def set_splitk_force(bn: int = 0, split: int = 0) -> None:
    """Force the split-K plan (tests / measurements); (0, 0) restores the planner."""
    check(load().ecadk_set_splitk_force(bn, split), "set_splitk_force")


def splitk_launches() -> int:
    return int(load().ecadk_splitk_launches())


def attention_ex(q, q_ld, k, v, kv_ld, bias, out, samples, heads, q_tokens, n_keys):
    """q_ld / kv_ld = 0: head-major operands; > 0: row-major [samples*tokens, ld] (see ecadk_attention_ex)."""
    check(load().ecadk_attention_ex(ptr(q), q_ld, ptr(k), ptr(v), kv_ld, ptr(bias), ptr(out), samples, heads, q_tokens,
                                    n_keys, stream_ptr()), "attention_ex")
    return out


def residual_ln(x, tokens, reuse=(), xb=None, h=None, shift_table=None, scale_table=None, shift_temb=None,
                scale_temb=None, temb_stride=0, eps=1e-6):
    """reuse: iterable of (cache, gate_table|None, gate_temb|None) tensors."""
    a = EcadkResidualLnArgs()
    a.x, a.xb, a.h = ptr(x), ptr(xb), ptr(h)
    a.rows, a.dim = x.shape[0], x.shape[1]
    a.tokens = tokens
    reuse = list(reuse)
    a.n_reuse = len(reuse)
    for j, (c, gt, ge) in enumerate(reuse):
        a.reuse[j].cache, a.reuse[j].gate_table, a.reuse[j].gate_temb = ptr(c), ptr(gt), ptr(ge)
    a.shift_table, a.scale_table = ptr(shift_table), ptr(scale_table)
    a.shift_temb, a.scale_temb = ptr(shift_temb), ptr(scale_temb)
    a.temb_stride, a.eps = temb_stride, eps
    check(load().ecadk_residual_ln(C.byref(a), stream_ptr()), "residual_ln")


def profile_start() -> None:
    check(load().ecadk_profile_start(), "profile_start")


def profile_stop() -> dict[str, dict[str, float]]:
    """Per kernel class: launches, summed CUDA-event ms, algorithmic flops / bytes of those launches."""
    recs = (EcadkProfileRecord * len(PROF_CLASSES))()
    check(load().ecadk_profile_stop(recs), "profile_stop")
    return {n: {"launches": int(r.launches), "total_ms": r.total_ms, "flops": r.flops, "bytes": r.bytes}
            for n, r in zip(PROF_CLASSES, recs)}


# ---------------------------------------------------------------------------------------------------
# NVTX.  The library names one range per executor call and per executed sub-block
# (include/ecad_b200.h: ecadk_set_nvtx); the host side adds the levels above it - denoising step, transformer forward,
# VAE decoder stage - through torch's bundled NVTX.  Everything is off unless ECADK_NVTX=1 or set_nvtx(True).
# ---------------------------------------------------------------------------------------------------
_nvtx_on = os.environ.get("ECADK_NVTX", "0") not in ("", "0")


def set_nvtx(on: bool) -> None:
    global _nvtx_on
    _nvtx_on = bool(on)
    check(load().ecadk_set_nvtx(int(_nvtx_on)), "set_nvtx")


def nvtx_enabled() -> bool:
    return _nvtx_on


def nvtx_ranges() -> int:
    """Ranges the library has pushed so far."""
    return int(load().ecadk_nvtx_ranges())


class nvtx_range:
    """``with nvtx_range("step 03"):`` - a host-side NVTX range when tracing is on, nothing otherwise."""

    __slots__ = ("name", "live")

    def __init__(self, name: str):
        self.name = name
        self.live = False

    def __enter__(self):
        if _nvtx_on:
            torch.cuda.nvtx.range_push(self.name)
            self.live = True
        return self

    def __exit__(self, *exc):
        if self.live:
            torch.cuda.nvtx.range_pop()
        return False
