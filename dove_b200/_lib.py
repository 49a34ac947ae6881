"""ctypes binding of libdove_b200.so (the C ABI declared in include/dove_b200.h).

PyTorch is used only for device memory and streams: every wrapper takes torch CUDA tensors, passes their raw
pointers and the current CUDA stream to the C entry point and raises `DoveError` on a non-zero return code.
There is no CPU fallback: loading fails loudly if the shared library is missing, and `init()` fails if there is
no sm_100 device.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libdove_b200.so"

EPI_BIAS, EPI_GELU_TANH, EPI_GATED_RES, EPI_ADD = 0, 1, 2, 3
DEFAULT_ATTN_VARIANT = -1     # auto: v4 kernel with 1/8 of the exps on the FMA pipe for >= 3000 rows, v2 kernel below


class DoveError(RuntimeError):
    pass


_SIGS = {
    "dove_abi_version": (c_int, []),
    "dove_init": (c_int, [c_int]),
    "dove_last_error": (c_char_p, []),
    "dove_num_sms": (c_int, []),
    "dove_set_option": (c_int, [c_char_p, c_int]),
    "dove_gemm_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                               c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "dove_gemv_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "dove_layernorm_mod_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dove_qk_norm_rope_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                       c_void_p, c_void_p, c_int, c_void_p]),
    "dove_gemm_qkv_norm_rope_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                             c_void_p, c_int, c_void_p]),
    "dove_attention_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "dove_patchify_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dove_unpatchify_velocity_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                              c_float, c_float, c_void_p]),
    "dove_velocity_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p]),
    "dove_conv_cl_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                  c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "dove_gn_finalize": (c_int, [c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p]),
    "dove_conv3d_causal_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "dove_gn_partial_floats": (c_size_t, [c_int64, c_int]),
    "dove_gn_stats_bf16": (c_int, [c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "dove_gn_apply_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "dove_time_pool_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "dove_upsample_nearest_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dove_pixels_to_cl_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dove_ncthw_to_cl_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "dove_cl_to_ncthw_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dove_gaussian_sample_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "dove_post_scale_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "dove_h2d_box_async": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "dove_upscale_normalize_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dove_blend_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int,
                                c_int64, c_int64, c_int64, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None
_inited_device = None


def load() -> ctypes.CDLL:
    """dlopen the in-tree library (no GPU needed); raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise DoveError(f"{LIB_PATH} is missing: run `python -m dove_b200.build` "
                            "(dove_b200 has no CPU / PyTorch fallback path)")
        lib = ctypes.CDLL(os.fspath(LIB_PATH))
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def init(device: int | None = None) -> None:
    """Bind the library to a CUDA device (idempotent)."""
    global _inited_device
    lib = load()
    if not torch.cuda.is_available():
        raise DoveError("no CUDA device: dove_b200 runs only on sm_100a GPUs (no CPU fallback)")
    if device is None:
        device = torch.cuda.current_device()
    if _inited_device == device:
        return
    rc = lib.dove_init(int(device))
    if rc != 0:
        raise DoveError(f"dove_init({device}) failed ({rc}): {lib.dove_last_error().decode()}")
    _inited_device = device


def set_option(name: str, value: int) -> None:
    rc = load().dove_set_option(name.encode(), int(value))
    if rc != 0:
        raise DoveError(load().dove_last_error().decode())


def _p(t):
    if t is None:
        return None
    assert t.is_cuda, "dove_b200 kernels take CUDA tensors only"
    return c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


launch_count = 0   # number of C-ABI compute calls issued (bench.py reports kernel launches from this)


_fn_cache = {}     # name -> bound foreign function (one attribute lookup per entry point, not per launch)


def _call(name, *args):
    global launch_count
    fn = _fn_cache.get(name)
    if fn is None:
        lib = load()
        if _inited_device is None:
            init()
        fn = _fn_cache[name] = getattr(lib, name)
    rc = fn(*args)
    launch_count += 1
    if rc != 0:
        raise DoveError(f"{name} failed ({rc}): {load().dove_last_error().decode()}")


def _bf16c(t):
    assert t.dtype == torch.bfloat16 and t.is_contiguous(), (t.dtype, t.is_contiguous())
    return t


# ---------------------------------------------------------------------------------------------- DiT ops
def gemm(a, w, out, bias=None, epilogue=EPI_BIAS, aux=None, gate0=None, gate1=None, split_row=0):
    """out[M,N] = epi(a[M,K] @ w[N,K]^T + bias).  a/out may be row-strided views (stride(1) == 1)."""
    assert a.dtype == w.dtype == out.dtype == torch.bfloat16 and a.stride(1) == 1 and out.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and w.stride(1) == 1 and tuple(out.shape) == (M, N)
    _call("dove_gemm_bf16", _p(a), a.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K, _p(bias),
          epilogue, _p(aux), aux.stride(0) if aux is not None else 0, _p(gate0), _p(gate1), split_row, _stream())
    return out


def gemv(x, w, b, out, silu_in=False):
    _call("dove_gemv_bf16", _p(_bf16c(x)), _p(_bf16c(w)), _p(b), _p(out), w.shape[0], w.shape[1], int(silu_in),
          _stream())
    return out


def layernorm_mod(x, out, ln_w, ln_b, eps, scale0=None, shift0=None, scale1=None, shift1=None, split_row=0):
    rows, D = x.shape
    _call("dove_layernorm_mod_bf16", _p(_bf16c(x)), _p(_bf16c(out)), rows, D, _p(ln_w), _p(ln_b), eps, _p(scale0),
          _p(shift0), _p(scale1), _p(shift1), split_row, _stream())
    return out


def qk_norm_rope(qkv, heads, q_w, q_b, k_w, k_b, eps, cos, sin, text_len):
    rows = qkv.shape[0]
    assert qkv.shape[1] == 3 * heads * 64
    if cos is not None:
        assert cos.dtype == torch.float32 and cos.is_contiguous() and cos.shape == (rows - text_len, 64)
    _call("dove_qk_norm_rope_bf16", _p(_bf16c(qkv)), rows, heads, _p(q_w), _p(q_b), _p(k_w), _p(k_b), eps, _p(cos),
          _p(sin), text_len, _stream())
    return qkv


def rope_tables_transposed(cos, sin):
    """[nv, 64] pair-duplicated RoPE tables -> ([32, nv], [32, nv]) fp32 contiguous, the layout of the fused QKV epilogue.
    Raises if the tables are not pair-duplicated (get_3d_rotary_pos_embed repeats every frequency twice)."""
    if not (torch.equal(cos[:, 0::2], cos[:, 1::2]) and torch.equal(sin[:, 0::2], sin[:, 1::2])):
        raise DoveError("fused QKV epilogue needs pair-duplicated RoPE tables (interleaved real layout)")
    return cos[:, 0::2].t().contiguous(), sin[:, 0::2].t().contiguous()


def gemm_qkv_norm_rope(a, w, out, bias, heads, q_w, q_b, k_w, k_b, eps, cos_t, sin_t, text_len):
    """out[M, 3*heads*64] = qkv projection with per-head q/k LayerNorm + RoPE fused in the GEMM epilogue.
    cos_t / sin_t: `rope_tables_transposed(cos, sin)` ([32, M - text_len] fp32) or None when no row is rotated."""
    M, K = a.shape
    assert a.dtype == w.dtype == out.dtype == torch.bfloat16 and a.stride(1) == 1 and out.stride(1) == 1
    assert tuple(w.shape) == (3 * heads * 64, K) and tuple(out.shape) == (M, 3 * heads * 64)
    if cos_t is not None:
        assert cos_t.dtype == torch.float32 and cos_t.is_contiguous() and cos_t.shape == (32, M - text_len)
        assert sin_t.dtype == torch.float32 and sin_t.is_contiguous() and sin_t.shape == cos_t.shape
    _call("dove_gemm_qkv_norm_rope_bf16", _p(a), a.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, heads, K,
          _p(_bf16c(bias)), _p(q_w), _p(q_b), _p(k_w), _p(k_b), eps, _p(cos_t), _p(sin_t), text_len, _stream())
    return out


def attention(qkv, out, heads, scale):
    rows = qkv.shape[0]
    assert qkv.shape[1] == 3 * heads * 64 and tuple(out.shape) == (rows, heads * 64)
    _call("dove_attention_bf16", _p(_bf16c(qkv)), _p(_bf16c(out)), rows, heads, scale, _stream())
    return out


def h2d_box(src_view, dst):
    """src_view: a [planes, rows, cols] strided view (unit stride in cols) of a HOST tensor (pinned for a truly
    asynchronous copy); dst: a dense CUDA tensor of the same shape and dtype.  One DMA, no staging buffer."""
    planes, rows, cols = src_view.shape
    assert not src_view.is_cuda and dst.is_cuda and dst.is_contiguous() and dst.dtype == src_view.dtype
    assert tuple(dst.shape) == (planes, rows, cols)
    sp, sr, sc = src_view.stride()
    eb = src_view.element_size()
    assert sc == 1 and sr >= cols and (planes == 1 or (sp % sr == 0 and sp // sr >= rows)), "unsupported source strides"
    _call("dove_h2d_box_async", c_void_p(src_view.data_ptr()), sr * eb, (sp // sr) if planes > 1 else rows,
          _p(dst), cols * eb, rows, planes, _stream())
    return dst


def patchify(latent, tokens):
    F, C, h, w = latent.shape
    _call("dove_patchify_bf16", _p(_bf16c(latent)), _p(_bf16c(tokens)), F, C, h, w, _stream())
    return tokens


def unpatchify_velocity(tokens, latent, x0, pred, F, C, h, w, a, b):
    _call("dove_unpatchify_velocity_bf16", _p(_bf16c(tokens)), _p(latent), _p(x0), _p(pred), F, C, h, w, a, b,
          _stream())


def velocity(sample, noise, out, a, b):
    _call("dove_velocity_bf16", _p(_bf16c(sample)), _p(_bf16c(noise)), _p(_bf16c(out)), sample.numel(), a, b,
          _stream())
    return out


# ---------------------------------------------------------------------------------------------- VAE ops
OUT_CL, OUT_PLANAR, OUT_PLANAR_POST, OUT_PLANAR_U8 = 0, 1, 2, 3


def conv_cl(x, w, bias, y, Tout, kt, kh, kw, stride, pad, Ho, Wo, cout_valid, epilogue=EPI_BIAS, aux=None,
            out_mode=0, plane_stride=0, gn_partial=None):
    """x [Tin,Hin,Win,Cin] channels-last; w [Cout_pad, kt*kh*kw*Cin]; y [Tout,Ho,Wo,ldy] (or planar)."""
    Tin, Hin, Win, Cin = x.shape
    assert Tin == Tout + kt - 1, (Tin, Tout, kt)
    Cout_pad = w.shape[0]
    ldy = y.shape[-1] if out_mode == 0 else (plane_stride or Tout * Ho * Wo)
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    assert y.dtype == (torch.uint8 if out_mode == OUT_PLANAR_U8 else torch.bfloat16)
    gn_done = ctypes.c_int(0)          # per call: the C entry point writes it before returning
    _call("dove_conv_cl_bf16", _p(x), _p(_bf16c(w)), _p(bias), _p(y), Tout, Hin, Win, Cin, Cout_pad,
          cout_valid, ldy, kt, kh, kw, stride, pad, Ho, Wo, epilogue, _p(aux),
          aux.shape[-1] if aux is not None else 0, out_mode, _p(gn_partial),
          ctypes.byref(gn_done) if gn_partial is not None else None, _stream())
    return (y, bool(gn_done.value)) if gn_partial is not None else y


def conv3d_causal(x, x_prev, w, bias, y, cout_valid, epilogue=EPI_BIAS, aux=None, out_mode=0, plane_stride=0,
                  gn_partial=None):
    """Causal 3x3x3 conv on the un-padded frame batch x [T,H,W,Cin]; x_prev [2,H,W,Cin] (cache view) or None."""
    T, H, W, Cin = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    assert y.dtype == (torch.uint8 if out_mode == OUT_PLANAR_U8 else torch.bfloat16)
    if x_prev is not None:
        assert x_prev.shape == (2, H, W, Cin) and x_prev.is_contiguous() and x_prev.dtype == torch.bfloat16
    ldy = y.shape[-1] if out_mode == 0 else (plane_stride or T * H * W)
    gn_done = ctypes.c_int(0)
    _call("dove_conv3d_causal_bf16", _p(x), _p(x_prev), _p(_bf16c(w)), _p(bias), _p(y), T, H, W, Cin, w.shape[0],
          cout_valid, ldy, epilogue, _p(aux), aux.shape[-1] if aux is not None else 0, out_mode, _p(gn_partial),
          ctypes.byref(gn_done) if gn_partial is not None else None, _stream())
    return (y, bool(gn_done.value)) if gn_partial is not None else y


def gn_finalize(partial, nvox, C, groups, eps, stats):
    _call("dove_gn_finalize", _p(partial), nvox, C, groups, eps, _p(stats), _stream())
    return stats


def gn_partial_floats(nvox, groups=32):
    return int(load().dove_gn_partial_floats(nvox, groups))


def gn_stats(x, C, groups, eps, partial, stats):
    nvox = x.numel() // C
    _call("dove_gn_stats_bf16", _p(_bf16c(x)), nvox, C, groups, eps, _p(partial), _p(stats), _stream())
    return stats


def gn_apply(x, out, T, H, W, C, groups, stats, gamma, beta, silu, zq_y=None, zq_b=None):
    Tz = hz = wz = 0
    if zq_y is not None:
        Tz, hz, wz, _ = zq_y.shape
    _call("dove_gn_apply_bf16", _p(x), _p(out), T, H, W, C, groups, _p(stats), _p(gamma), _p(beta), int(silu),
          _p(zq_y), _p(zq_b), Tz, hz, wz, _stream())
    return out


def time_pool(x, y, T, frame_elems):
    _call("dove_time_pool_bf16", _p(x), _p(y), T, frame_elems, _stream())
    return y


def upsample_nearest(x, y, T, H, W, C, time_x2):
    _call("dove_upsample_nearest_bf16", _p(x), _p(y), T, H, W, C, int(time_x2), _stream())
    return y


def pixels_to_cl(x, y, T, H, W, Cpad):
    assert x.dtype in (torch.float32, torch.bfloat16) and x.is_contiguous()
    _call("dove_pixels_to_cl_bf16", _p(x), int(x.dtype == torch.float32), _p(y), T, H, W, Cpad, _stream())
    return y


def ncthw_to_cl(x, y, C, T, H, W, Cpad, scale=1.0):
    _call("dove_ncthw_to_cl_bf16", _p(_bf16c(x)), _p(y), C, T, H, W, Cpad, scale, _stream())
    return y


def cl_to_ncthw(x, y, C, T, H, W, ldx):
    _call("dove_cl_to_ncthw_bf16", _p(x), _p(y), C, T, H, W, ldx, _stream())
    return y


def gaussian_sample(moments, noise, z, nvox, scaling):
    _call("dove_gaussian_sample_bf16", _p(moments), _p(_bf16c(noise)), _p(z), nvox, scaling, _stream())
    return z


def post_scale(x, y):
    _call("dove_post_scale_bf16", _p(_bf16c(x)), _p(y), x.numel(), _stream())
    return y


def blend(a, b, axis, extent, layout):
    """In-place linear blend of tile `b` with its neighbour `a` along `axis` ("v" rows / "h" cols).
    layout "cl": tensors [T, Y, X, C];  layout "planar": tensors [C, T, Y, X]."""
    if layout == "cl":
        T, Ya, Xa, C = a.shape
        _, Yb, Xb, _ = b.shape
        outer, inner = T, C
        a_s = (a.stride(0), a.stride(1), a.stride(2))
        b_s = (b.stride(0), b.stride(1), b.stride(2))
    else:
        Cc, T, Ya, Xa = a.shape
        _, _, Yb, Xb = b.shape
        assert a.stride(1) == Ya * Xa and a.stride(0) == T * Ya * Xa and b.stride(1) == Yb * Xb
        outer, inner = Cc * T, 1
        a_s = (Ya * Xa, a.stride(2), a.stride(3))
        b_s = (Yb * Xb, b.stride(2), b.stride(3))
    if axis == "v":
        e = min(Ya, Yb, extent)
        assert Xa == Xb
        args = (outer, e, Xb, inner, a_s[0], a_s[1], a_s[2], Ya, b_s[0], b_s[1], b_s[2])
    else:
        e = min(Xa, Xb, extent)
        assert Ya == Yb
        args = (outer, e, Yb, inner, a_s[0], a_s[2], a_s[1], Xa, b_s[0], b_s[2], b_s[1])
    if e > 0:
        _call("dove_blend_bf16", _p(a), _p(b), *args, _stream())
    return b


def upscale_normalize(lr, out, scale):
    F, C, h, w = lr.shape
    assert C == 3 and lr.dtype == torch.float32 and lr.is_contiguous() and out.dtype == torch.float32
    _call("dove_upscale_normalize_f32", _p(lr), _p(out), F, h, w, scale, _stream())
    return out
