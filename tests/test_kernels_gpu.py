"""GPU parity tests of every C-ABI kernel against the fp32 arithmetic of the corresponding reference op
(torch fp32 on the same bf16 inputs, TF32 off), through the ctypes binding.  Tolerance: rel-L2 <= 2e-3 vs the
fp32 result rounded to bf16 (one bf16 ulp is 3.9e-3; identical rounding points give ~1e-5..1e-3)."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import bf, rb, rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-3


@pytest.fixture(scope="module")
def L():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from dove_b200 import _lib
    _lib.init(0)
    return _lib


def randn(*s, std=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return bf(torch.randn(*s, generator=g, device="cuda") * std)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 256, 128), (129, 64, 64), (200, 16, 64), (777, 32, 192),
                                   (1000, 128, 3072), (2500, 3072, 1024), (1111, 768, 256),
                                   (4097, 512, 256), (2304, 256, 64)])   # M >= 2048, N % 256 == 0: CTA-pair kernel
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm(L, M, N, K, epi):
    a = randn(M, K, seed=1)
    w = randn(N, K, std=K ** -0.5, seed=2)
    bias = randn(N, std=0.1, seed=3)
    aux = randn(M, N, seed=4)
    g0, g1 = randn(N, std=0.5, seed=5), randn(N, std=0.5, seed=6)
    split = M // 3
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.gemm(a, w, out, bias, epi, aux if epi >= 2 else None, g0 if epi == 2 else None, g1 if epi == 2 else None, split)
    torch.cuda.synchronize()
    r = rb(a.float() @ w.float().t() + bias.float())
    if epi == 1:
        ref = F.gelu(r, approximate="tanh")
    elif epi == 2:
        gate = torch.cat([g0.float().expand(split, N), g1.float().expand(M - split, N)], 0)
        ref = aux.float() + rb(gate * r)
    elif epi == 3:
        ref = r + aux.float()
    else:
        ref = r
    assert torch.isfinite(out.float()).all()
    e = rel_l2(out, rb(ref))
    print(f"gemm M{M} N{N} K{K} epi{epi}: rel_l2={e:.3e}")
    assert e < TOL


def test_gemm_strided(L):
    """A and C as column slices of wider buffers (the fused QKV / token layouts)."""
    M, N, K = 500, 256, 128
    abuf = randn(M, 3 * K, seed=1)
    cbuf = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    w = randn(N, K, std=K ** -0.5, seed=2)
    L.gemm(abuf[:, K:2 * K], w, cbuf[:, N:], None, 0)
    torch.cuda.synchronize()
    ref = abuf[:, K:2 * K].float() @ w.float().t()
    assert rel_l2(cbuf[:, N:], rb(ref)) < TOL
    assert cbuf[:, :N].abs().max().item() == 0


def conv_ref(x_cl, w, bias, kt, kh, kw, stride, pad, cin, cout):
    """x_cl [Tin,H,W,Cin] (already temporally padded) -> [Tout,Ho,Wo,Cout] fp32 via F.conv3d."""
    x = x_cl.float().permute(3, 0, 1, 2)[None]                       # [1,C,T,H,W]
    wt = w.float().reshape(cout, kt, kh, kw, cin).permute(0, 4, 1, 2, 3)
    if stride == 2:
        x = F.pad(x, (0, 1, 0, 1))
        y = F.conv3d(x, wt, bias.float(), stride=(1, 2, 2))
    else:
        y = F.conv3d(x, wt, bias.float(), padding=(0, pad, pad))
    return y[0].permute(1, 2, 3, 0)


@pytest.mark.parametrize("cin,cout,T,H,W,kt,stride", [
    (64, 128, 3, 20, 24, 3, 1), (128, 128, 2, 16, 160, 3, 1), (128, 256, 2, 33, 40, 3, 1),
    (64, 64, 1, 8, 8, 3, 1), (128, 128, 2, 32, 48, 1, 2), (256, 256, 1, 24, 20, 1, 2),
    (128, 128, 3, 16, 16, 1, 1), (512, 512, 2, 12, 20, 3, 1), (64, 16, 2, 16, 24, 3, 1)])
def test_conv(L, cin, cout, T, H, W, kt, stride):
    kh = kw = 3
    pad = 1 if stride == 1 else 0
    Tin = T + kt - 1
    x = randn(Tin, H, W, cin, seed=1)
    K = kt * kh * kw * cin
    w = randn(cout, K, std=K ** -0.5, seed=2)
    bias = randn(cout, std=0.1, seed=3)
    Ho, Wo = (H, W) if stride == 1 else (H // 2, W // 2)
    y = torch.full((T, Ho, Wo, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.conv_cl(x, w, bias, y, T, kt, kh, kw, stride, pad, Ho, Wo, cout)
    torch.cuda.synchronize()
    ref = conv_ref(x, w, bias, kt, kh, kw, stride, pad, cin, cout)
    assert torch.isfinite(y.float()).all()
    e = rel_l2(y, rb(ref))
    print(f"conv cin{cin} cout{cout} T{T} {H}x{W} kt{kt} s{stride}: rel_l2={e:.3e}")
    assert e < TOL
    # residual-add epilogue
    aux = randn(T, Ho, Wo, cout, seed=7)
    y2 = torch.empty_like(y)
    L.conv_cl(x, w, bias, y2, T, kt, kh, kw, stride, pad, Ho, Wo, cout, epilogue=L.EPI_ADD, aux=aux)
    torch.cuda.synchronize()
    assert rel_l2(y2, rb(rb(ref) + aux.float())) < TOL


def test_conv_planar_out(L):
    """conv_out of the decoder: 128 -> 3 channels (weights padded to 16 rows), planar NCDHW output."""
    cin, T, H, W = 128, 2, 16, 40
    x = randn(T + 2, H, W, cin, seed=1)
    K = 27 * cin
    w = torch.zeros(16, K, device="cuda", dtype=torch.bfloat16)
    w[:3] = randn(3, K, std=K ** -0.5, seed=2)
    bias = torch.zeros(16, device="cuda", dtype=torch.bfloat16)
    bias[:3] = randn(3, std=0.1, seed=3)
    y = torch.full((3, T, H, W), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.conv_cl(x, w, bias, y, T, 3, 3, 3, 1, 1, H, W, 3, out_mode=1)
    torch.cuda.synchronize()
    ref = conv_ref(x, w[:3], bias[:3], 3, 3, 3, 1, 1, cin, 3).permute(3, 0, 1, 2)
    assert rel_l2(y, rb(ref)) < TOL
    # fused post-processing epilogues (ref inference_script.py:501 and the savers' uint8 quantisation :124/:143/:168):
    # bit-exact functions of the raw bf16 output of the same kernel
    y2 = torch.empty_like(y)
    L.conv_cl(x, w, bias, y2, T, 3, 3, 3, 1, 1, H, W, 3, out_mode=L.OUT_PLANAR_POST)
    y3 = torch.empty((3, T, H, W), device="cuda", dtype=torch.uint8)
    L.conv_cl(x, w, bias, y3, T, 3, 3, 3, 1, 1, H, W, 3, out_mode=L.OUT_PLANAR_U8)
    torch.cuda.synchronize()
    unit = (y * 0.5 + 0.5).clamp(0.0, 1.0)                   # torch bf16 ops, as the reference applies them
    assert torch.equal(y2, unit)
    assert torch.equal(y3, (unit.float() * 255).clamp(0, 255).to(torch.uint8))
    assert 0 < y3.float().mean().item() < 255


@pytest.mark.parametrize("T,H,W,prev", [(2, 12, 256, True), (1, 9, 368, False), (3, 5, 130, True)])
def test_conv_narrow_output_kernel(L, T, H, W, prev):
    """conv_out (128 -> 3, weights padded to 16 rows) on rows >= 128 wide: the resident-weight / halo-box kernel
    (conv3.cu) against fp32 conv3d and against the generic kernel on the same problem, every planar output mode, with
    the causal cache frames through the second tensor map and with frame-0 replication."""
    cin = 128
    x = randn(T, H, W, cin, seed=1)
    pv = randn(2, H, W, cin, seed=2) if prev else None
    K = 27 * cin
    w = torch.zeros(16, K, device="cuda", dtype=torch.bfloat16)
    w[:3] = randn(3, K, std=K ** -0.5, seed=3)
    bias = torch.zeros(16, device="cuda", dtype=torch.bfloat16)
    bias[:3] = randn(3, std=0.1, seed=4)
    outs = {}
    for variant in (1, 0):
        L.set_option("conv2cta", variant)
        try:
            y = torch.full((3, T, H, W), float("nan"), device="cuda", dtype=torch.bfloat16)
            L.conv3d_causal(x, pv, w, bias, y, 3, out_mode=L.OUT_PLANAR)
            y2 = torch.empty_like(y)
            L.conv3d_causal(x, pv, w, bias, y2, 3, out_mode=L.OUT_PLANAR_POST)
            y3 = torch.empty((3, T, H, W), device="cuda", dtype=torch.uint8)
            L.conv3d_causal(x, pv, w, bias, y3, 3, out_mode=L.OUT_PLANAR_U8)
            torch.cuda.synchronize()
            outs[variant] = (y, y2, y3)
        finally:
            L.set_option("conv2cta", 1)
    padded = torch.cat([pv, x], 0) if prev else torch.cat([x[:1], x[:1], x], 0)
    ref = conv_ref(padded, w[:3], bias[:3], 3, 3, 3, 1, 1, cin, 3).permute(3, 0, 1, 2)
    y, y2, y3 = outs[1]
    assert torch.isfinite(y.float()).all()
    e, e2 = rel_l2(y, rb(ref)), rel_l2(y, outs[0][0])
    print(f"conv narrow T{T} {H}x{W} prev={prev}: vs fp32 {e:.3e}, vs generic kernel {e2:.3e}")
    assert e < TOL and e2 < TOL
    unit = (y * 0.5 + 0.5).clamp(0.0, 1.0)
    assert torch.equal(y2, unit)
    assert torch.equal(y3, (unit.float() * 255).clamp(0, 255).to(torch.uint8))


@pytest.mark.parametrize("rows,heads", [(128, 1), (200, 2), (482, 3), (1000, 2), (2304, 4), (19426, 2)])
def test_attention(L, rows, heads):
    qkv = randn(rows, 3 * heads * 64, seed=rows)
    out = torch.full((rows, heads * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.attention(qkv, out, heads, 0.125)
    torch.cuda.synchronize()
    q, k, v = [t.float().reshape(rows, heads, 64).transpose(0, 1) for t in qkv.chunk(3, dim=1)]
    p = torch.softmax(q @ k.transpose(1, 2) * 0.125, dim=-1)
    ref = (p @ v).transpose(0, 1).reshape(rows, heads * 64)
    assert torch.isfinite(out.float()).all()
    e = rel_l2(out, ref)
    print(f"attention rows{rows} heads{heads}: rel_l2={e:.3e}")
    assert e < 5e-3    # P is rounded to bf16 before the PV matmul (as flash SDPA does)


def _attn_ref(qkv, rows, heads, scale):
    q, k, v = [t.float().reshape(rows, heads, 64).transpose(0, 1) for t in qkv.chunk(3, dim=1)]
    p = torch.softmax(q @ k.transpose(1, 2) * scale, dim=-1)
    return (p @ v).transpose(0, 1).reshape(rows, heads * 64)


@pytest.mark.parametrize("variant", [-1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11])
@pytest.mark.parametrize("rows,heads", [(128, 1), (300, 2), (482, 3), (1000, 2), (4000, 2), (64, 1), (65, 1)])
def test_attention_variants(L, variant, rows, heads):
    """Every attention kernel variant (v2: one query tile per CTA; v3: two query tiles per CTA sharing K/V, with 0..5/8
    of the exponentials on the FMA-pipe polynomial; v4 = variants 7..11: S row held in registers (setmaxnreg) so that
    S(j+1) is issued while softmax(j) runs, 0..4/8 emulated; -1: automatic choice) against fp32 softmax attention, ragged
    sizes included."""
    L.set_option("attn_variant", variant)
    try:
        qkv = randn(rows, 3 * heads * 64, seed=rows + 1)
        out = torch.full((rows, heads * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        L.attention(qkv, out, heads, 0.125)
        torch.cuda.synchronize()
        assert torch.isfinite(out.float()).all()
        e = rel_l2(out, _attn_ref(qkv, rows, heads, 0.125))
        print(f"attention variant{variant} rows{rows} heads{heads}: rel_l2={e:.3e}")
        assert e < 5e-3
    finally:
        L.set_option("attn_variant", L.DEFAULT_ATTN_VARIANT)


@pytest.mark.parametrize("variant", [0, 2, 3, 5, 7, 8, 10])
def test_attention_wide_logit_range(L, variant):
    """Scores spanning tens of nats, a dominant key that appears only in the LAST key tile and one far-below-everything
    key: the steady-state softmax (no row-max pass, stale reference max) must detect the overflow through the row sum
    and fall back to the exact rescaling path; far-negative scores must not break the FMA-pipe exponential."""
    rows, heads = 1500, 2
    L.set_option("attn_variant", variant)
    try:
        g = torch.Generator(device="cuda").manual_seed(5)
        q = torch.randn(rows, heads * 64, generator=g, device="cuda") * 3
        k = torch.randn(rows, heads * 64, generator=g, device="cuda") * 3
        v = torch.randn(rows, heads * 64, generator=g, device="cuda")
        k[-3] = q[7] * 4                      # a late key that dominates (at least) query 7
        k[5] = -q[9] * 40                     # ~ -1e4 nats for query 9: far below the clamp of the polynomial path
        qkv = bf(torch.cat([q, k, v], dim=1))
        out = torch.full((rows, heads * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        L.attention(qkv, out, heads, 0.125)
        torch.cuda.synchronize()
        assert torch.isfinite(out.float()).all()
        ref = _attn_ref(qkv, rows, heads, 0.125)
        e = rel_l2(out, ref)
        print(f"attention wide-range variant{variant}: rel_l2={e:.3e}")
        assert e < 1e-2
        assert rel_l2(out[7], ref[7]) < 1e-2
    finally:
        L.set_option("attn_variant", L.DEFAULT_ATTN_VARIANT)


def test_layernorm_mod(L):
    rows, D, split = 333, 768, 100
    x = randn(rows, D, seed=1)
    w, b = randn(D, seed=2), randn(D, std=0.1, seed=3)
    sc = [randn(D, std=0.3, seed=4 + i) for i in range(4)]
    out = torch.empty_like(x)
    L.layernorm_mod(x, out, w, b, 1e-5, sc[0], sc[1], sc[2], sc[3], split)
    torch.cuda.synchronize()
    n = rb(F.layer_norm(x.float(), (D,), w.float(), b.float(), 1e-5))
    scale = torch.cat([sc[0].float().expand(split, D), sc[2].float().expand(rows - split, D)])
    shift = torch.cat([sc[1].float().expand(split, D), sc[3].float().expand(rows - split, D)])
    ref = rb(n * rb(1 + scale)) + shift
    assert rel_l2(out, rb(ref)) < 1e-3
    L.layernorm_mod(x, out, w, b, 1e-5)
    torch.cuda.synchronize()
    assert rel_l2(out, n) < 1e-3


def test_qk_norm_rope(L):
    from dove_b200.embeddings import get_3d_rotary_pos_embed
    heads, text = 3, 10
    T, h, w = 2, 4, 6
    rows = text + T * h * w
    qkv = randn(rows, 3 * heads * 64, seed=1)
    ws = [randn(64, seed=2), randn(64, std=0.1, seed=3), randn(64, seed=4), randn(64, std=0.1, seed=5)]
    cos, sin = get_3d_rotary_pos_embed(64, None, (h, w), T, grid_type="slice", max_size=(h, w), device="cuda")
    ref = qkv.clone().float()
    for i in range(2):
        t = ref[:, i * heads * 64:(i + 1) * heads * 64].reshape(rows, heads, 64)
        t = rb(F.layer_norm(t, (64,), ws[2 * i].float(), ws[2 * i + 1].float(), 1e-6))
        tv = t[text:]
        xr, xi = tv.reshape(rows - text, heads, 32, 2).unbind(-1)
        rot = torch.stack([-xi, xr], dim=-1).flatten(2)
        t[text:] = tv * cos[:, None] + rot * sin[:, None]
        ref[:, i * heads * 64:(i + 1) * heads * 64] = t.reshape(rows, heads * 64)
    L.qk_norm_rope(qkv, heads, *ws, 1e-6, cos.contiguous(), sin.contiguous(), text)
    torch.cuda.synchronize()
    assert rel_l2(qkv, rb(ref)) < 1e-3
    assert torch.equal(qkv[:, 2 * heads * 64:].float(), ref[:, 2 * heads * 64:])   # v untouched


@pytest.mark.parametrize("rows,heads,text", [(300, 4, 10), (2600, 4, 230), (1000, 8, 1000)])
def test_gemm_qkv_norm_rope_fused(L, rows, heads, text):
    """QKV projection with the per-head q/k LayerNorm + RoPE in the GEMM epilogue (one kernel) against the fp32
    arithmetic of Linear -> LayerNorm(64) -> apply_rotary_emb, and against the two-kernel path it replaces."""
    from dove_b200.embeddings import get_3d_rotary_pos_embed
    D = heads * 64
    nv = rows - text
    cos = sin = None
    if nv > 0:
        T, h = 2, 5
        w_ = nv // (T * h)
        assert T * h * w_ == nv
        cos, sin = get_3d_rotary_pos_embed(64, None, (h, w_), T, grid_type="slice", max_size=(h, w_), device="cuda")
        cos, sin = cos.contiguous(), sin.contiguous()
    a = randn(rows, D, seed=1)
    w = randn(3 * D, D, std=D ** -0.5, seed=2)
    bias = randn(3 * D, std=0.1, seed=3)
    ws = [randn(64, seed=4), randn(64, std=0.1, seed=5), randn(64, seed=6), randn(64, std=0.1, seed=7)]
    fused = torch.full((rows, 3 * D), float("nan"), device="cuda", dtype=torch.bfloat16)
    ct, st = L.rope_tables_transposed(cos, sin) if nv > 0 else (None, None)
    L.gemm_qkv_norm_rope(a, w, fused, bias, heads, *ws, 1e-6, ct, st, text)
    two = torch.empty_like(fused)
    L.gemm(a, w, two, bias)
    L.qk_norm_rope(two, heads, *ws, 1e-6, cos, sin, text)
    torch.cuda.synchronize()
    ref = rb(a.float() @ w.float().t() + bias.float())
    for i in range(2):
        t = ref[:, i * D:(i + 1) * D].reshape(rows, heads, 64)
        t = rb(F.layer_norm(t, (64,), ws[2 * i].float(), ws[2 * i + 1].float(), 1e-6))
        if nv > 0:
            tv = t[text:]
            xr, xi = tv.reshape(nv, heads, 32, 2).unbind(-1)
            rot = torch.stack([-xi, xr], dim=-1).flatten(2)
            t[text:] = tv * cos[:, None] + rot * sin[:, None]
        ref[:, i * D:(i + 1) * D] = t.reshape(rows, D)
    assert torch.isfinite(fused.float()).all()
    e, e2 = rel_l2(fused, rb(ref)), rel_l2(fused, two)
    print(f"gemm_qkv_norm_rope rows{rows} heads{heads}: vs fp32 {e:.3e}, vs two-kernel path {e2:.3e}")
    assert e < TOL and e2 < TOL
    assert torch.equal(fused[:, 2 * D:], two[:, 2 * D:])          # v columns: plain Linear, identical


def test_gemv(L):
    N, K = 1000, 512
    x, w, b = randn(K, seed=1), randn(N, K, std=K ** -0.5, seed=2), randn(N, std=0.1, seed=3)
    y = torch.empty(N, device="cuda", dtype=torch.bfloat16)
    L.gemv(x, w, b, y, silu_in=True)
    torch.cuda.synchronize()
    ref = rb(F.silu(x.float())) @ w.float().t() + b.float()
    assert rel_l2(y, rb(ref)) < 1e-3


def test_patchify_roundtrip(L):
    Fr, C, h, w = 4, 16, 6, 10
    lat = randn(Fr, C, h, w, seed=1)
    tok = torch.empty(Fr // 2 * h // 2 * w // 2, C * 8, device="cuda", dtype=torch.bfloat16)
    L.patchify(lat, tok)
    ref = lat.permute(0, 2, 3, 1).reshape(Fr // 2, 2, h // 2, 2, w // 2, 2, C).permute(0, 2, 4, 6, 1, 3, 5)
    torch.cuda.synchronize()
    assert torch.equal(tok, ref.reshape(tok.shape))      # bit-exact index bookkeeping
    pred = torch.empty_like(lat)
    x0 = torch.empty_like(lat)
    L.unpatchify_velocity(tok, lat, x0, pred, Fr, C, h, w, 0.625, 0.78125)
    torch.cuda.synchronize()
    assert torch.equal(pred, lat)
    ref0 = rb(0.625 * lat.float()) - rb(0.78125 * lat.float())
    assert torch.equal(x0.float(), rb(ref0))


@pytest.mark.parametrize("C", [128, 256, 512])
@pytest.mark.parametrize("spatial", [False, True])
def test_groupnorm(L, C, spatial):
    T, H, W = 3, 8, 12
    x = bf(randn(T, H, W, C, seed=1).float() * 2 + 0.5)
    gamma, beta = randn(C, seed=2), randn(C, std=0.2, seed=3)
    partial = torch.empty(L.gn_partial_floats(T * H * W), device="cuda", dtype=torch.float32)
    stats = torch.empty(64, device="cuda", dtype=torch.float32)
    L.gn_stats(x, C, 32, 1e-6, partial, stats)
    out = torch.empty_like(x)
    zy = zb = None
    if spatial:
        zy, zb = randn(2, H // 4, W // 4, C, seed=4), randn(2, H // 4, W // 4, C, seed=5)
    L.gn_apply(x, out, T, H, W, C, 32, stats, gamma, beta, True, zy, zb)
    torch.cuda.synchronize()
    xn = x.float().permute(3, 0, 1, 2)[None]
    ref = rb(F.group_norm(xn, 32, gamma.float(), beta.float(), 1e-6))
    if spatial:
        def up(z):   # SpatialNorm3D nearest interpolation with the odd-T first-frame split
            z = z.float().permute(3, 0, 1, 2)[None]
            first = F.interpolate(z[:, :, :1], size=(1, H, W))
            rest = F.interpolate(z[:, :, 1:], size=(T - 1, H, W))
            return torch.cat([first, rest], dim=2)
        ref = rb(rb(ref * up(zy)) + up(zb))
    ref = F.silu(ref)[0].permute(1, 2, 3, 0)
    e = rel_l2(out, rb(ref))
    print(f"groupnorm C{C} spatial{spatial}: {e:.3e}")
    assert e < 1e-3


def test_time_pool_and_upsample(L):
    H, W, C = 6, 10, 128
    for T in (1, 2, 5, 8, 9):
        x = randn(T, H, W, C, seed=T)
        Tout = 1 + (T - 1) // 2 if T % 2 else T // 2
        y = torch.empty(Tout, H, W, C, device="cuda", dtype=torch.bfloat16)
        L.time_pool(x, y, T, H * W * C)
        xf = x.float()
        if T % 2:
            parts = [xf[:1]] + ([(xf[1::2] + xf[2::2]) * 0.5] if T > 1 else [])
            ref = torch.cat(parts)
        else:
            ref = (xf[0::2] + xf[1::2]) * 0.5
        torch.cuda.synchronize()
        assert torch.equal(y.float(), rb(ref)), T
        for tx2 in (False, True):
            xn = xf.permute(3, 0, 1, 2)[None]
            if tx2 and T > 1 and T % 2:
                first = F.interpolate(xn[:, :, 0], scale_factor=2.0)
                rest = F.interpolate(xn[:, :, 1:], scale_factor=2.0)
                r = torch.cat([first[:, :, None], rest], dim=2)
            elif tx2 and T > 1:
                r = F.interpolate(xn, scale_factor=2.0)
            else:
                r = F.interpolate(xn, scale_factor=(1.0, 2.0, 2.0))
            r = r[0].permute(1, 2, 3, 0)
            u = torch.empty(r.shape, device="cuda", dtype=torch.bfloat16)
            L.upsample_nearest(x, u, T, H, W, C, tx2)
            torch.cuda.synchronize()
            assert torch.equal(u.float(), r), (T, tx2)


def test_layout_and_sampling(L):
    T, H, W = 3, 8, 12
    pix = torch.rand(3, T, H, W, device="cuda") * 2 - 1
    y = torch.empty(T, H, W, 64, device="cuda", dtype=torch.bfloat16)
    L.pixels_to_cl(pix, y, T, H, W, 64)
    torch.cuda.synchronize()
    assert torch.equal(y[..., :3].float(), rb(pix.permute(1, 2, 3, 0)))
    assert y[..., 3:].abs().max().item() == 0
    lat = randn(16, T, H, W, seed=1)
    cl = torch.empty(T, H, W, 64, device="cuda", dtype=torch.bfloat16)
    L.ncthw_to_cl(lat, cl, 16, T, H, W, 64, 1 / 0.7)
    torch.cuda.synchronize()
    assert torch.equal(cl[..., :16].float(), rb(lat.float().permute(1, 2, 3, 0) * (1 / 0.7)))
    back = torch.empty(16, T, H, W, device="cuda", dtype=torch.bfloat16)
    L.ncthw_to_cl(lat, cl, 16, T, H, W, 64, 1.0)
    L.cl_to_ncthw(cl, back, 16, T, H, W, 64)
    torch.cuda.synchronize()
    assert torch.equal(back, lat)
    mom = randn(T, H, W, 32, seed=2)
    noise = randn(16, T, H, W, seed=3)
    z = torch.empty(16, T, H, W, device="cuda", dtype=torch.bfloat16)
    L.gaussian_sample(mom, noise, z, T * H * W, 0.7)
    torch.cuda.synchronize()
    m = mom.float().permute(3, 0, 1, 2)
    std = rb(torch.exp(rb(0.5 * m[16:].clamp(-30, 20))))
    ref = rb(rb(m[:16] + rb(std * noise.float())) * 0.7)
    assert rel_l2(z, ref) < 1e-3
    o = torch.empty_like(noise)
    L.post_scale(noise, o)
    torch.cuda.synchronize()
    assert torch.equal(o.float(), rb(rb(rb(noise.float() * 0.5) + 0.5).clamp(0, 1)))


@pytest.mark.parametrize("variant", [1, 2, 0])
@pytest.mark.parametrize("cin,cout,T,H,W,kt", [(128, 128, 2, 6, 256, 3), (128, 256, 2, 5, 384, 3), (256, 256, 2, 4, 320, 1),
                                               (256, 128, 1, 8, 256, 3), (64, 128, 3, 3, 640, 3),
                                               # Cout = 128 on >= 4096-voxel frames: swapped-operand kernel (N = 256 voxels)
                                               (128, 128, 2, 16, 256, 3), (256, 128, 1, 24, 200, 3), (64, 128, 2, 64, 64, 1),
                                               (128, 128, 3, 33, 136, 3),
                                               # ... and, from 256 voxels of width, its halo-row variant (conv4.cu): exact /
                                               # ragged widths (narrow last MMA), kt = 1 and 3, 64..256 in-channels
                                               (128, 128, 2, 9, 512, 3), (64, 128, 1, 17, 496, 3), (256, 128, 1, 16, 256, 1),
                                               (128, 128, 1, 5, 1280, 3), (128, 128, 2, 8, 752, 1), (128, 128, 2, 12, 368, 3),
                                               (128, 128, 1, 16, 264, 3), (64, 128, 2, 11, 300, 1)])
def test_conv_cta_pair(L, cin, cout, T, H, W, kt, variant):
    """Wide stride-1 3x3(x3) convs on the specialised kernels: CTA pair (cta_group::2 + W-tap reuse through shifted smem
    descriptors, Cout = 256), swapped operands (Cout = 128) and its halo-row variant.  variant 1 = shipping dispatch,
    2 = without the halo-row kernel (the generic swapped-operand kernel on the same problem), 0 = 1-CTA kernel."""
    L.set_option("conv2cta", variant)
    try:
        x = randn(T + kt - 1, H, W, cin, seed=1)
        K = kt * 9 * cin
        w = randn(cout, K, std=K ** -0.5, seed=2)
        bias = randn(cout, std=0.1, seed=3)
        aux = randn(T, H, W, cout, seed=7)
        y = torch.full((T, H, W, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
        L.conv_cl(x, w, bias, y, T, kt, 3, 3, 1, 1, H, W, cout, epilogue=L.EPI_ADD, aux=aux)
        torch.cuda.synchronize()
        ref = conv_ref(x, w, bias, kt, 3, 3, 1, 1, cin, cout)
        e = rel_l2(y, rb(rb(ref) + aux.float()))
        print(f"conv2cta variant{variant} cin{cin} cout{cout} T{T} {H}x{W} kt{kt}: rel_l2={e:.3e}")
        assert torch.isfinite(y.float()).all()
        assert e < TOL
    finally:
        L.set_option("conv2cta", 1)


def test_preprocess_matches_reference_recipe(L):
    """ref inference_script.py:192-235 + :670-679: pad F to 8k+1 / HW to x16, bilinear x4 on 0..255 floats, x/255*2-1."""
    from dove_b200.runner import preprocess_frames, remove_padding_and_extra_frames
    g = torch.Generator().manual_seed(0)
    frames = torch.randint(0, 256, (10, 45, 70, 3), generator=g, dtype=torch.uint8)
    out, pf, ph, pw = preprocess_frames(frames, 4, "cuda")
    assert (pf, ph, pw) == (7, 3, 10) and out.shape == (1, 3, 17, 192, 320)
    x = torch.cat([frames, frames[-1:].repeat(7, 1, 1, 1)], 0)
    x = F.pad(x, (0, 0, 0, 10, 0, 3)).float().permute(0, 3, 1, 2)
    ref = (F.interpolate(x, scale_factor=4, mode="bilinear") / 255.0 * 2.0 - 1.0).permute(1, 0, 2, 3)[None]
    torch.cuda.synchronize()
    assert (out.cpu() - ref).abs().max().item() < 2e-6
    assert remove_padding_and_extra_frames(out, pf, ph, pw).shape == (1, 3, 10, 180, 280)


@pytest.mark.parametrize("cin,cout,T,H,W", [(64, 64, 3, 12, 20), (128, 128, 2, 16, 256), (128, 256, 3, 6, 384),
                                            (256, 128, 1, 24, 200), (512, 512, 2, 12, 20), (128, 128, 3, 9, 512)])
def test_conv3d_causal_zero_copy_cache(L, cin, cout, T, H, W):
    """dove_conv3d_causal_bf16 (two preceding frames through a second tensor map / frame-0 replication) must be
    bit-identical to the same kernel run on the materialised padded input, for the generic, swapped-operand (halo-row
    variant at 128 -> 128, 16 x 256 and 9 x 512) and CTA-pair kernels."""
    x = randn(T, H, W, cin, seed=1)
    prev = randn(2, H, W, cin, seed=2)
    K = 27 * cin
    w = randn(cout, K, std=K ** -0.5, seed=3)
    bias = randn(cout, std=0.1, seed=4)
    aux = randn(T, H, W, cout, seed=5)
    for pv in (prev, None):
        padded = torch.cat([pv, x], 0) if pv is not None else torch.cat([x[:1], x[:1], x], 0)
        y_ref = torch.empty(T, H, W, cout, device="cuda", dtype=torch.bfloat16)
        L.conv_cl(padded.contiguous(), w, bias, y_ref, T, 3, 3, 3, 1, 1, H, W, cout, epilogue=L.EPI_ADD, aux=aux)
        y = torch.full_like(y_ref, float("nan"))
        L.conv3d_causal(x, pv, w, bias, y, cout, epilogue=L.EPI_ADD, aux=aux)
        torch.cuda.synchronize()
        assert torch.equal(y, y_ref), (cin, cout, pv is None)
    ref = conv_ref(torch.cat([prev, x], 0), w, bias, 3, 3, 3, 1, 1, cin, cout)
    assert rel_l2(y_ref if False else y, rb(rb(conv_ref(torch.cat([x[:1], x[:1], x], 0), w, bias, 3, 3, 3, 1, 1, cin, cout))
                                            + aux.float())) < TOL
    del ref


@pytest.mark.parametrize("cin,cout,T,H,W", [(128, 128, 2, 16, 256), (256, 128, 1, 24, 200), (128, 256, 2, 6, 384),
                                            (256, 512, 2, 5, 320), (64, 128, 2, 12, 20), (256, 256, 2, 26, 23),
                                            (512, 512, 1, 13, 12), (128, 128, 2, 11, 496)])
def test_conv_fused_groupnorm_stats(L, cin, cout, T, H, W):
    """GroupNorm statistics of the conv OUTPUT accumulated in the conv epilogue (swapped-operand and CTA-pair kernels)
    equal the statistics of a separate pass over the stored tensor (generic, swapped-operand and CTA-pair kernels)."""
    x = randn(T, H, W, cin, seed=1)
    K = 27 * cin
    w = randn(cout, K, std=K ** -0.5, seed=3)
    bias = randn(cout, std=0.3, seed=4)
    aux = randn(T, H, W, cout, seed=5)
    y = torch.empty(T, H, W, cout, device="cuda", dtype=torch.bfloat16)
    partial = torch.full((L.gn_partial_floats(0),), float("nan"), device="cuda", dtype=torch.float32)
    _, done = L.conv3d_causal(x, None, w, bias, y, cout, epilogue=L.EPI_ADD, aux=aux, gn_partial=partial)
    fused = torch.empty(64, device="cuda", dtype=torch.float32)
    ref = torch.empty(64, device="cuda", dtype=torch.float32)
    partial2 = torch.empty_like(partial)
    L.gn_stats(y, cout, 32, 1e-6, partial2, ref)
    torch.cuda.synchronize()
    assert done                                      # every conv kernel with Cout % 128 == 0 fuses the statistics
    if done:
        L.gn_finalize(partial, T * H * W, cout, 32, 1e-6, fused)
        torch.cuda.synchronize()
        assert torch.isfinite(fused).all()
        assert torch.allclose(fused, ref, rtol=2e-4, atol=2e-5), (fused - ref).abs().max()


def test_h2d_box_strided_dma(L):
    """dove_h2d_box_async: a chunk x tile view of a pinned host clip lands on the device bit for bit (fp32 and uint8,
    first / interior / last tile, single-frame chunk) — the per-unit H2D of runner.super_resolve."""
    from dove_b200.runner import _unit_to_device
    torch.manual_seed(3)
    clip = torch.rand(1, 3, 9, 40, 56).pin_memory()
    for (t0, t1, h0, h1, w0, w1) in [(0, 9, 0, 16, 0, 24), (2, 7, 8, 40, 16, 56), (4, 5, 24, 40, 32, 56), (0, 9, 0, 40, 0, 56)]:
        unit = clip[:, :, t0:t1, h0:h1, w0:w1]
        got = _unit_to_device(unit, "cuda")
        torch.cuda.synchronize()
        assert got.is_contiguous() and torch.equal(got.cpu(), unit.contiguous())
    u8 = (clip * 255).to(torch.uint8).pin_memory()
    dst = torch.empty(5, 16, 24, dtype=torch.uint8, device="cuda")
    L.h2d_box(u8[0, 1, 2:7, 8:24, 16:40], dst)
    torch.cuda.synchronize()
    assert torch.equal(dst.cpu(), u8[0, 1, 2:7, 8:24, 16:40])
    pageable = torch.rand(1, 3, 4, 32, 32)                  # pageable source: still correct (the driver stages it)
    got = _unit_to_device(pageable[:, :, 1:3, 8:24, 0:16], "cuda")
    torch.cuda.synchronize()
    assert torch.equal(got.cpu(), pageable[:, :, 1:3, 8:24, 0:16])
