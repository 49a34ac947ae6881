"""Pins of the CPU oracle (SURVEY.md section 8c "independent pins") and oracle self-consistency.  CPU only."""
import torch

from dove_b200.weights import dit_param_spec, init_state_dict, param_count, vae_param_spec
from oracle.dit import OracleCogVideoXTransformer3DModel, get_3d_rotary_pos_embed
from oracle.scheduler import OracleCogVideoXDPMScheduler
from oracle.vae import CausalConv3d, OracleAutoencoderKLCogVideoX


def test_parameter_count_pin():
    v, d = param_count(vae_param_spec()), param_count(dit_param_spec())
    assert v == 215_583_907 and d == 5_570_676_352      # incl. patch_embed.proj.bias (3072)
    published = 5787.19e6            # assets/Quantitative-2.png
    assert abs((v + d) - published) / published < 5e-4


def test_published_mac_count_pin():
    """assets/Quantitative-2.png: 504.81 T MACs for DOVE at 33x720x1280 (module MACs: the SDPA core is not counted).
    The analytic model of the restated architecture gives 504.60 T (0.04 %), SURVEY.md section 6 / 8c pin (2)."""
    from dove_b200.workmodel import clip_macs
    m = clip_macs(33, 720, 1280)
    assert m["tokens"] == 18226
    assert abs(m["module_macs"] / 1e12 - 504.81) / 504.81 < 1e-3
    assert abs(m["module_macs"] / 1e12 - 504.60) < 0.01
    # SURVEY 8d FLOP table (2 * MACs, SDPA core included): cfg-1, cfg-2, cfg-4 chunk
    for shape, tflop in (((8, 256, 256), 19.0), ((33, 768, 1280), 1271.1), ((25, 1088, 1920), 2309.6)):
        assert abs(2 * clip_macs(*shape)["total"] / 1e12 - tflop) < 0.06, shape


def test_prompt_embedding_fixture():
    """The shipped pre-computed T5 embedding of "" (ref :580-590): key, shape, dtype and statistics."""
    import hashlib
    from pathlib import Path
    from safetensors.torch import load_file
    f = Path(__file__).resolve().parent / "golden" / "empty_prompt_embedding.safetensors"
    sd = load_file(str(f))
    assert list(sd) == ["prompt_embedding"]
    t = sd["prompt_embedding"]
    assert t.shape == (226, 4096) and t.dtype == torch.bfloat16
    assert abs(t.float().mean().item() - 0.00175) < 1e-4 and abs(t.float().std().item() - 0.1485) < 1e-3
    assert abs(t.float().abs().max().item() - 6.47) < 0.01
    assert hashlib.sha256(f.read_bytes()).hexdigest() == "49738b5f634bc7c7ebad8e0ba01bf8c4eb5930b84c38d00f78dc0d5bc0a417cc"


def test_strict_checkpoint_keys():
    """A checkpoint tensor the build does not consume (or a missing one) must raise, never be dropped silently
    (the CogVideoX-1.5 patch_embed.proj.bias exists although config.patch_bias is false)."""
    names = {n for n, _, _ in dit_param_spec()}
    assert "patch_embed.proj.bias" in names and "patch_embed.proj.weight" in names
    from dove_b200.vae import _TrackedStateDict
    sd = _TrackedStateDict({"a": 1, "b": 2}, "vae")
    assert sd["a"] == 1 and "b" in sd
    try:
        sd.check_all_consumed()
        raise AssertionError("unconsumed key not reported")
    except KeyError as e:
        assert "unconsumed" in str(e)
    try:
        sd["c"]
        raise AssertionError("missing key not reported")
    except KeyError as e:
        assert "missing" in str(e)


def test_spec_matches_oracle_modules():
    with torch.device("meta"):
        vae = OracleAutoencoderKLCogVideoX()
        dit = OracleCogVideoXTransformer3DModel(num_layers=2)
    assert {k: tuple(t.shape) for k, t in vae.state_dict().items()} == {n: tuple(s) for n, s, _ in vae_param_spec()}
    assert {k: tuple(t.shape) for k, t in dit.state_dict().items()} == \
        {n: tuple(s) for n, s, _ in dit_param_spec(dict(num_layers=2))}


def test_scheduler_constants():
    s = OracleCogVideoXDPMScheduler()
    a = s.alphas_cumprod[399]
    assert abs(a.item() - 0.393544057504) < 1e-11
    assert abs(a.float().sqrt().item() - 0.6273309) < 1e-6 and abs((1 - a.float()).sqrt().item() - 0.7787528) < 1e-6
    ab = s.alphas_cumprod.to(torch.bfloat16)[399]
    assert ab.item() == 0.392578125 and (ab ** 0.5).item() == 0.625 and ((1 - ab) ** 0.5).item() == 0.78125
    assert s.alphas_cumprod[-1].item() == 0.0           # zero terminal SNR
    from dove_b200.scheduler import CogVideoXDPMScheduler
    p = CogVideoXDPMScheduler()
    assert torch.equal(p.alphas_cumprod, s.alphas_cumprod)
    assert p.coefficients(399, torch.bfloat16) == (0.625, 0.78125)


def test_vae_integers():
    v = OracleAutoencoderKLCogVideoX.__new__(OracleAutoencoderKLCogVideoX)
    torch.nn.Module.__init__(v)
    v.tile_sample_min_height, v.tile_sample_min_width = 240, 360
    v.tile_latent_min_height, v.tile_latent_min_width = 30, 45
    v.tile_overlap_factor_height, v.tile_overlap_factor_width = 1 / 6, 1 / 5
    enc, dec = v.tile_ints()
    assert (enc["stride_h"], enc["stride_w"], enc["blend_h"], enc["blend_w"], enc["limit_h"], enc["limit_w"]) == \
        (200, 288, 5, 9, 25, 36)
    assert (dec["stride_h"], dec["stride_w"], dec["blend_h"], dec["blend_w"], dec["limit_h"], dec["limit_w"]) == \
        (25, 36, 40, 72, 200, 288)
    fb = OracleAutoencoderKLCogVideoX.frame_batches
    assert fb(33, 8) == [(0, 9), (9, 17), (17, 25), (25, 33)] and fb(9, 2) == [(0, 3), (3, 5), (5, 7), (7, 9)]
    assert fb(8, 8) == [(0, 8)] and fb(54, 8)[0] == (0, 14)
    from dove_b200.vae import AutoencoderKLCogVideoX
    assert AutoencoderKLCogVideoX.frame_batches(33, 8) == fb(33, 8)
    assert [AutoencoderKLCogVideoX.latent_frames(f) for f in (8, 9, 33, 54, 129)] == [2, 3, 9, 14, 33]


def test_rope_index_order_and_product_table():
    cos, sin = get_3d_rotary_pos_embed(64, None, (3, 4), 2, grid_type="slice", max_size=(3, 4), device="cpu")
    assert cos.shape == (24, 64) and cos.dtype == torch.float32
    t, y, x = 1, 2, 3
    row = (t * 3 + y) * 4 + x
    f_t = 1.0 / (10000 ** (torch.arange(0, 16, 2).float() / 16))
    f_h = 1.0 / (10000 ** (torch.arange(0, 24, 2).float() / 24))
    exp = torch.cat([(t * f_t).cos().repeat_interleave(2), (y * f_h).cos().repeat_interleave(2),
                     (x * f_h).cos().repeat_interleave(2)])
    assert torch.allclose(cos[row], exp, atol=1e-6)
    from dove_b200.embeddings import get_3d_rotary_pos_embed as ours
    oc, os_ = ours(64, None, (3, 4), 2, grid_type="slice", max_size=(3, 4), device="cpu")
    assert torch.equal(oc, cos) and torch.equal(os_, sin)           # bit-exact bookkeeping


def test_causal_conv_cache_equivalence():
    """frame-batched conv with cache == single pass (convs only; GroupNorm extents differ by design)."""
    torch.manual_seed(0)
    c = CausalConv3d(4, 6, 3)
    x = torch.randn(1, 4, 9, 5, 7)
    full, _ = c(x)
    a, cache = c(x[:, :, :5])
    b, _ = c(x[:, :, 5:], cache)
    assert torch.allclose(torch.cat([a, b], 2), full, atol=1e-5)


def test_vae_shapes_small():
    torch.manual_seed(0)
    sd = init_state_dict(vae_param_spec(), 1, "cpu", torch.float32)
    vae = OracleAutoencoderKLCogVideoX()
    vae.load_state_dict(sd)
    with torch.no_grad():
        m = vae.encode(torch.rand(1, 3, 9, 16, 16) * 2 - 1).latent_dist
        assert m.parameters.shape == (1, 32, 3, 2, 2)
        out = vae.decode(m.mode()).sample
    assert out.shape == (1, 3, 9, 16, 16) and torch.isfinite(out).all()


def test_frame_bookkeeping_against_oracle_control_flow():
    """The product's frame-count bookkeeping (latent_frames, decoded frame counts, tile integers) against the oracle's
    actual encode/decode control flow, evaluated shape-only on the meta device."""
    from dove_b200.vae import AutoencoderKLCogVideoX as P
    with torch.device("meta"):
        vae = OracleAutoencoderKLCogVideoX()
        for F in (8, 9, 17, 24, 25, 33, 54):
            m = vae._encode(torch.empty(1, 3, F, 32, 48))
            assert m.shape == (1, 32, P.latent_frames(F), 4, 6), F
            d = vae._decode(torch.empty(1, 16, m.shape[2], 4, 6))
            # decoded frame count: first latent batch keeps frame 0 single (odd) -> what the product pre-computes
            total = 0
            for s, e in P.frame_batches(m.shape[2], 2):
                t = e - s
                for _ in range(2):
                    t = (1 + 2 * (t - 1) if t % 2 else 2 * t) if t > 1 else t
                total += t
            assert d.shape == (1, 3, total, 32, 48), (F, d.shape, total)
    p = P.__new__(P)
    from types import SimpleNamespace
    from dove_b200.weights import VAE_CONFIG
    p.config = SimpleNamespace(**VAE_CONFIG)
    enc, dec = p.tile_ints()
    assert (enc["stride_h"], enc["stride_w"], enc["blend_h"], enc["blend_w"], enc["limit_h"], enc["limit_w"]) == \
        (200, 288, 5, 9, 25, 36)
    assert (dec["stride_h"], dec["stride_w"], dec["blend_h"], dec["blend_w"], dec["limit_h"], dec["limit_w"]) == \
        (25, 36, 40, 72, 200, 288)


def test_oracle_blend_ramp():
    a = torch.ones(1, 1, 1, 6, 4)
    b = torch.zeros(1, 1, 1, 6, 4)
    out = OracleAutoencoderKLCogVideoX.blend_v(a, b.clone(), 4)
    assert torch.allclose(out[0, 0, 0, :, 0], torch.tensor([1.0, 0.75, 0.5, 0.25, 0.0, 0.0]))
    out = OracleAutoencoderKLCogVideoX.blend_h(a, b.clone(), 2)
    assert torch.allclose(out[0, 0, 0, 0], torch.tensor([1.0, 0.5, 0.0, 0.0]))
