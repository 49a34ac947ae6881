"""Analytic work model of the hot path: multiply-accumulates of every conv / linear, per stage.

Used by (a) tests/test_oracle_pins.py to pin the architecture restatement against the reference's published
504.81 T MACs at 33x720x1280 (assets/Quantitative-2.png via README.md:278-282 — the figure counts module MACs, i.e.
it excludes the SDPA QK^T / PV matmuls), and (b) bench.py to turn measured times into fractions of the FLOP roofline
(FLOPs = 2 * MACs, biases / norms / elementwise excluded, untiled VAE).
"""
from __future__ import annotations

from .weights import DIT_CONFIG, VAE_CONFIG


def _pool_t(t):
    return 1 + (t - 1) // 2 if t % 2 else t // 2


def _up_t(t):
    return (1 + 2 * (t - 1) if t % 2 else 2 * t) if t > 1 else t


def frame_batches(n, bs):
    nb, rem = max(n // bs, 1), n % bs
    return [(bs * i + (0 if i == 0 else rem), min(bs * (i + 1) + rem, n)) for i in range(nb)]


def vae_encoder_macs(F, H, W, cfg=None):
    c = dict(VAE_CONFIG)
    c.update(cfg or {})
    boc, lpb, zc = list(c["block_out_channels"]), c["layers_per_block"], c["latent_channels"]
    total = 0
    for s, e in frame_batches(F, 8):
        t, h, w = e - s, H, W
        total += 27 * c["in_channels"] * boc[0] * t * h * w
        ch = boc[0]
        for i, oc in enumerate(boc):
            for j in range(lpb):
                cin = ch if j == 0 else oc
                total += 27 * cin * oc * t * h * w + 27 * oc * oc * t * h * w
                if cin != oc:
                    total += cin * oc * t * h * w
            ch = oc
            if i != len(boc) - 1:
                if i < 2:
                    t = _pool_t(t)
                h, w = h // 2, w // 2
                total += 9 * oc * oc * t * h * w
        total += 2 * 2 * 27 * ch * ch * t * h * w                      # mid block
        total += 27 * ch * 2 * zc * t * h * w                          # conv_out
    return total


def vae_decoder_macs(Tz, h, w, cfg=None):
    c = dict(VAE_CONFIG)
    c.update(cfg or {})
    rev, lpb, zc = list(c["block_out_channels"])[::-1], c["layers_per_block"], c["latent_channels"]
    total = 0
    for s, e in frame_batches(Tz, 2):
        t, hh, ww = e - s, h, w

        def resnet(cin, cout):
            m = 27 * cin * cout * t * hh * ww + 27 * cout * cout * t * hh * ww
            m += 2 * zc * cin * t * hh * ww + 2 * zc * cout * t * hh * ww     # SpatialNorm3D conv_y / conv_b (1x1x1)
            if cin != cout:
                m += cin * cout * t * hh * ww
            return m

        total += 27 * zc * rev[0] * t * hh * ww
        total += 2 * resnet(rev[0], rev[0])
        ch = rev[0]
        for i, oc in enumerate(rev):
            for j in range(lpb + 1):
                total += resnet(ch if j == 0 else oc, oc)
            ch = oc
            if i != len(rev) - 1:
                if i < 2:
                    t = _up_t(t)
                hh, ww = hh * 2, ww * 2
                total += 9 * oc * oc * t * hh * ww
        total += 2 * zc * ch * t * hh * ww                              # norm_out SpatialNorm3D
        total += 27 * ch * c["out_channels"] * t * hh * ww              # conv_out
    return total


def latent_frames(F):
    return sum(_pool_t(_pool_t(e - s)) for s, e in frame_batches(F, 8))


def dit_macs(F, H, W, cfg=None, text_len=226, include_sdpa=True):
    """One forward of the DiT on the latent of an F x H x W pixel clip -> (linear MACs, SDPA-core MACs, tokens)."""
    c = dict(DIT_CONFIG)
    c.update(cfg or {})
    D = c["num_attention_heads"] * c["attention_head_dim"]
    te, pt, p = c["time_embed_dim"], c["patch_size_t"], c["patch_size"]
    tl = latent_frames(F)
    tl += tl % pt
    nv = (tl // pt) * (H // 8 // p) * (W // 8 // p)
    n = nv + text_len
    pfeat = c["in_channels"] * p * p * pt
    lin = nv * pfeat * D + text_len * c["text_embed_dim"] * D + D * te + te * te
    per_layer = n * (4 * D * D + 2 * c["ff_mult"] * D * D) + 2 * te * 6 * D
    lin += c["num_layers"] * per_layer + te * 2 * D + nv * D * pfeat
    sdpa = c["num_layers"] * 2 * n * n * D if include_sdpa else 0
    return lin, sdpa, n


def clip_macs(F, H, W, dit_cfg=None):
    """Per-stage MACs of one unit (untiled VAE)."""
    enc = vae_encoder_macs(F, H, W)
    lin, sdpa, n = dit_macs(F, H, W, dit_cfg)
    dec = vae_decoder_macs(latent_frames(F), H // 8, W // 8)
    return dict(encoder=enc, dit_linear=lin, dit_sdpa=sdpa, decoder=dec, tokens=n,
                module_macs=enc + lin + dec, total=enc + lin + sdpa + dec)
