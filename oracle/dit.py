"""Oracle: CogVideoX-1.5 DiT forward (restatement of diffusers `CogVideoXTransformer3DModel`).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows, by upstream module path,
`diffusers/models/transformers/cogvideox_transformer_3d.py` (CogVideoXBlock,
CogVideoXTransformer3DModel.forward), `models/attention_processor.py`
(Attention, CogVideoXAttnProcessor2_0), `models/normalization.py` (CogVideoXLayerNormZero,
AdaLayerNorm), `models/attention.py::FeedForward`, `models/activations.py::GELU`,
`models/embeddings.py` (CogVideoXPatchEmbed, Timesteps, TimestepEmbedding,
get_timestep_embedding, get_3d_rotary_pos_embed, get_1d_rotary_pos_embed, apply_rotary_emb).
Reference call sites: /root/reference/inference_script.py:364-392 (RoPE tables),
:483-489 (transformer call).  Parameter names match diffusers state-dict keys.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

DIT_CONFIG = dict(
    num_attention_heads=48, attention_head_dim=64, in_channels=16, out_channels=16,
    flip_sin_to_cos=True, freq_shift=0, time_embed_dim=512, text_embed_dim=4096, num_layers=42,
    patch_size=2, patch_size_t=2, patch_bias=False, max_text_seq_length=226,
    norm_eps=1e-5, use_rotary_positional_embeddings=True, use_learned_positional_embeddings=False,
    sample_height=300, sample_width=300, sample_frames=81, temporal_compression_ratio=4,
    ff_mult=4,
)


# ---- embeddings -----------------------------------------------------------------------------
def get_timestep_embedding(timesteps, dim, flip_sin_to_cos=True, freq_shift=0.0, max_period=10000):
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def get_1d_rotary_pos_embed(dim, pos, theta=10000.0):
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32, device=pos.device)[: dim // 2] / dim))
    freqs = torch.outer(pos, freqs)
    return (freqs.cos().repeat_interleave(2, dim=1).float(),
            freqs.sin().repeat_interleave(2, dim=1).float())


def get_3d_rotary_pos_embed(embed_dim, crops_coords, grid_size, temporal_size, theta=10000,
                            use_real=True, grid_type="slice", max_size=None, device=None):
    """grid_type="slice" branch only (the one /root/reference/inference_script.py:382-390 uses)."""
    assert grid_type == "slice" and use_real
    max_h, max_w = max_size
    gh, gw = grid_size
    grid_h = torch.arange(max_h, device=device, dtype=torch.float32)
    grid_w = torch.arange(max_w, device=device, dtype=torch.float32)
    grid_t = torch.arange(temporal_size, device=device, dtype=torch.float32)
    dim_t, dim_h, dim_w = embed_dim // 4, embed_dim // 8 * 3, embed_dim // 8 * 3
    t_cos, t_sin = get_1d_rotary_pos_embed(dim_t, grid_t, theta)
    h_cos, h_sin = get_1d_rotary_pos_embed(dim_h, grid_h, theta)
    w_cos, w_sin = get_1d_rotary_pos_embed(dim_w, grid_w, theta)
    t_cos, t_sin = t_cos[:temporal_size], t_sin[:temporal_size]
    h_cos, h_sin = h_cos[:gh], h_sin[:gh]
    w_cos, w_sin = w_cos[:gw], w_sin[:gw]

    def combine(ft, fh, fw):
        ft = ft[:, None, None, :].expand(-1, gh, gw, -1)
        fh = fh[None, :, None, :].expand(temporal_size, -1, gw, -1)
        fw = fw[None, None, :, :].expand(temporal_size, gh, -1, -1)
        return torch.cat([ft, fh, fw], dim=-1).reshape(temporal_size * gh * gw, -1)

    return combine(t_cos, h_cos, w_cos), combine(t_sin, h_sin, w_sin)


def apply_rotary_emb(x, freqs_cis):
    cos, sin = freqs_cis
    cos, sin = cos[None, None].to(x.device), sin[None, None].to(x.device)
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    return (x.float() * cos + x_rot.float() * sin).to(x.dtype)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class PatchEmbed(nn.Module):
    def __init__(self, p, p_t, cin, dim, text_dim, bias):
        super().__init__()
        self.p, self.p_t = p, p_t
        # diffusers CogVideoXPatchEmbed: the CogVideoX-1.5 branch (patch_size_t is not None) builds
        # nn.Linear(in*p*p*p_t, dim) with the DEFAULT bias (the `bias` / config.patch_bias argument only reaches the
        # 1.0 Conv2d branch), and the released 1.5 checkpoints ship a trained patch_embed.proj.bias.
        self.proj = nn.Linear(cin * p * p * p_t, dim, bias=True if p_t is not None else bias)
        self.text_proj = nn.Linear(text_dim, dim)

    def forward(self, text, image):
        text = self.text_proj(text)
        b, f, c, h, w = image.shape
        p, pt = self.p, self.p_t
        x = image.permute(0, 1, 3, 4, 2).reshape(b, f // pt, pt, h // p, p, w // p, p, c)
        x = x.permute(0, 1, 3, 5, 7, 2, 4, 6).flatten(4, 7).flatten(1, 3)
        x = self.proj(x)
        return torch.cat([text, x], dim=1).contiguous()


# ---- block ---------------------------------------------------------------------------------
class LayerNormZero(nn.Module):
    def __init__(self, cond_dim, dim, eps):
        super().__init__()
        self.linear = nn.Linear(cond_dim, 6 * dim)
        self.norm = nn.LayerNorm(dim, eps=eps, elementwise_affine=True)

    def forward(self, h, e, temb):
        shift, scale, gate, eshift, escale, egate = self.linear(F.silu(temb)).chunk(6, dim=1)
        h = self.norm(h) * (1 + scale)[:, None, :] + shift[:, None, :]
        e = self.norm(e) * (1 + escale)[:, None, :] + eshift[:, None, :]
        return h, e, gate[:, None, :], egate[:, None, :]


class Attention(nn.Module):
    def __init__(self, dim, heads, head_dim):
        super().__init__()
        self.heads = heads
        inner = heads * head_dim
        self.to_q = nn.Linear(dim, inner)
        self.to_k = nn.Linear(dim, inner)
        self.to_v = nn.Linear(dim, inner)
        self.norm_q = nn.LayerNorm(head_dim, eps=1e-6, elementwise_affine=True)
        self.norm_k = nn.LayerNorm(head_dim, eps=1e-6, elementwise_affine=True)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Identity()])

    def forward(self, h, e, rope):
        tl = e.size(1)
        x = torch.cat([e, h], dim=1)
        b = x.shape[0]
        q, k, v = self.to_q(x), self.to_k(x), self.to_v(x)
        hd = q.shape[-1] // self.heads
        q = q.view(b, -1, self.heads, hd).transpose(1, 2)
        k = k.view(b, -1, self.heads, hd).transpose(1, 2)
        v = v.view(b, -1, self.heads, hd).transpose(1, 2)
        q, k = self.norm_q(q), self.norm_k(k)
        if rope is not None:
            q[:, :, tl:] = apply_rotary_emb(q[:, :, tl:], rope)
            k[:, :, tl:] = apply_rotary_emb(k[:, :, tl:], rope)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, self.heads * hd)
        o = self.to_out[0](o)
        e_out, h_out = o.split([tl, o.size(1) - tl], dim=1)
        return h_out, e_out


class GELUProj(nn.Module):
    def __init__(self, din, dout):
        super().__init__()
        self.proj = nn.Linear(din, dout)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class FeedForward(nn.Module):
    def __init__(self, dim, mult):
        super().__init__()
        self.net = nn.ModuleList([GELUProj(dim, dim * mult), nn.Identity(), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class CogVideoXBlock(nn.Module):
    def __init__(self, dim, heads, head_dim, time_dim, eps, ff_mult):
        super().__init__()
        self.norm1 = LayerNormZero(time_dim, dim, eps)
        self.attn1 = Attention(dim, heads, head_dim)
        self.norm2 = LayerNormZero(time_dim, dim, eps)
        self.ff = FeedForward(dim, ff_mult)

    def forward(self, h, e, temb, rope):
        tl = e.size(1)
        nh, ne, gate, egate = self.norm1(h, e, temb)
        ah, ae = self.attn1(nh, ne, rope)
        h = h + gate * ah
        e = e + egate * ae
        nh, ne, gate, egate = self.norm2(h, e, temb)
        ff = self.ff(torch.cat([ne, nh], dim=1))
        h = h + gate * ff[:, tl:]
        e = e + egate * ff[:, :tl]
        return h, e


class AdaLayerNorm(nn.Module):
    def __init__(self, cond_dim, out_dim, eps):
        super().__init__()
        self.linear = nn.Linear(cond_dim, out_dim)
        self.norm = nn.LayerNorm(out_dim // 2, eps=eps, elementwise_affine=True)

    def forward(self, x, temb):
        shift, scale = self.linear(F.silu(temb)).chunk(2, dim=1)   # chunk_dim=1: shift first
        return self.norm(x) * (1 + scale[:, None, :]) + shift[:, None, :]


class OracleCogVideoXTransformer3DModel(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        cfg = dict(DIT_CONFIG)
        cfg.update(kw)
        self.config = SimpleNamespace(**cfg)
        c = self.config
        dim = c.num_attention_heads * c.attention_head_dim
        self.patch_embed = PatchEmbed(c.patch_size, c.patch_size_t, c.in_channels, dim, c.text_embed_dim,
                                      c.patch_bias)
        self.time_embedding = TimestepEmbedding(dim, c.time_embed_dim)
        self.transformer_blocks = nn.ModuleList(
            [CogVideoXBlock(dim, c.num_attention_heads, c.attention_head_dim, c.time_embed_dim, c.norm_eps,
                            c.ff_mult) for _ in range(c.num_layers)])
        self.norm_final = nn.LayerNorm(dim, c.norm_eps, elementwise_affine=True)
        self.norm_out = AdaLayerNorm(c.time_embed_dim, 2 * dim, c.norm_eps)
        self.proj_out = nn.Linear(dim, c.patch_size * c.patch_size * c.patch_size_t * c.out_channels)
        self.dim = dim

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def forward(self, hidden_states, encoder_hidden_states, timestep, image_rotary_emb=None,
                return_dict=False, **_):
        c = self.config
        b, f, ch, hh, ww = hidden_states.shape
        t_emb = get_timestep_embedding(timestep, self.dim, c.flip_sin_to_cos, c.freq_shift)
        t_emb = t_emb.to(dtype=hidden_states.dtype)
        emb = self.time_embedding(t_emb)
        x = self.patch_embed(encoder_hidden_states, hidden_states)
        tl = encoder_hidden_states.shape[1]
        e, h = x[:, :tl], x[:, tl:]
        for blk in self.transformer_blocks:
            h, e = blk(h, e, emb, image_rotary_emb)
        h = self.norm_final(h)
        h = self.norm_out(h, emb)
        h = self.proj_out(h)
        p, pt = c.patch_size, c.patch_size_t
        out = h.reshape(b, (f + pt - 1) // pt, hh // p, ww // p, -1, pt, p, p)
        out = out.permute(0, 1, 5, 4, 2, 6, 3, 7).flatten(6, 7).flatten(4, 5).flatten(1, 2)
        return (out,)
