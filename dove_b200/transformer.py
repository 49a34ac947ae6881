"""B200-native CogVideoX-1.5 DiT forward (host orchestration over libdove_b200 kernels).

Mirror of the diffusers `CogVideoXTransformer3DModel` call surface the reference uses
(ref: /root/reference/inference_script.py:483-489; config attributes read at :411, :373-383, :434, :478).

Token layout in HBM: ONE residual-stream buffer x [N, D] bf16 with the 226 text tokens first and the
T/2*h/2*w/2 video tokens after them (the reference keeps two tensors and concatenates them 3x per block; all
ops are row-wise so a single buffer with a `split_row` is equivalent).  Per block (6 kernel launches):
  layernorm_mod (LN + AdaLN-Zero modulate, text/video segments)
  -> fused QKV GEMM [N, 3D] (tcgen05) with per-head q/k LayerNorm + 3-D RoPE in its epilogue
  -> attention (tcgen05 flash attention)
  -> out-proj GEMM with gate*y + residual epilogue                -> layernorm_mod
  -> FF1 GEMM with GELU-tanh epilogue -> FF2 GEMM with gate*y + residual epilogue.
The timestep is the constant 399 in DOVE (ref: :459-464), so the timestep embedding and all 42x2 modulation
6-tuples are input independent: they are computed once per distinct timestep value with the same bf16 op
sequence (dove_gemv_bf16) and cached.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch

from . import _lib as L
from .weights import DIT_CONFIG

BF = torch.bfloat16


class _Block:
    pass


class CogVideoXTransformer3DModel:
    def __init__(self, state_dict, config=None, device="cuda"):
        cfg = dict(DIT_CONFIG)
        cfg.update(config or {})
        self.config = SimpleNamespace(**cfg)
        c = self.config
        if c.attention_head_dim != 64:
            raise NotImplementedError("attention kernel is specialised for head_dim 64 (CogVideoX)")
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise L.DoveError("dove_b200.CogVideoXTransformer3DModel runs on CUDA (sm_100a) only")
        L.init(self._device.index if self._device.index is not None else torch.cuda.current_device())
        self.heads = c.num_attention_heads
        self.dim = self.heads * c.attention_head_dim
        if self.dim % 256 != 0:
            raise NotImplementedError("inner dim must be a multiple of 256")
        self._mods = {}
        self._rope_t = None
        self.fuse_qk_norm_rope = True      # False: separate dove_qk_norm_rope_bf16 pass (kept for A/B tests)
        self._build(state_dict)

    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return BF

    def to(self, *a, **k):
        return self

    def _build(self, sd):
        dev = self._device
        used = set()

        def g(k):
            if k not in sd:
                raise KeyError(f"transformer checkpoint is missing '{k}' (strict loading: no silent defaults)")
            used.add(k)
            return sd[k].to(device=dev, dtype=BF).contiguous()

        c = self.config
        self.w_patch = g("patch_embed.proj.weight")                       # [D, 128]
        # CogVideoX-1.5 (patch_size_t set): nn.Linear with its default bias regardless of config.patch_bias
        # (diffusers CogVideoXPatchEmbed); the released checkpoints carry the tensor.
        self.b_patch = g("patch_embed.proj.bias") if (c.patch_size_t is not None or c.patch_bias) else None
        self.w_text, self.b_text = g("patch_embed.text_proj.weight"), g("patch_embed.text_proj.bias")
        self.te1 = (g("time_embedding.linear_1.weight"), g("time_embedding.linear_1.bias"))
        self.te2 = (g("time_embedding.linear_2.weight"), g("time_embedding.linear_2.bias"))
        self.blocks = []
        for i in range(c.num_layers):
            p = f"transformer_blocks.{i}"
            b = _Block()
            b.mod1 = (g(f"{p}.norm1.linear.weight"), g(f"{p}.norm1.linear.bias"))
            b.ln1 = (g(f"{p}.norm1.norm.weight"), g(f"{p}.norm1.norm.bias"))
            b.wqkv = torch.cat([g(f"{p}.attn1.to_q.weight"), g(f"{p}.attn1.to_k.weight"),
                                g(f"{p}.attn1.to_v.weight")], 0).contiguous()
            b.bqkv = torch.cat([g(f"{p}.attn1.to_q.bias"), g(f"{p}.attn1.to_k.bias"),
                                g(f"{p}.attn1.to_v.bias")], 0).contiguous()
            b.qn = (g(f"{p}.attn1.norm_q.weight"), g(f"{p}.attn1.norm_q.bias"))
            b.kn = (g(f"{p}.attn1.norm_k.weight"), g(f"{p}.attn1.norm_k.bias"))
            b.wo, b.bo = g(f"{p}.attn1.to_out.0.weight"), g(f"{p}.attn1.to_out.0.bias")
            b.mod2 = (g(f"{p}.norm2.linear.weight"), g(f"{p}.norm2.linear.bias"))
            b.ln2 = (g(f"{p}.norm2.norm.weight"), g(f"{p}.norm2.norm.bias"))
            b.w1, b.b1 = g(f"{p}.ff.net.0.proj.weight"), g(f"{p}.ff.net.0.proj.bias")
            b.w2, b.b2 = g(f"{p}.ff.net.2.weight"), g(f"{p}.ff.net.2.bias")
            self.blocks.append(b)
        self.ln_final = (g("norm_final.weight"), g("norm_final.bias"))
        self.mod_out = (g("norm_out.linear.weight"), g("norm_out.linear.bias"))
        self.ln_out = (g("norm_out.norm.weight"), g("norm_out.norm.bias"))
        self.w_proj, self.b_proj = g("proj_out.weight"), g("proj_out.bias")
        extra = sorted(set(sd) - used)
        if extra:      # a tensor this implementation does not consume would be a silently dropped parameter
            raise KeyError(f"transformer checkpoint has {len(extra)} unconsumed tensors, e.g. {extra[:4]}")
        self._text_cache = {}

    # ---- timestep-dependent constants ---------------------------------------------------------------
    def _modulation(self, t: int):
        """temb = time_embedding(Timesteps(t).to(bf16)); per block chunk(6) of Linear(SiLU(temb))."""
        if t in self._mods:
            return self._mods[t]
        c, dev, D = self.config, self._device, self.dim
        half = D // 2
        # Timesteps(): fp32 sinusoid, flip_sin_to_cos -> [cos | sin], then cast to bf16 (tiny, load-time only)
        exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=dev) / (half - c.freq_shift)
        ang = float(t) * torch.exp(exponent)
        emb = torch.cat([torch.sin(ang), torch.cos(ang)])
        if c.flip_sin_to_cos:
            emb = torch.cat([emb[half:], emb[:half]])
        t_emb = emb.to(BF).contiguous()
        h1 = torch.empty(c.time_embed_dim, dtype=BF, device=dev)
        temb = torch.empty(c.time_embed_dim, dtype=BF, device=dev)
        L.gemv(t_emb, self.te1[0], self.te1[1], h1)
        L.gemv(h1, self.te2[0], self.te2[1], temb, silu_in=True)
        mods = []
        for b in self.blocks:
            m1 = torch.empty(6 * D, dtype=BF, device=dev)
            m2 = torch.empty(6 * D, dtype=BF, device=dev)
            L.gemv(temb, b.mod1[0], b.mod1[1], m1, silu_in=True)
            L.gemv(temb, b.mod2[0], b.mod2[1], m2, silu_in=True)
            # chunk order: shift, scale, gate, enc_shift, enc_scale, enc_gate
            mods.append((m1.view(6, D), m2.view(6, D)))
        mo = torch.empty(2 * D, dtype=BF, device=dev)
        L.gemv(temb, self.mod_out[0], self.mod_out[1], mo, silu_in=True)
        self._mods[t] = (mods, mo.view(2, D))       # norm_out chunk order: shift, scale
        return self._mods[t]

    def _text_tokens(self, text):
        """text_proj(prompt embedding).  DOVE always feeds the SAME constant empty-prompt embedding (ref
        inference_script.py:580-590, :423-428), so the projection is computed once per distinct tensor and cached
        (keyed on the tensor's storage + version counter, so an in-place edit invalidates it)."""
        key = (text.data_ptr(), text._version, tuple(text.shape))
        hit = self._text_cache.get(key)
        if hit is None:
            out = torch.empty(text.shape[0], self.dim, dtype=BF, device=self._device)
            L.gemm(text.contiguous(), self.w_text, out, self.b_text)
            self._text_cache.clear()
            self._text_cache[key] = hit = (text, out)      # holding `text` keeps the pointer key valid
        return hit[1]

    # ---- forward ------------------------------------------------------------------------------------
    def forward_tokens(self, latent, text, t: int, rope):
        """latent [F,16,h,w] bf16 (F even), text [n_text, 4096] bf16, rope (cos, sin) fp32 [Nv,64]
        -> proj_out tokens [Nv, 128] bf16 (patch feature order (C, pt, ph, pw))."""
        c, dev, D, H = self.config, self._device, self.dim, self.heads
        F, C, h, w = latent.shape
        nt = text.shape[0]
        nv = (F // 2) * (h // 2) * (w // 2)
        N = nt + nv
        mods, mod_out = self._modulation(t)
        x = torch.empty(N, D, dtype=BF, device=dev)
        x[:nt] = self._text_tokens(text)
        tok = torch.empty(nv, C * 8, dtype=BF, device=dev)
        L.patchify(latent.contiguous(), tok)
        L.gemm(tok, self.w_patch, x[nt:], self.b_patch)
        cos, sin = (rope[0].contiguous(), rope[1].contiguous()) if rope is not None else (None, None)
        n1 = torch.empty(N, D, dtype=BF, device=dev)
        qkv = torch.empty(N, 3 * D, dtype=BF, device=dev)
        att = torch.empty(N, D, dtype=BF, device=dev)
        ff = torch.empty(N, c.ff_mult * D, dtype=BF, device=dev)
        scale = 1.0 / math.sqrt(c.attention_head_dim)
        fuse_qk = self.fuse_qk_norm_rope and D % 256 == 0
        cos_t = sin_t = None
        if fuse_qk and rope is not None:       # transposed tables of the fused epilogue, cached per table object
            key = (cos.data_ptr(), tuple(cos.shape))
            if self._rope_t is None or self._rope_t[0] != key:
                self._rope_t = (key, cos, L.rope_tables_transposed(cos, sin))
            cos_t, sin_t = self._rope_t[2]
        for b, (m1, m2) in zip(self.blocks, mods):
            L.layernorm_mod(x, n1, b.ln1[0], b.ln1[1], c.norm_eps, m1[4], m1[3], m1[1], m1[0], nt)
            if fuse_qk:     # QKV projection with per-head q/k LayerNorm + RoPE in the GEMM epilogue
                L.gemm_qkv_norm_rope(n1, b.wqkv, qkv, b.bqkv, H, b.qn[0], b.qn[1], b.kn[0], b.kn[1], 1e-6, cos_t, sin_t,
                                     nt if rope is not None else N)
            else:
                L.gemm(n1, b.wqkv, qkv, b.bqkv)
                L.qk_norm_rope(qkv, H, b.qn[0], b.qn[1], b.kn[0], b.kn[1], 1e-6, cos, sin, nt if rope is not None else N)
            L.attention(qkv, att, H, scale)
            L.gemm(att, b.wo, x, b.bo, L.EPI_GATED_RES, aux=x, gate0=m1[5], gate1=m1[2], split_row=nt)
            L.layernorm_mod(x, n1, b.ln2[0], b.ln2[1], c.norm_eps, m2[4], m2[3], m2[1], m2[0], nt)
            L.gemm(n1, b.w1, ff, b.b1, L.EPI_GELU_TANH)
            L.gemm(ff, b.w2, x, b.b2, L.EPI_GATED_RES, aux=x, gate0=m2[5], gate1=m2[2], split_row=nt)
        xv = x[nt:]
        nf = n1[:nv]
        L.layernorm_mod(xv, nf, self.ln_final[0], self.ln_final[1], c.norm_eps)
        no = att[:nv]
        L.layernorm_mod(nf, no, self.ln_out[0], self.ln_out[1], c.norm_eps, mod_out[1], mod_out[0], None, None, 0)
        out = torch.empty(nv, self.w_proj.shape[0], dtype=BF, device=dev)
        L.gemm(no, self.w_proj, out, self.b_proj)
        return out

    def __call__(self, hidden_states, encoder_hidden_states, timestep, image_rotary_emb=None, return_dict=False,
                 **kwargs):
        """hidden_states [1,F,16,h,w]; encoder_hidden_states [1,n_text,4096]; timestep LongTensor[1] -> (pred,)"""
        assert hidden_states.shape[0] == 1, "batch 1 (DOVE processes one chunk x tile unit at a time)"
        t = int(timestep.reshape(-1)[0].item()) if torch.is_tensor(timestep) else int(timestep)
        lat = hidden_states[0].to(device=self._device, dtype=BF).contiguous()
        txt = encoder_hidden_states[0].to(device=self._device, dtype=BF).contiguous()
        F, C, h, w = lat.shape
        tok = self.forward_tokens(lat, txt, t, image_rotary_emb)
        pred = torch.empty_like(lat)
        L.unpatchify_velocity(tok, None, None, pred, F, C, h, w, 0.0, 0.0)
        out = pred[None]
        return (out,) if not return_dict else SimpleNamespace(sample=out)
