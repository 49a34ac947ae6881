"""Model configurations and weights for the DOVE hot path (CogVideoX-1.5-5B DiT + CogVideoX 3-D VAE).

* `VAE_CONFIG`, `DIT_CONFIG`, `SCHED_CONFIG`: the `config.json` values DOVE keeps from the base checkpoint
  (ref: /root/reference/finetune/scripts/prepare_sft_ckpt.py:49-67; values in SURVEY.md appendix A.0).
* `vae_param_spec` / `dit_param_spec`: every parameter (diffusers state-dict key, shape, kind) — the layout
  `CogVideoXPipeline.from_pretrained` (ref: inference_script.py:613) reads from `diffusion_pytorch_model*.safetensors`.
* `init_state_dict`: deterministic random initialisation (no network, no released weights): conv/linear weights
  N(0, 1/fan_in), biases N(0, 0.02^2), norm scales 1 + 0.1 N, norm shifts 0.05 N — every bias / affine / gate
  path is exercised (AdaLN-Zero gates are non-zero, otherwise each DiT block would be the identity).
* `load_safetensors_dir`: real-weight loading from a diffusers directory layout (sharded safetensors + index).
"""
from __future__ import annotations

import json
import math
import zlib
from pathlib import Path

import torch

VAE_CONFIG = dict(
    in_channels=3, out_channels=3, latent_channels=16,
    block_out_channels=(128, 256, 256, 512), layers_per_block=3,
    norm_eps=1e-6, norm_num_groups=32, temporal_compression_ratio=4,
    sample_height=480, sample_width=720, scaling_factor=0.7,
)
DIT_CONFIG = dict(
    num_attention_heads=48, attention_head_dim=64, in_channels=16, out_channels=16,
    flip_sin_to_cos=True, freq_shift=0, time_embed_dim=512, text_embed_dim=4096, num_layers=42,
    patch_size=2, patch_size_t=2, patch_bias=False, max_text_seq_length=226,
    norm_eps=1e-5, use_rotary_positional_embeddings=True, use_learned_positional_embeddings=False,
    sample_height=300, sample_width=300, sample_frames=81, temporal_compression_ratio=4, ff_mult=4,
)
SCHED_CONFIG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                    beta_schedule="scaled_linear", prediction_type="v_prediction",
                    rescale_betas_zero_snr=True, snr_shift_scale=1.0, timestep_spacing="trailing")


def _conv3(name, cin, cout, k):
    return [(f"{name}.conv.weight", (cout, cin, k, k, k), "w"), (f"{name}.conv.bias", (cout,), "b")]


def _norm(name, c):
    return [(f"{name}.weight", (c,), "nw"), (f"{name}.bias", (c,), "nb")]


def _spatial_norm(name, c, zc):
    return _norm(f"{name}.norm_layer", c) + _conv3(f"{name}.conv_y", zc, c, 1) + _conv3(f"{name}.conv_b", zc, c, 1)


def _resnet(name, cin, cout, zc=None):
    out = []
    out += _spatial_norm(f"{name}.norm1", cin, zc) if zc else _norm(f"{name}.norm1", cin)
    out += _spatial_norm(f"{name}.norm2", cout, zc) if zc else _norm(f"{name}.norm2", cout)
    out += _conv3(f"{name}.conv1", cin, cout, 3) + _conv3(f"{name}.conv2", cout, cout, 3)
    if cin != cout:
        out += [(f"{name}.conv_shortcut.weight", (cout, cin, 1, 1, 1), "w"), (f"{name}.conv_shortcut.bias", (cout,), "b")]
    return out


def vae_param_spec(cfg=None):
    c = dict(VAE_CONFIG)
    c.update(cfg or {})
    boc, lpb, zc = list(c["block_out_channels"]), c["layers_per_block"], c["latent_channels"]
    s = _conv3("encoder.conv_in", c["in_channels"], boc[0], 3)
    oc = boc[0]
    for i, ch in enumerate(boc):
        ic, oc = oc, ch
        for j in range(lpb):
            s += _resnet(f"encoder.down_blocks.{i}.resnets.{j}", ic if j == 0 else oc, oc)
        if i != len(boc) - 1:
            s += [(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (oc, oc, 3, 3), "w"),
                  (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (oc,), "b")]
    for j in range(2):
        s += _resnet(f"encoder.mid_block.resnets.{j}", boc[-1], boc[-1])
    s += _norm("encoder.norm_out", boc[-1]) + _conv3("encoder.conv_out", boc[-1], 2 * zc, 3)
    rev = boc[::-1]
    s += _conv3("decoder.conv_in", zc, rev[0], 3)
    for j in range(2):
        s += _resnet(f"decoder.mid_block.resnets.{j}", rev[0], rev[0], zc)
    oc = rev[0]
    for i, ch in enumerate(rev):
        pc, oc = oc, ch
        for j in range(lpb + 1):
            s += _resnet(f"decoder.up_blocks.{i}.resnets.{j}", pc if j == 0 else oc, oc, zc)
        if i != len(rev) - 1:
            s += [(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight", (oc, oc, 3, 3), "w"),
                  (f"decoder.up_blocks.{i}.upsamplers.0.conv.bias", (oc,), "b")]
    s += _spatial_norm("decoder.norm_out", rev[-1], zc) + _conv3("decoder.conv_out", rev[-1], c["out_channels"], 3)
    return s


def _linear(name, din, dout, bias=True):
    return [(f"{name}.weight", (dout, din), "w")] + ([(f"{name}.bias", (dout,), "b")] if bias else [])


def dit_param_spec(cfg=None):
    c = dict(DIT_CONFIG)
    c.update(cfg or {})
    hd = c["attention_head_dim"]
    dim = c["num_attention_heads"] * hd
    te = c["time_embed_dim"]
    pfeat = c["in_channels"] * c["patch_size"] ** 2 * c["patch_size_t"]
    # CogVideoX-1.5 (patch_size_t set): diffusers builds the patch projection as nn.Linear with its default bias,
    # whatever config.patch_bias says (it only applies to the 1.0 Conv2d branch) — the checkpoint has the tensor.
    has_pbias = c["patch_size_t"] is not None or c["patch_bias"]
    s = _linear("patch_embed.proj", pfeat, dim, has_pbias) + _linear("patch_embed.text_proj", c["text_embed_dim"], dim)
    s += _linear("time_embedding.linear_1", dim, te) + _linear("time_embedding.linear_2", te, te)
    for i in range(c["num_layers"]):
        b = f"transformer_blocks.{i}"
        s += _linear(f"{b}.norm1.linear", te, 6 * dim) + _norm(f"{b}.norm1.norm", dim)
        s += _linear(f"{b}.attn1.to_q", dim, dim) + _linear(f"{b}.attn1.to_k", dim, dim) + _linear(f"{b}.attn1.to_v", dim, dim)
        s += _norm(f"{b}.attn1.norm_q", hd) + _norm(f"{b}.attn1.norm_k", hd)
        s += _linear(f"{b}.attn1.to_out.0", dim, dim)
        s += _linear(f"{b}.norm2.linear", te, 6 * dim) + _norm(f"{b}.norm2.norm", dim)
        s += _linear(f"{b}.ff.net.0.proj", dim, c["ff_mult"] * dim) + _linear(f"{b}.ff.net.2", c["ff_mult"] * dim, dim)
    s += _norm("norm_final", dim)
    s += _linear("norm_out.linear", te, 2 * dim) + _norm("norm_out.norm", dim)
    s += _linear("proj_out", dim, pfeat)
    return s


def param_count(spec) -> int:
    return sum(math.prod(shape) for _, shape, _ in spec)


def init_state_dict(spec, seed: int = 1234, device="cpu", dtype=torch.float32) -> dict:
    """Deterministic per-tensor init (seeded by crc32(name) ^ seed, so order/device-placement independent
    within one device type)."""
    out = {}
    dev = torch.device(device)
    for name, shape, kind in spec:
        g = torch.Generator(device=dev).manual_seed((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
        r = torch.randn(shape, generator=g, device=dev, dtype=torch.float32)
        if kind == "w":
            fan_in = math.prod(shape[1:])
            t = r * (1.0 / math.sqrt(fan_in))
        elif kind == "b":
            t = r * 0.02
        elif kind == "nw":
            t = 1.0 + 0.1 * r
        else:
            t = 0.05 * r
        out[name] = t.to(dtype)
    return out


def load_safetensors_dir(path) -> dict:
    """Read a diffusers component directory (`diffusion_pytorch_model.safetensors` or the sharded
    `diffusion_pytorch_model-0000x-of-0000y.safetensors` + `.index.json` layout written by
    ref: finetune/scripts/prepare_sft_ckpt.py:17-41)."""
    from safetensors.torch import load_file
    path = Path(path)
    idx = path / "diffusion_pytorch_model.safetensors.index.json"
    files = sorted(set(json.loads(idx.read_text())["weight_map"].values())) if idx.exists() else \
        ["diffusion_pytorch_model.safetensors"]
    sd = {}
    for f in files:
        sd.update(load_file(str(path / f)))
    return sd
