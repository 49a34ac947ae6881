"""Builds matching (oracle fp32, oracle bf16, dove_b200) model triples from one bf16-rounded state dict."""
import copy

import torch

from dove_b200.weights import dit_param_spec, init_state_dict, vae_param_spec
from oracle.dit import OracleCogVideoXTransformer3DModel
from oracle.pipeline import OraclePipe
from oracle.vae import OracleAutoencoderKLCogVideoX

SMALL_DIT = dict(num_layers=2, num_attention_heads=4)           # dim 256
WIDE_DIT = dict(num_layers=2)                                   # 48 heads, dim 3072 (full width, shrunken depth)


def state_dicts(dit_cfg, seed=1234):
    vsd = init_state_dict(vae_param_spec(), seed, "cpu", torch.bfloat16)
    dsd = init_state_dict(dit_param_spec(dit_cfg), seed, "cpu", torch.bfloat16)
    return vsd, dsd


def oracle_vae(vsd, device, dtype):
    m = OracleAutoencoderKLCogVideoX()
    m.load_state_dict({k: v.float() for k, v in vsd.items()})
    return m.to(device=device, dtype=dtype).eval()


def oracle_dit(dsd, dit_cfg, device, dtype):
    m = OracleCogVideoXTransformer3DModel(**dit_cfg)
    m.load_state_dict({k: v.float() for k, v in dsd.items()})
    return m.to(device=device, dtype=dtype).eval()


def oracle_pipe(vsd, dsd, dit_cfg, device, dtype):
    return OraclePipe(oracle_vae(vsd, device, dtype), oracle_dit(dsd, dit_cfg, device, dtype))


def b200_pipe(vsd, dsd, dit_cfg, device="cuda"):
    from dove_b200.pipeline import CogVideoXPipeline
    from dove_b200.transformer import CogVideoXTransformer3DModel
    from dove_b200.vae import AutoencoderKLCogVideoX
    return CogVideoXPipeline(AutoencoderKLCogVideoX(vsd, None, device),
                             CogVideoXTransformer3DModel(dsd, dit_cfg, device))


def prompt_embedding(device="cpu"):
    """The reference's shipped empty-prompt T5 embedding [226, 4096] bf16 (tests/golden fixture, ref :580-590)."""
    from pathlib import Path
    from dove_b200.pipeline import load_prompt_embedding
    return load_prompt_embedding(Path(__file__).resolve().parent / "golden" / "empty_prompt_embedding.safetensors").to(device)
