"""Executes the REFERENCE's own `process_video` source (ast-extracted from /root/reference/inference_script.py,
unmodified) against the oracle pipe and checks that `oracle.pipeline.oracle_process_video` — the restatement the
GPU parity tests compare against — returns the same tensor bit for bit.  Build-container only (the GPU box has no
/root/reference); CPU, fp32, tiny DiT."""
import ast
from pathlib import Path
from typing import Dict, Tuple

import pytest
import torch

REF = Path("/root/reference/inference_script.py")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference checkout not present")


def _reference_namespace():
    from oracle.dit import get_3d_rotary_pos_embed
    tree = ast.parse(REF.read_text())
    ns = dict(torch=torch, Dict=Dict, Tuple=Tuple, CogVideoXPipeline=object,
              get_3d_rotary_pos_embed=get_3d_rotary_pos_embed)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("no_grad", "prepare_rotary_positional_embeddings",
                                                               "process_video"):
            exec(compile(ast.Module([node], []), str(REF), "exec"), ns)
    return ns


def test_reference_process_video_runs_on_oracle_pipe():
    import models
    from dove_b200.pipeline import synthetic_prompt_embedding
    from oracle.pipeline import oracle_process_video
    cfg = dict(num_layers=1, num_attention_heads=4)
    vsd, dsd = models.state_dicts(cfg)
    pipe = models.oracle_pipe(vsd, dsd, cfg, "cpu", torch.float32)
    ns = _reference_namespace()
    emb = synthetic_prompt_embedding().float()
    torch.manual_seed(0)
    video = torch.rand(1, 3, 9, 16, 16) * 2 - 1
    torch.manual_seed(42)
    ref_out = ns["process_video"](pipe=pipe, video=video, prompt="", noise_step=0, sr_noise_step=399,
                                  empty_prompt_embedding=emb)
    torch.manual_seed(42)
    ours = oracle_process_video(pipe, video, emb)
    assert ref_out.shape == (1, 3, 9, 16, 16)
    assert torch.equal(ref_out, ours)


def test_reference_loop_matches_runner():
    """The reference's chunk x tile loop + stitch (ref :682-729, restated inline with ITS bookkeeping functions)
    equals runner.super_resolve on the same unit function."""
    from dove_b200.runner import super_resolve
    tree = ast.parse(REF.read_text())
    ns = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("make_temporal_chunks", "make_spatial_tiles",
                                                               "get_valid_tile_region"):
            exec(compile(ast.Module([node], []), str(REF), "exec"), ns)
    fn = lambda u, k=None, s=None: torch.tanh(u * 1.7) * 0.5 + 0.5
    video = torch.rand(1, 3, 41, 96, 128) * 2 - 1
    chunk_len, overlap_t, tile, ov = 25, 8, (64, 64), (32, 32)
    out = torch.zeros_like(video)
    cnt = torch.zeros_like(video, dtype=torch.int)
    for (t0, t1) in ns["make_temporal_chunks"](41, chunk_len, overlap_t):
        for (h0, h1, w0, w1) in ns["make_spatial_tiles"](96, 128, tile, ov):
            r = fn(video[:, :, t0:t1, h0:h1, w0:w1])
            g = ns["get_valid_tile_region"](t0, t1, h0, h1, w0, w1, video.shape, overlap_t, ov[0], ov[1])
            out[:, :, g["out_t_start"]:g["out_t_end"], g["out_h_start"]:g["out_h_end"], g["out_w_start"]:g["out_w_end"]] = \
                r[:, :, g["valid_t_start"]:g["valid_t_end"], g["valid_h_start"]:g["valid_h_end"], g["valid_w_start"]:g["valid_w_end"]]
            cnt[:, :, g["out_t_start"]:g["out_t_end"], g["out_h_start"]:g["out_h_end"], g["out_w_start"]:g["out_w_end"]] += 1
    assert (cnt == 1).all()
    ours = super_resolve(video, fn, chunk_len=chunk_len, overlap_t=overlap_t, tile_size_hw=tile, overlap_hw=ov)
    assert torch.equal(ours, out)
