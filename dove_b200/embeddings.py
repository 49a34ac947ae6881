"""3-D rotary position tables (mirror of `diffusers.models.embeddings.get_3d_rotary_pos_embed`, the symbol
/root/reference/inference_script.py:17 imports and :382-390 calls with grid_type="slice").

Index bookkeeping is bit-exact: table row = (t * grid_h + y) * grid_w + x; columns 0..15 carry the temporal
frequencies, 16..39 height, 40..63 width, each frequency repeated twice (interleaved pairs);
freq_i = theta^(-2i/dim_axis).  Tables are fp32 and computed once per (T, h, w) on the target device (they are
inputs of the `dove_qk_norm_rope_bf16` kernel), then cached.
"""
from __future__ import annotations

import torch

_cache: dict = {}


def _axis_table(dim: int, n: int, theta: float, device):
    pos = torch.arange(n, device=device, dtype=torch.float32)
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32, device=device)[: dim // 2] / dim))
    ang = torch.outer(pos, inv)
    return (ang.cos().repeat_interleave(2, dim=1).float(), ang.sin().repeat_interleave(2, dim=1).float())


def get_3d_rotary_pos_embed(embed_dim, crops_coords, grid_size, temporal_size, theta: int = 10000,
                            use_real: bool = True, grid_type: str = "slice", max_size=None, device=None):
    if grid_type != "slice" or not use_real:
        raise NotImplementedError("dove_b200 implements the grid_type='slice' real-valued tables DOVE uses")
    gh, gw = grid_size
    mh, mw = max_size if max_size is not None else grid_size
    key = (embed_dim, gh, gw, mh, mw, temporal_size, theta, str(device))
    if key in _cache:
        return _cache[key]
    dt, dh, dw = embed_dim // 4, embed_dim // 8 * 3, embed_dim // 8 * 3
    tables = []
    (tc, ts), (hc, hs), (wc, ws) = (_axis_table(dt, temporal_size, theta, device), _axis_table(dh, mh, theta, device),
                                    _axis_table(dw, mw, theta, device))
    for ft, fh, fw in ((tc, hc, wc), (ts, hs, ws)):
        ft = ft[:temporal_size, None, None, :].expand(-1, gh, gw, -1)
        fh = fh[None, :gh, None, :].expand(temporal_size, -1, gw, -1)
        fw = fw[None, None, :gw, :].expand(temporal_size, gh, -1, -1)
        tables.append(torch.cat([ft, fh, fw], dim=-1).reshape(temporal_size * gh * gw, embed_dim).contiguous())
    _cache[key] = (tables[0], tables[1])
    return _cache[key]
