"""Integer bookkeeping of the reference's unit decomposition (bit-exact; also the multi-GPU sharding unit).

Restates ref: /root/reference/inference_script.py:249-279 (`make_temporal_chunks`), :282-329
(`make_spatial_tiles`), :332-361 (`get_valid_tile_region`), :228-232 / :238-246 (padding rules) and the
write-count self check :724-729.  Same names, argument meaning and error behaviour as the reference.
"""
from __future__ import annotations


def make_temporal_chunks(F, chunk_len, overlap_t=8):
    """[(t_start, t_end)]; chunk_len == 0 -> one chunk.  A short last chunk is merged into its predecessor."""
    if chunk_len == 0:
        return [(0, F)]
    stride = chunk_len - overlap_t
    if stride <= 0:
        raise ValueError("chunk_len must be greater than overlap")
    starts = list(range(0, F - overlap_t, stride))
    if starts[-1] + chunk_len < F:
        starts.append(F - chunk_len)
    chunks = [(s, min(s + chunk_len, F)) for s in starts]
    if len(chunks) >= 2 and chunks[-1][1] - chunks[-1][0] < chunk_len:
        end = chunks.pop()[1]
        chunks[-1] = (chunks[-1][0], end)
    return chunks


def _axis_starts(size, tile, overlap):
    stride = tile - overlap
    starts = list(range(0, size - overlap, stride))
    if not starts or starts[-1] + tile < size:
        starts.append(size - tile)
    if len(starts) >= 2 and starts[-1] + tile > size:      # merge an overhanging last tile
        starts.pop()
    return starts, stride


def make_spatial_tiles(H, W, tile_size_hw, overlap_hw=(32, 32)):
    """[(h_start, h_end, w_start, w_end)]; a tile size of 0 disables tiling."""
    th, tw = tile_size_hw
    oh, ow = overlap_hw
    if th == 0 or tw == 0:
        return [(0, H, 0, W)]
    if th - oh <= 0 or tw - ow <= 0:
        raise ValueError("Tile size must be greater than overlap")
    hs, sh = _axis_starts(H, th, oh)
    ws, sw = _axis_starts(W, tw, ow)
    tiles = []
    for h0 in hs:
        h1 = min(h0 + th, H)
        if h1 + sh > H:
            h1 = H
        for w0 in ws:
            w1 = min(w0 + tw, W)
            if w1 + sw > W:
                w1 = W
            tiles.append((h0, h1, w0, w1))
    return tiles


def get_valid_tile_region(t_start, t_end, h_start, h_end, w_start, w_end, video_shape, overlap_t, overlap_h,
                          overlap_w):
    """Interior region of a unit that is kept: half the overlap is trimmed on every interior side."""
    _, _, F, H, W = video_shape
    out = {}
    for name, s, e, full, ov in (("t", t_start, t_end, F, overlap_t), ("h", h_start, h_end, H, overlap_h),
                                 ("w", w_start, w_end, W, overlap_w)):
        n = e - s
        vs = 0 if s == 0 else ov // 2
        ve = n if e == full else n - ov // 2
        out[f"valid_{name}_start"], out[f"valid_{name}_end"] = vs, ve
        out[f"out_{name}_start"], out[f"out_{name}_end"] = s + vs, s + ve
    return out


def enumerate_units(video_shape, chunk_len, overlap_t, tile_size_hw, overlap_hw):
    """Units in the reference's loop order (ref :690-691): for chunk: for tile."""
    _, _, F, H, W = video_shape
    return [(c, t) for c in make_temporal_chunks(F, chunk_len, overlap_t)
            for t in make_spatial_tiles(H, W, tile_size_hw, overlap_hw)]


def unit_cost(unit):
    """Work proxy of a unit: voxels (frames x rows x cols).  The VAE cost is linear in it; attention adds a mildly
    super-linear term, so bigger units are placed first."""
    (t0, t1), (h0, h1, w0, w1) = unit
    return (t1 - t0) * (h1 - h0) * (w1 - w0)


def partition_units(units, world_size, balance=True):
    """Static assignment of units to ranks.  balance=True: longest-processing-time-first greedy
    (units sorted by cost, ties by index, each given to the least-loaded rank, ties by rank) — e.g. the 8 cfg-3 tiles
    (6 of 416x352 + 2 of 416x416) on 2 or 4 ranks get one wide tile per rank instead of both on the last rank.
    balance=False: plain round-robin (rank r owns units r, r+W, ...).  Deterministic: every rank computes the same map."""
    if not balance:
        return [list(range(r, len(units), world_size)) for r in range(world_size)]
    order = sorted(range(len(units)), key=lambda k: (-unit_cost(units[k]), k))
    loads = [0] * world_size
    parts = [[] for _ in range(world_size)]
    for k in order:
        r = min(range(world_size), key=lambda i: (loads[i], i))
        parts[r].append(k)
        loads[r] += unit_cost(units[k])
    return [sorted(p) for p in parts]


def frame_padding(F):
    """Frames appended (last frame repeated) so that F = 8k+1 (ref :215-221)."""
    rem = (F - 1) % 8
    return 0 if rem == 0 else 8 - rem


def spatial_padding(n):
    """Bottom/right zero padding to a multiple of 16 (ref :223-227)."""
    return (16 - n % 16) % 16
