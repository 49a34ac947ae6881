"""Clip-level driver: temporal-chunk x spatial-tile unit loop, stitching, and unit sharding over GPUs.

Mirrors the per-video hot loop of the reference (ref: /root/reference/inference_script.py:682-729):
units = chunks x tiles in loop order; each unit is super-resolved independently; only its "valid" interior
(half the overlap trimmed on interior sides, no blending) is written; every output voxel must be written exactly
once or the run aborts.  B200-first differences: the stitched output stays in HBM (one D2H at the end instead of
one synchronous D2H per unit, ref :712-717), and with world_size > 1 the units are statically partitioned
(cost-balanced, deterministic) over ranks (no data-path collective while computing) followed by ONE all-gather of the packed valid
regions — NCCL over NVLink on GPUs, gloo in the CPU tests of this host logic.

RNG: the reference draws each unit's latent noise from the global generator in loop order (seed 42 once,
ref :577), so unit k's noise depends on all earlier units.  `noise_mode="global"` reproduces that (single rank
only).  `noise_mode="per_unit"` seeds unit k with `seed + k`, which makes results independent of the rank
count; it is required (and the default) when world_size > 1.
"""
from __future__ import annotations

import torch

from .bookkeeping import enumerate_units, get_valid_tile_region, partition_units


class StitchError(RuntimeError):
    """write count != 1 somewhere (the reference prints and exit()s, ref :724-729)."""


def _unit_regions(units, video_shape, overlap_t, overlap_hw):
    regs = []
    for (t0, t1), (h0, h1, w0, w1) in units:
        regs.append(get_valid_tile_region(t0, t1, h0, h1, w0, w1, video_shape, overlap_t, overlap_hw[0],
                                          overlap_hw[1]))
    return regs


def _region_numel(r, C):
    return (C * (r["out_t_end"] - r["out_t_start"]) * (r["out_h_end"] - r["out_h_start"]) *
            (r["out_w_end"] - r["out_w_start"]))


_DTYPE_NAMES = {"bfloat16": torch.bfloat16, "float32": torch.float32, "float16": torch.float16, "uint8": torch.uint8}
def _unit_to_device(unit, dev):
    """A unit is a strided [1,C,t,h,w] view of the clip.  When the clip lives in host memory only THIS unit crosses
    PCIe (ref :407 copies per unit too): one strided DMA per channel straight out of the (pinned) clip
    (dove_h2d_box_async = cudaMemcpy3DAsync) — no staging buffer, no CPU pass over the payload, no stream sync."""
    dev = torch.device(dev)
    if unit.device.type != "cpu" or dev.type != "cuda":
        return unit.to(dev)
    B, C, t, h, w = unit.shape
    st = unit.stride()
    if B != 1 or st[4] != 1 or (t > 1 and st[2] % st[3] != 0):
        return unit.to(dev)                                # exotic layout: let torch stage it
    from . import _lib as L
    out = torch.empty((1, C, t, h, w), dtype=unit.dtype, device=dev)
    for c in range(C):
        L.h2d_box(unit[0, c], out[0, c])
    return out


def super_resolve(video, process_fn, *, chunk_len=0, overlap_t=8, tile_size_hw=(0, 0), overlap_hw=(32, 32),
                  out_device=None, out_dtype=None, group=None, noise_mode=None, seed=42, timings=None):
    """video: [1,3,F,H,W] float in [-1,1] (already x4-upscaled and padded, ref :670-679), on the host or the GPU.
    process_fn(unit_video, unit_index, generator_seed_or_None) -> [1,3,t,h,w] tensor in [0,1] (or uint8 0..255);
    it may declare `.out_dtype` / `.device` attributes (make_process_fn does) so that ranks WITHOUT units agree with
    the working ranks on the gather buffer's dtype and device.
    Returns the stitched [1,3,F,H,W] clip (on every rank when sharded).  `timings` (dict, optional) receives
    compute_ms / collective_ms / units measured with CUDA events on this rank."""
    import torch.distributed as dist
    sharded = group is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
    rank = dist.get_rank(group) if sharded else 0
    world = dist.get_world_size(group) if sharded else 1
    if noise_mode is None:
        noise_mode = "per_unit" if world > 1 else "global"
    if world > 1 and noise_mode != "per_unit":
        raise ValueError("sharded runs need noise_mode='per_unit' (the global RNG stream is sequential)")
    B, C, F, H, W = video.shape
    units = enumerate_units(video.shape, chunk_len, overlap_t, tile_size_hw, overlap_hw)
    regions = _unit_regions(units, video.shape, overlap_t, overlap_hw)
    parts = partition_units(units, world)
    mine = parts[rank]

    # Output device / dtype must be the same on EVERY rank, including ranks that own no unit (fewer units than
    # ranks): they are taken from the arguments or from what process_fn declares — never from a local result.
    dev = torch.device(out_device or getattr(process_fn, "device", None) or video.device)
    dt = out_dtype or getattr(process_fn, "out_dtype", None)
    use_cuda_events = timings is not None and dev.type == "cuda"
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if use_cuda_events else None
    if ev:
        ev[0].record()

    results = {}
    for k in mine:
        (t0, t1), (h0, h1, w0, w1) = units[k]
        unit = video[:, :, t0:t1, h0:h1, w0:w1]
        res = process_fn(unit, k, (seed + k) if noise_mode == "per_unit" else None)
        r = regions[k]
        results[k] = res[:, :, r["valid_t_start"]:r["valid_t_end"], r["valid_h_start"]:r["valid_h_end"],
                         r["valid_w_start"]:r["valid_w_end"]]
    if ev:
        ev[1].record()
    if dt is None:
        local = next(iter(results.values())).dtype if results else None
        if world > 1:      # undeclared output dtype: agree on it explicitly (tiny host-side exchange)
            names = [None] * world
            dist.all_gather_object(names, str(local).replace("torch.", "") if local is not None else None, group=group)
            seen = sorted({n for n in names if n is not None})
            if len(seen) > 1:
                raise StitchError(f"ranks disagree on the output dtype: {seen}")
            dt = _DTYPE_NAMES[seen[0]] if seen else video.dtype
        else:
            dt = local if local is not None else video.dtype

    out = torch.zeros((B, C, F, H, W), dtype=dt, device=dev)
    count = torch.zeros((F, H, W), dtype=torch.int32, device=dev)

    def write(k, block):
        r = regions[k]
        sl = (slice(r["out_t_start"], r["out_t_end"]), slice(r["out_h_start"], r["out_h_end"]),
              slice(r["out_w_start"], r["out_w_end"]))
        out[(slice(None), slice(None)) + sl] = block.to(device=dev, dtype=dt)
        count[sl] += 1

    if world == 1:
        for k, block in results.items():
            write(k, block)
    else:
        lens = [sum(_region_numel(regions[k], B * C) for k in p) for p in parts]
        maxlen = max(max(lens), 1)
        send = torch.zeros(maxlen, dtype=dt, device=dev)
        off = 0
        for k in mine:
            n = _region_numel(regions[k], B * C)
            send[off:off + n] = results[k].to(device=dev, dtype=dt).reshape(-1)
            off += n
        recv = torch.empty(world * maxlen, dtype=dt, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)          # the single collective of the path
        for r_ in range(world):
            off = r_ * maxlen
            for k in parts[r_]:
                rg = regions[k]
                shp = (B, C, rg["out_t_end"] - rg["out_t_start"], rg["out_h_end"] - rg["out_h_start"],
                       rg["out_w_end"] - rg["out_w_start"])
                n = _region_numel(rg, B * C)
                write(k, recv[off:off + n].view(shp))
                off += n
    if ev:
        ev[2].record()
    if not bool((count == 1).all()):
        bad = int((count != 1).sum())
        raise StitchError(f"write count != 1 at {bad} positions (overlaps must be even and tiles must cover the clip)")
    if timings is not None:
        timings["units"] = len(mine)
        if ev:
            torch.cuda.synchronize(dev)
            timings["compute_ms"] = ev[0].elapsed_time(ev[1])
            timings["collective_ms"] = ev[1].elapsed_time(ev[2])     # pack + all-gather + unpack/stitch
    return out


def make_process_fn(pipe, empty_prompt_embedding, sr_noise_step=399, noise_step=0, output="unit", use_graph=False):
    """process_fn for `super_resolve` backed by the fused device path of `pipe`.  output="uint8": units come back
    quantised as the reference's savers would (trunc(x*255)), so the gather and the D2H move 1 byte per element.

    use_graph: units of one decomposition share a shape, and a small unit (one of 8 tiles of a 720p clip: ~4 400 launches
    of ~40 us) is bounded by the host's launch rate, so the whole one-step pipeline of a unit shape is captured ONCE into
    a CUDA graph (on the SECOND eager run of that shape) and replayed for every later unit: inputs (the unit's
    pixels, its latent noise — drawn outside the graph from the unit's generator, exactly the eager draw) are copied
    into the graph's static buffers, the result is cloned out.  Bit-identical to the eager path (tests)."""
    graphs, seen = {}, {}

    def fn(unit, k, unit_seed):
        g = None
        if unit_seed is not None:      # every random draw of this unit (latent noise, add_noise) is seeded per unit
            g = torch.Generator(device=pipe.device).manual_seed(int(unit_seed))
        x = _unit_to_device(unit, pipe.device)
        if not use_graph or noise_step != 0:
            return pipe.one_step_sr(x, empty_prompt_embedding, sr_noise_step=sr_noise_step, noise_step=noise_step,
                                    output=output, generator=g)
        _, _, F, H, W = x.shape
        noise = torch.randn((1, 16, pipe.vae.latent_frames(F), H // 8, W // 8), device=pipe.device, dtype=torch.bfloat16,
                            generator=g)                                   # the draw one_step_sr would make itself
        key = (tuple(x.shape), x.dtype)
        ent = graphs.get(key)
        if ent is None:
            out = pipe.one_step_sr(x, empty_prompt_embedding, sr_noise_step=sr_noise_step, noise=noise, output=output)
            seen[key] = seen.get(key, 0) + 1
            if fn.graph_error is None and seen[key] >= 2:      # a shape is captured when it comes back (ragged edge
                try:                                           # units that appear once per clip never pay a capture)
                    sx, sn = x.clone(), noise.clone()
                    torch.cuda.synchronize(pipe.device)
                    graph = torch.cuda.CUDAGraph()
                    # thread_local: CUDA calls of OTHER threads (NCCL watchdog event queries, samplers) must not
                    # invalidate the capture
                    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                        so = pipe.one_step_sr(sx, empty_prompt_embedding, sr_noise_step=sr_noise_step, noise=sn,
                                              output=output)
                    while len(graphs) >= 3:                                # each graph pins a pipeline's worth of
                        graphs.pop(next(iter(graphs)))                     # activations: keep the 3 newest shapes
                    graphs[key] = (graph, sx, sn, so)
                except RuntimeError as ex:     # capture refused (e.g. no memory for the private pool): stay on the
                    fn.graph_error = f"{type(ex).__name__}: {ex}"          # eager GPU path, and say so (fn.graph_error)
                    torch.cuda.synchronize(pipe.device)
            return out
        graph, sx, sn, so = ent
        sx.copy_(x)
        sn.copy_(noise)
        graph.replay()
        return so.clone()
    fn.out_dtype = torch.uint8 if output == "uint8" else torch.bfloat16
    fn.device = pipe.device
    fn.graph_error = None
    fn.uses_graph = lambda: bool(graphs)
    return fn


def preprocess_frames(frames_u8, upscale=4, device="cuda"):
    """GPU version of ref inference_script.py:192-235 + :670-679 (decode excluded): frames [F,H,W,3] uint8 ->
    ([1,3,F',4H',4W'] fp32 in [-1,1] on the device, pad_f, pad_h, pad_w).  Frames are padded to 8k+1 by repeating the
    last frame, H/W to multiples of 16 with zeros (bottom/right), then bilinear x`upscale` on the 0..255 floats and
    x/255*2-1 run in one kernel.  Only the small low-resolution clip crosses PCIe."""
    from . import _lib as L
    from .bookkeeping import frame_padding, spatial_padding
    F, H, W, C = frames_u8.shape
    pad_f, pad_h, pad_w = frame_padding(F), spatial_padding(H), spatial_padding(W)
    x = frames_u8.to(device, non_blocking=True)
    if pad_f:
        x = torch.cat([x, x[-1:].repeat(pad_f, 1, 1, 1)], dim=0)
    if pad_h or pad_w:
        x = torch.nn.functional.pad(x, (0, 0, 0, pad_w, 0, pad_h))
    lr = x.permute(0, 3, 1, 2).float().contiguous()                       # [F,3,h,w] 0..255
    Fp, _, h, w = lr.shape
    out = torch.empty(1, 3, Fp, h * upscale, w * upscale, dtype=torch.float32, device=lr.device)
    L.upscale_normalize(lr, out[0], upscale)
    return out, pad_f, pad_h, pad_w


def remove_padding_and_extra_frames(video, pad_f, pad_h, pad_w, upscale=4):
    """ref inference_script.py:238-246 (called with pad*4 at :731)."""
    if pad_f > 0:
        video = video[:, :, :-pad_f]
    if pad_h > 0:
        video = video[:, :, :, :-pad_h * upscale]
    if pad_w > 0:
        video = video[:, :, :, :, :-pad_w * upscale]
    return video
