"""Host logic of the multi-GPU path on CPU: world_size-2 `gloo` run of runner.super_resolve (static unit
partition + ONE all_gather) must equal the single-process run bit for bit, on every rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dove_b200.bookkeeping import enumerate_units, partition_units
from dove_b200.runner import super_resolve


def _unit_fn(unit, k, seed):
    # deterministic stand-in for the one-step SR of a unit; uses the per-unit seed like the real path
    g = torch.Generator().manual_seed(int(seed))
    noise = torch.randn(unit.shape, generator=g)
    return (torch.tanh(unit + 0.1 * noise) * 0.5 + 0.5)


def _unit_fn_bf16(unit, k, seed):
    return _unit_fn(unit, k, seed).bfloat16()


CASES = [dict(shape=(1, 3, 33, 96, 160), chunk_len=0, overlap_t=8, tile_size_hw=(52, 44), overlap_hw=(8, 8)),
         dict(shape=(1, 3, 57, 64, 64), chunk_len=25, overlap_t=8, tile_size_hw=(48, 48), overlap_hw=(16, 16)),
         dict(shape=(1, 3, 9, 32, 32), chunk_len=0, overlap_t=8, tile_size_hw=(0, 0), overlap_hw=(32, 32)),
         # ONE unit on two ranks, bf16 results: the rank without units must agree on the gather buffer's dtype
         dict(shape=(1, 3, 9, 32, 32), chunk_len=0, overlap_t=8, tile_size_hw=(0, 0), overlap_hw=(32, 32), fn="bf16")]


def _worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    outs = []
    for c in CASES:
        torch.manual_seed(0)
        v = torch.rand(c["shape"]) * 2 - 1
        kw = {k: c[k] for k in ("chunk_len", "overlap_t", "tile_size_hw", "overlap_hw")}
        outs.append(super_resolve(v, _unit_fn_bf16 if c.get("fn") == "bf16" else _unit_fn, **kw))
    torch.save(outs, os.path.join(outdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_equals_single_process(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    outdir = str(tmp_path)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, outdir)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    got = {r: torch.load(os.path.join(outdir, f"rank{r}.pt")) for r in range(2)}
    for i, c in enumerate(CASES):
        torch.manual_seed(0)
        v = torch.rand(c["shape"]) * 2 - 1
        kw = {k: c[k] for k in ("chunk_len", "overlap_t", "tile_size_hw", "overlap_hw")}
        single = super_resolve(v, _unit_fn_bf16 if c.get("fn") == "bf16" else _unit_fn, noise_mode="per_unit", **kw)
        for r in range(2):
            assert got[r][i].dtype == single.dtype and torch.equal(got[r][i], single), (i, r)


def test_equal_tile_decomposition_of_cfg3():
    """bench.py's N > 1 decomposition: `--tile_size_hw 416 368 --overlap_hw 64 64` on 768x1280 -> 8 IDENTICAL tiles
    (2 x 4 of 416x368), write count exactly 1, every side a multiple of 16, one unit per rank at N = 8."""
    from dove_b200.bookkeeping import get_valid_tile_region, unit_cost
    shape = (1, 3, 33, 768, 1280)
    units = enumerate_units(shape, 0, 8, (416, 368), (64, 64))
    assert len(units) == 8
    assert {(h1 - h0, w1 - w0) for _, (h0, h1, w0, w1) in units} == {(416, 368)}
    assert len({unit_cost(u) for u in units}) == 1
    cnt = torch.zeros(768, 1280, dtype=torch.int32)
    for (t0, t1), (h0, h1, w0, w1) in units:
        g = get_valid_tile_region(t0, t1, h0, h1, w0, w1, shape, 8, 64, 64)
        cnt[g["out_h_start"]:g["out_h_end"], g["out_w_start"]:g["out_w_end"]] += 1
    assert bool((cnt == 1).all())
    assert [len(p) for p in partition_units(units, 8)] == [1] * 8
    # cfg-4: 9 chunks of exactly 25 frames x 8 equal 576x528 tiles = 72 units, 9 per rank
    shape4 = (1, 3, 129, 1088, 1920)
    u4 = enumerate_units(shape4, 25, 12, (576, 528), (64, 64))
    assert len(u4) == 72 and {(t1 - t0, h1 - h0, w1 - w0) for (t0, t1), (h0, h1, w0, w1) in u4} == {(25, 576, 528)}
    assert [len(p) for p in partition_units(u4, 8)] == [9] * 8
    c4 = torch.zeros(129, 1088 // 16, 1920 // 16, dtype=torch.int32)
    for (t0, t1), (h0, h1, w0, w1) in u4:
        g = get_valid_tile_region(t0, t1, h0, h1, w0, w1, shape4, 12, 64, 64)
        assert all(g[k] % 16 == 0 for k in ("out_h_start", "out_h_end", "out_w_start", "out_w_end"))
        c4[g["out_t_start"]:g["out_t_end"], g["out_h_start"] // 16:g["out_h_end"] // 16,
           g["out_w_start"] // 16:g["out_w_end"] // 16] += 1
    assert bool((c4 == 1).all())


def test_partition_covers_all_units_once():
    shape = (1, 3, 33, 768, 1280)
    units = enumerate_units(shape, 0, 8, (416, 352), (64, 64))
    assert len(units) == 8                                   # exactly 8 spatial tiles for cfg-3
    from dove_b200.bookkeeping import unit_cost
    for world in (1, 2, 4, 8, 3):
        for balance in (True, False):
            parts = partition_units(units, world, balance=balance)
            assert sorted(k for p in parts for k in p) == list(range(len(units)))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    # the two 416-wide tiles (indices 3 and 7) must not land on the same rank when balancing over 2 or 4 ranks
    for world in (2, 4):
        loads = [sum(unit_cost(units[k]) for k in p) for p in partition_units(units, world)]
        rr = [sum(unit_cost(units[k]) for k in p) for p in partition_units(units, world, balance=False)]
        assert max(loads) < max(rr)
        assert all(not ({3, 7} <= set(p)) for p in partition_units(units, world))
