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


CASES = [dict(shape=(1, 3, 33, 96, 160), chunk_len=0, overlap_t=8, tile_size_hw=(52, 44), overlap_hw=(8, 8)),
         dict(shape=(1, 3, 57, 64, 64), chunk_len=25, overlap_t=8, tile_size_hw=(48, 48), overlap_hw=(16, 16)),
         dict(shape=(1, 3, 9, 32, 32), chunk_len=0, overlap_t=8, tile_size_hw=(0, 0), overlap_hw=(32, 32))]


def _worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    outs = []
    for c in CASES:
        torch.manual_seed(0)
        v = torch.rand(c["shape"]) * 2 - 1
        kw = {k: c[k] for k in ("chunk_len", "overlap_t", "tile_size_hw", "overlap_hw")}
        outs.append(super_resolve(v, _unit_fn, **kw))
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
        single = super_resolve(v, _unit_fn, noise_mode="per_unit", **kw)
        for r in range(2):
            assert torch.equal(got[r][i], single), (i, r)


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
