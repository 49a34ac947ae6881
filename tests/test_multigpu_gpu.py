"""NCCL N-rank run == 1-rank run (needs >= 2 GPUs; `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`).

Two ranks, one process per GPU, `runner.super_resolve` over the real device path with per-unit seeds and ONE
all_gather_into_tensor over NCCL: the stitched clip must equal, bit for bit and on every rank, the clip a single process
stitches from the same units — including a clip with FEWER units than ranks (a rank without work must still enter the
collective with the right device and dtype) and the uint8 output mode."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [dict(shape=(1, 3, 9, 96, 96), chunk_len=0, overlap_t=8, tile_size_hw=(64, 64), overlap_hw=(32, 32), output="unit"),
         dict(shape=(1, 3, 9, 32, 48), chunk_len=0, overlap_t=8, tile_size_hw=(0, 0), overlap_hw=(32, 32), output="unit"),
         # temporal chunks (0,17) + (9,33) (the short tail is merged, ref :274-277), uint8 output
         dict(shape=(1, 3, 33, 32, 48), chunk_len=17, overlap_t=8, tile_size_hw=(0, 0), overlap_hw=(32, 32), output="uint8")]


def _pipe(device):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import models
    vsd, dsd = models.state_dicts(models.SMALL_DIT)
    return models.b200_pipe(vsd, dsd, models.SMALL_DIT, device), models.prompt_embedding()


def _run_cases(pipe, emb, on_host):
    from dove_b200.runner import make_process_fn, super_resolve
    outs = []
    for c in CASES:
        torch.manual_seed(0)
        v = torch.rand(c["shape"]) * 2 - 1
        if not on_host:
            v = v.to(pipe.device)
        kw = {k: c[k] for k in ("chunk_len", "overlap_t", "tile_size_hw", "overlap_hw")}
        t = {}
        outs.append(super_resolve(v, make_process_fn(pipe, emb, output=c["output"]), noise_mode="per_unit", seed=42,
                                  timings=t, **kw).cpu())
        assert "units" in t
    return outs


def _worker(rank, world, port, outdir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pipe, emb = _pipe(dev)
    outs = _run_cases(pipe, emb, on_host=(rank == 1))      # rank 1 feeds its units from (pinned-staged) host memory
    torch.save(outs, os.path.join(outdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    pipe, emb = _pipe(torch.device("cuda", 0))
    single = _run_cases(pipe, emb, on_host=False)
    del pipe
    torch.cuda.empty_cache()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    for r in range(2):
        got = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        for i, c in enumerate(CASES):
            assert got[i].dtype == single[i].dtype == (torch.uint8 if c["output"] == "uint8" else torch.bfloat16)
            assert torch.equal(got[i], single[i]), (r, i)
