"""GroupNorm-apply(+SiLU) / SpatialNorm3D-apply micro-benchmark at the cfg-2 decoder / encoder shapes: CUDA-event time,
effective bandwidth (read x + write out, 4 B per element) against the measured HBM copy peak.
    python profiles/microbench_gn.py > profiles/r02_microbench_gn.txt        (on a B200)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dove_b200 import _lib as L   # noqa: E402

pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
HBM = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
L.init(0)
print(f"# HBM copy peak {HBM} GB/s")
for (T, H, W, C) in [(8, 768, 1280, 128), (8, 384, 640, 256), (4, 192, 320, 512)]:
    x = torch.randn(T, H, W, C, device="cuda").bfloat16()
    out = torch.empty_like(x)
    gamma = torch.ones(C, device="cuda", dtype=torch.bfloat16)
    beta = torch.zeros(C, device="cuda", dtype=torch.bfloat16)
    stats = torch.tensor([0.0, 1.0] * 32, device="cuda", dtype=torch.float32)
    zy = torch.randn(2, 96, 160, C, device="cuda").bfloat16()
    zb = torch.randn(2, 96, 160, C, device="cuda").bfloat16()
    for spatial in (False, True):
        fn = lambda: L.gn_apply(x, out, T, H, W, C, 32, stats, gamma, beta, True, zy if spatial else None,
                                zb if spatial else None)
        for _ in range(3):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        gbs = 2 * x.numel() * 2 / ms / 1e6
        print(f"gn_apply T{T} {H}x{W} C{C} spatial={spatial}: {ms:.3f} ms  {gbs:.0f} GB/s  frac_of_hbm={gbs / HBM:.2f}")
    # accuracy of the one-MUFU SiLU (tanh.approx) against fp32 x*sigmoid(x) on the normalised input
    L.gn_apply(x, out, T, H, W, C, 32, stats, gamma, beta, True, None, None)
    ref = torch.nn.functional.silu(x.float()).bfloat16()
    torch.cuda.synchronize()
    d = (out.float() - ref.float())
    print(f"   silu: rel_l2 vs bf16(fp32 silu) = {(d.norm() / ref.float().norm()).item():.2e}, "
          f"elements differing = {(d != 0).float().mean().item() * 100:.2f} %, max abs = {d.abs().max().item():.3e}")
