"""GroupNorm-apply(+SiLU) micro-benchmark: variants of the kernel's unroll / occupancy (dove_set_option gn_variant)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dove_b200 import _lib as L   # noqa: E402

L.init(0)
T, H, W, C = 8, 768, 1280, 128
x = torch.randn(T, H, W, C, device="cuda").bfloat16()
out = torch.empty_like(x)
gamma = torch.ones(C, device="cuda", dtype=torch.bfloat16)
beta = torch.zeros(C, device="cuda", dtype=torch.bfloat16)
stats = torch.tensor([0.0, 1.0] * 32, device="cuda", dtype=torch.float32)
zy = torch.randn(2, 96, 160, C, device="cuda").bfloat16()
zb = torch.randn(2, 96, 160, C, device="cuda").bfloat16()
partial = torch.empty(L.gn_partial_floats(0), device="cuda", dtype=torch.float32)
for variant in (0,):
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
        print(f"gn_apply variant{variant} spatial={spatial}: {ms:.3f} ms  {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s")
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    L.gn_stats(x, C, 32, 1e-6, partial, stats)
e.record()
torch.cuda.synchronize()
print(f"gn_stats: {s.elapsed_time(e) / 10:.3f} ms")
