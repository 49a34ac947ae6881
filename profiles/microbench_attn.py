"""Attention kernel variants at the cfg-2 shape (N = 19 426 tokens, 48 heads, d = 64) and the tile shape of the 8-GPU
decomposition (N = 2 618): CUDA-event time with the L2 flushed between iterations, TFLOP/s (4 N^2 heads d), fraction of
the measured cuBLAS bf16 burst peak, error against fp32 softmax attention on the first 512 query rows.
    python profiles/microbench_attn.py > profiles/r02_microbench_attn.txt        (on a B200)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dove_b200 import _lib as L   # noqa: E402

pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
BURST = json.load(open(pk))["bf16_tflops"] if os.path.exists(pk) else 1590.0
NAMES = {0: "v2 (1 Q tile/CTA, 2 CTAs/SM)", 1: "v3 emu 0/8", 2: "v3 emu 1/8", 3: "v3 emu 2/8", 4: "v3 emu 3/8",
         5: "v3 emu 4/8", 6: "v3 emu 5/8", 7: "v4 emu 0/8", 8: "v4 emu 1/8", 9: "v4 emu 2/8", 10: "v4 emu 3/8",
         11: "v4 emu 4/8"}


def timeit(fn, flush, iters=8, warm=3):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in ev:
        flush.zero_()
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    t = sorted(s.elapsed_time(e) for s, e in ev)
    return t[len(t) // 2]


VARIANTS = [int(v) for v in sys.argv[1:]] or list(range(12))
NS = [int(v) for v in os.environ.get("DOVE_ATTN_NS", "19426,2618,32866").split(",")]   # 3216 / 4978 = cfg-3 / cfg-4 units


def main():
    L.init(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    heads = 48
    for n in NS:
        g = torch.Generator(device="cuda").manual_seed(n)
        qkv = torch.randn(n, 3 * heads * 64, generator=g, device="cuda").bfloat16()
        q, k, v = [t.float().reshape(n, heads, 64).transpose(0, 1) for t in qkv.chunk(3, dim=1)]
        p = torch.softmax(q[:, :512] @ k.transpose(1, 2) * 0.125, dim=-1)
        ref = (p @ v).transpose(0, 1).reshape(512, heads * 64)
        del q, k, v, p
        print(f"## N = {n}, {heads} heads (burst peak {BURST} TFLOP/s)")
        for var in VARIANTS:
            L.set_option("attn_variant", var)
            out = torch.zeros(n, heads * 64, device="cuda", dtype=torch.bfloat16)
            ms = timeit(lambda: L.attention(qkv, out, heads, 0.125), flush)
            err = ((out[:512].float() - ref).norm() / ref.norm()).item()
            tf = 4.0 * n * n * heads * 64 / ms / 1e9
            print(f"variant {var} {NAMES[var]:30s}: {ms:8.3f} ms  {tf:7.1f} TFLOP/s  frac_of_burst={tf / BURST:.3f}  "
                  f"rel_l2 vs fp32 = {err:.2e}", flush=True)
    L.set_option("attn_variant", L.DEFAULT_ATTN_VARIANT)


if __name__ == "__main__":
    main()
