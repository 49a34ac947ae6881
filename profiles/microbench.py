"""cfg-5 kernel micro-benchmarks (BASELINE.json configs[4]): DiT attention seq-len sweep and CausalConv3d (T,H,W)
sweep through the C ABI, CUDA-event timed (5 warm-ups, 10 timed, inputs > L2 or L2 flushed by the 256 MB scratch
write between iterations).  Prints achieved TFLOP/s and fraction of the measured cuBLAS bf16 peaks.
    python profiles/microbench.py > profiles/r01_microbench.txt        (on a B200)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dove_b200 import _lib as L   # noqa: E402

PEAKS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else \
    {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
BURST = PEAKS["bf16_tflops"]
flush = None


def timeit(fn, iters=10, warm=5):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in ev:
        flush.zero_()
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    return sum(s.elapsed_time(e) for s, e in ev) / iters


def main():
    L.init(0)
    print(f"# peaks: burst cuBLAS bf16 {BURST} TFLOP/s (kernels timed alone), HBM {PEAKS['hbm_gbs']} GB/s")
    print("## attention (48 heads, d=64): FLOPs = 4*N^2*3072")
    for n in (2048, 4096, 8192, 16384, 19426, 32768, 65536):
        qkv = torch.randn(n, 9216, device="cuda").bfloat16()
        out = torch.empty(n, 3072, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: L.attention(qkv, out, 48, 0.125), iters=5 if n > 20000 else 10)
        tf = 4.0 * n * n * 3072 / ms / 1e9
        print(f"attn N={n:6d}: {ms:9.3f} ms  {tf:8.1f} TFLOP/s  frac_of_burst={tf / BURST:.3f}  "
              f"min-traffic {8 * n * 3072 / 1e6:.0f} MB")
    print("## CausalConv3d 3x3x3 / Conv2d 3x3 (channels-last, bias+residual epilogue): FLOPs = 2*taps*Cin*Cout*T*H*W")
    for (cin, cout, kt, T, H, W) in [(128, 128, 3, 8, 768, 1280), (256, 128, 3, 8, 768, 1280), (256, 256, 3, 8, 384, 640),
                                     (256, 256, 3, 4, 192, 320), (512, 512, 3, 2, 96, 160), (128, 128, 3, 8, 384, 640),
                                     (256, 256, 1, 8, 768, 1280), (128, 128, 3, 2, 240, 360)]:
        x = torch.randn(T + kt - 1, H, W, cin, device="cuda").bfloat16()
        w = (torch.randn(cout, kt * 9 * cin, device="cuda") * (kt * 9 * cin) ** -0.5).bfloat16()
        b = torch.zeros(cout, device="cuda", dtype=torch.bfloat16)
        aux = torch.randn(T, H, W, cout, device="cuda").bfloat16()
        y = torch.empty(T, H, W, cout, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: L.conv_cl(x, w, b, y, T, kt, 3, 3, 1, 1, H, W, cout, epilogue=L.EPI_ADD, aux=aux), iters=5)
        fl = 2.0 * kt * 9 * cin * cout * T * H * W
        by = 2.0 * (cin * (T + kt - 1) + 2 * cout * T) * H * W
        print(f"conv {cin:3d}->{cout:3d} kt{kt} T{T} {H}x{W}: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  "
              f"frac_of_burst={fl / ms / 1e9 / BURST:.3f}  algorithmic {by / 1e9:.2f} GB -> {by / ms / 1e6:.0f} GB/s")
    print("## dense GEMM (DiT linears, M = 19426)")
    for (n, k, epi, name) in [(9216, 3072, 0, "QKV"), (3072, 3072, 2, "out+gate"), (12288, 3072, 1, "FF1+GELU"),
                              (3072, 12288, 2, "FF2+gate")]:
        a = torch.randn(19426, k, device="cuda").bfloat16()
        w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
        bias = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
        aux = torch.randn(19426, n, device="cuda").bfloat16()
        g = torch.randn(n, device="cuda").bfloat16()
        c = torch.empty(19426, n, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: L.gemm(a, w, c, bias, epi, aux if epi == 2 else None, g if epi == 2 else None,
                                   g if epi == 2 else None, 226))
        tf = 2.0 * 19426 * n * k / ms / 1e9
        print(f"gemm {name:9s} N={n:5d} K={k:5d}: {ms:7.3f} ms  {tf:8.1f} TFLOP/s  frac_of_burst={tf / BURST:.3f}")


if __name__ == "__main__":
    main()
