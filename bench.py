#!/usr/bin/env python
"""bench.py — 720p frames/s of DOVE's one-step VSR hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 3 --warmup 3            # this repo (libdove_b200 kernels)
    python bench.py --impl reference --gpus 1 ...            # the reference's arithmetic on the host CPU cores
    python bench.py --impl gpu_library ...                   # the reference's torch-op path on the same B200
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (VAE encode -> 42-layer DiT at t=399 -> VAE decode) over one synthetic
33-frame 768x1280 clip (cfg-2: LR 33x180x320 -> rows padded to 192 -> x4).  N = 1 runs the clip
as ONE unit (untiled, 19 426-token attention); N > 1 shards the same clip into 8 EQUAL script-level spatial tiles
(`--tile_size_hw 416 368 --overlap_hw 64 64`, ref inference_script.py:282-361: 2 x 4 tiles of 416x368, write count 1)
cost-balanced over the ranks, with ONE NCCL all-gather of the valid pixels (uint8, as the reference's savers quantise
them).  `--workload cfg4` runs the 129-frame 1088x1920 long clip (`--chunk_len 25 --overlap_t 12` x `--tile_size_hw 576
528 --overlap_hw 64 64`: 9 chunks x 8 tiles = 72 units of 25x576x528).  Weights: random-init CogVideoX-1.5-5B DiT +
CogVideoX VAE (no network), bf16; prompt = the reference's shipped empty-prompt embedding.

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.  Every step's
working set (>= 2.3 GB activations per frame batch) exceeds the 126 MB L2, so no explicit flush is needed.
`value`: inputs resident in HBM.  `e2e`: pinned host fp32 clip -> public API -> pinned host result, H2D and D2H inside
the timed region (N = 1: `process_video`, bf16 result as the reference returns it; N > 1: `runner.super_resolve`,
each rank copies only its own tiles, uint8 result, D2H on rank 0).
`families`: EVERY C-ABI call inside the timed region is bracketed by CUDA events on the launching stream; the table
(conv / attn / gemm / norm / other / idle) sums to the step.  `roofline`: the conv launch class with the largest summed
time (cfg-2: the 128->128 3x3x3 causal convs at full resolution on `conv_trans_halo_kernel`): algorithmic
FLOPs of one launch / its mean event duration, against MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "cfg2": dict(frames=33, height=768, width=1280, chunk_len=0, overlap_t=8,
                 tiled=dict(tile_size_hw=(416, 368), overlap_hw=(64, 64))),
    "cfg4": dict(frames=129, height=1088, width=1920, chunk_len=25, overlap_t=12,
                 tiled=dict(tile_size_hw=(576, 528), overlap_hw=(64, 64))),
}
CFG1 = dict(frames=8, height=256, width=256)
METRIC = "720p_frames_per_sec_one_step_vsr_33f"
PROMPT_FIXTURE = ROOT / "tests" / "golden" / "empty_prompt_embedding.safetensors"
NCU_TRAFFIC_FILE = ROOT / "profiles" / "r02_ncu_traffic.json"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "gpu_library"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--layers", type=int, default=42, help="DiT depth (42 = CogVideoX-1.5-5B; smaller = debug only)")
    ap.add_argument("--tiled", action="store_true", help="N = 1: run the N > 1 unit decomposition on one GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=30.0, help="budget of timed CPU work in the cpu_baseline leg")
    ap.add_argument("--profile", action="store_true", help="for ncu: honour --warmup < 3, skip e2e / cfg-1 / CPU legs")
    ap.add_argument("--channels-last", action="store_true", help="gpu_library: channels_last_3d memory format")
    ap.add_argument("--debug-small-cpu", action="store_true",
                    help="contract tests only: shrink the CPU arm's clip to 8x32x32 (says so in cpu_baseline.sample)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
def synthetic_clip(F, H, W, seed=0):
    """cfg-2 style input: uint8 noise LR clip -> bilinear x4 on 0..255 floats -> /255*2-1 (ref :670-679)."""
    g = torch.Generator().manual_seed(seed)
    lr = torch.randint(0, 256, (F, 3, H // 4, W // 4), generator=g, dtype=torch.uint8).float()
    up = torch.nn.functional.interpolate(lr, scale_factor=4, mode="bilinear")
    return (up / 255.0 * 2.0 - 1.0).permute(1, 0, 2, 3)[None].contiguous()        # [1,3,F,H,W] fp32


def cfg1_clip():
    """SURVEY 8d / BASELINE cfg-1: torch.manual_seed(0); rand(1,3,8,256,256)*2-1, fp32, straight into process_video."""
    g = torch.Generator().manual_seed(0)
    return torch.rand(1, 3, CFG1["frames"], CFG1["height"], CFG1["width"], generator=g) * 2 - 1


def prompt_embedding():
    from dove_b200.pipeline import load_prompt_embedding, synthetic_prompt_embedding
    if PROMPT_FIXTURE.exists():          # the reference's shipped e3b0...b855.safetensors (ref :580-590)
        return load_prompt_embedding(PROMPT_FIXTURE), "shipped empty-prompt embedding (e3b0...b855.safetensors)"
    return synthetic_prompt_embedding(), "synthetic stand-in with the shipped file's statistics"


def config_dict(args, world):
    """The SAME dict in every arm (ours / reference / gpu_library): it names the workload, not the implementation."""
    w = WORKLOADS[args.workload]
    F, H, W = w["frames"], w["height"], w["width"]
    if args.workload == "cfg2":
        desc = (f"cfg-2: {F}x{H}x{W} one-step VSR (VAE encode -> {args.layers}-layer DiT t=399 -> VAE decode), random-init "
                "CogVideoX-1.5-5B + VAE, untiled VAE")
    else:
        desc = (f"cfg-4: {F}x{H}x{W} long clip, --chunk_len 25 --overlap_t 12 (9 chunks) x --tile_size_hw 576 528 "
                f"--overlap_hw 64 64 (8 tiles) = 72 units, {args.layers}-layer DiT, untiled VAE")
    return {"workload": desc,
            "decomposition": {"1": "one unit (whole clip)" if args.workload == "cfg2" and not args.tiled else "units",
                              "N>1": "8 equal spatial tiles 416x368/64 + 1 all-gather (uint8)"
                              if args.workload == "cfg2" else "72 units over ranks + 1 all-gather (uint8)"},
            "l2": "working set >> L2 (>= 2.3 GB per activation), no flush needed",
            "baseline_note": "vs_baseline = value / 2.21 f/s (published 14.90 s per 33x720x1280 clip, 1xA100)"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.samples, self.reasons, self.proc, self.idx = [], set(), None, gpu_index
        self.max_mhz = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(float(p[1]))
                self.max_mhz = float(p[2])
                for n, v in zip(names, p[4:8]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        p = json.loads(f.read_text())
        return dict(tflops=p["bf16_tflops_sustained"], tflops_burst=p["bf16_tflops"], hbm=p["hbm_gbs"], source="measured")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback")


# ------------------------------------------------------------------------------------------------ per-call timing
FAMILY = {"dove_conv3d_causal_bf16": "conv", "dove_conv_cl_bf16": "conv", "dove_attention_bf16": "attn",
          "dove_gemm_bf16": "gemm", "dove_gemm_qkv_norm_rope_bf16": "gemm", "dove_gn_apply_bf16": "norm", "dove_gn_finalize": "norm", "dove_gn_stats_bf16": "norm",
          "dove_layernorm_mod_bf16": "norm", "dove_qk_norm_rope_bf16": "norm"}


def _v(a):
    return a.value if hasattr(a, "value") else a


class CallTimer:
    """Brackets every C-ABI call with CUDA events on the launching stream (the current torch stream, which is the
    stream handed to the library).  Conv / GEMM / attention calls also get their algorithmic FLOPs from the call's own
    arguments (real, un-padded input channels through the weight-pointer registry)."""

    def __init__(self, L, vae):
        self.L, self.rec = L, []
        self.cin_real = {c.w.data_ptr(): c.cin for c in _all_convs(vae)}

    def __enter__(self):
        self.orig = self.L._call

        Event, orig, rec = torch.cuda.Event, self.orig, self.rec

        def timed(name, *args):        # as little host work as possible inside the timed region: the work model runs later
            s, e = Event(enable_timing=True), Event(enable_timing=True)
            s.record()
            orig(name, *args)
            e.record()
            rec.append((name, s, e, args))
        self.L._call = timed
        return self

    def __exit__(self, *a):
        self.L._call = self.orig

    def _work(self, name, a):
        if name == "dove_conv3d_causal_bf16":
            T, H, W, cin_pad, cout = (_v(a[i]) for i in (5, 6, 7, 8, 10))
            cin = self.cin_real.get(_v(a[2]), cin_pad)
            return 2.0 * 27 * cin * cout * T * H * W, ("conv", cin, cout, 3, T, H, W, 1), \
                2.0 * ((cin_pad * (T + 2) + cout) * H * W) + 2.0 * 27 * cin_pad * cout
        if name == "dove_conv_cl_bf16":
            T, cin_pad, cout, kt, kh, kw, stride, Ho, Wo = (_v(a[i]) for i in (4, 7, 9, 11, 12, 13, 14, 16, 17))
            Hin, Win = _v(a[5]), _v(a[6])
            cin = self.cin_real.get(_v(a[1]), cin_pad)
            return 2.0 * kt * kh * kw * cin * cout * T * Ho * Wo, ("conv", cin, cout, kt, T, Ho, Wo, stride), \
                2.0 * (cin_pad * (T + kt - 1) * Hin * Win + cout * T * Ho * Wo) + 2.0 * kt * kh * kw * cin_pad * cout
        if name == "dove_gemm_bf16":
            M, N, K = (_v(a[i]) for i in (6, 7, 8))
            return 2.0 * M * N * K, ("gemm", M, N, K), 2.0 * (M * K + N * K + M * N)
        if name == "dove_gemm_qkv_norm_rope_bf16":
            M, heads, K = (_v(a[i]) for i in (6, 7, 8))
            N = 3 * heads * 64
            return 2.0 * M * N * K, ("gemm+qk_norm_rope", M, N, K), 2.0 * (M * K + N * K + M * N)
        if name == "dove_attention_bf16":
            rows, heads = _v(a[2]), _v(a[3])
            return 4.0 * rows * rows * heads * 64, ("attn", rows, heads), 2.0 * 4 * rows * heads * 64
        return 0.0, (name,), 0.0

    def result(self, step_ms_total, steps, pk):
        torch.cuda.synchronize()
        fam, classes = {}, {}
        for name, s, e, args in self.rec:
            flops, key, nbytes = self._work(name, args)
            ms = s.elapsed_time(e)
            f = fam.setdefault(FAMILY.get(name, "other"), [0, 0.0, 0.0])
            f[0] += 1
            f[1] += ms
            f[2] += flops
            c = classes.setdefault(key, [0, 0.0, 0.0, 0.0])
            c[0] += 1
            c[1] += ms
            c[2] += flops
            c[3] += nbytes
        busy = sum(v[1] for v in fam.values())
        table = {}
        for k in ("conv", "attn", "gemm", "norm", "other"):
            n, ms, fl = fam.get(k, [0, 0.0, 0.0])
            table[k] = {"launches_per_step": n / steps, "ms_per_step": ms / steps, "share_of_step": ms / step_ms_total}
            if fl > 0 and ms > 0:
                table[k]["tflops"] = fl / (ms / 1e3) / 1e12
                table[k]["frac_of_peak"] = table[k]["tflops"] / pk["tflops"]
        table["idle"] = {"ms_per_step": (step_ms_total - busy) / steps, "share_of_step": (step_ms_total - busy) / step_ms_total}
        table["sum_check"] = sum(v["share_of_step"] for v in table.values())
        return table, classes


def _all_convs(vae):
    out = [vae.enc_conv_in, vae.enc_conv_out, vae.dec_conv_in, vae.dec_conv_out]
    for res, samp, _ in vae.enc_down + vae.dec_up:
        for r in res:
            out += [r.conv1, r.conv2]
        if samp is not None:
            out.append(samp)
    for r in vae.enc_mid + vae.dec_mid:
        out += [r.conv1, r.conv2]
    return out


def conv_kernel_name(cin, cout, kt, Ho, Wo, stride):
    """Which kernel the library dispatches this conv class to (csrc/gemm.cu conv_impl)."""
    if stride == 1 and cout % 256 == 0 and Wo >= 256:
        return "conv2cta_kernel<256>"
    if stride == 1 and cout == 128 and Ho * Wo >= 4096:
        if Wo >= 256:                                              # same rule as conv_impl (gemm.cu)
            return "conv_trans_halo_kernel"
        return "umma_gemm_kernel<256,conv,trans>"
    bn = 256 if cout % 256 == 0 else 128 if cout % 128 == 0 else 64 if cout % 64 == 0 else 32 if cout % 32 == 0 else 16
    return f"umma_gemm_kernel<{bn},conv>"


def roofline(pk, table, classes, ms_total, steps):
    """Dominant kernel = the conv launch class with the largest summed time in the step.  achieved = algorithmic FLOPs
    of ONE such launch / its mean CUDA-event duration.  `traffic`: dram read+write bytes of one launch of the same class
    from the committed ncu --set full capture of this round (profiles/r02_ncu_traffic.json), next to the algorithmic
    bytes (activations in + out once, weights once)."""
    convs = {k: v for k, v in classes.items() if k[0] == "conv"}
    if not convs:
        return None
    key, (n, ms, fl, by) = max(convs.items(), key=lambda kv: kv[1][1])
    _, cin, cout, kt, T, Ho, Wo, stride = key
    ach = fl / (ms / 1e3) / 1e12
    traffic = None
    if NCU_TRAFFIC_FILE.exists():
        traffic = json.loads(NCU_TRAFFIC_FILE.read_text()).get(f"conv {cin}->{cout} kt{kt} T{T} {Ho}x{Wo}")
    return {"bound": "tensor",
            "kernel": f"{conv_kernel_name(cin, cout, kt, Ho, Wo, stride)}: tcgen05 implicit-GEMM conv, class Cin{cin} "
                      f"Cout{cout} kt{kt} T{T} {Ho}x{Wo}",
            "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
            "traffic": traffic, "algorithmic_bytes": by / n,
            "flops_per_launch": fl / n, "ms_per_launch": ms / n, "launches_per_step": n / steps,
            "peak_source": pk["source"] + " (sustained cuBLAS bf16: kernel timed inside a long power-capped step; burst "
                           + str(pk["tflops_burst"]) + ")",
            "class_share_of_step": ms / ms_total,
            "conv_family": {"achieved": table["conv"].get("tflops"), "frac": table["conv"].get("frac_of_peak"),
                            "launches_per_step": table["conv"]["launches_per_step"],
                            "share_of_step": table["conv"]["share_of_step"]}}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def build_oracle_pipe(layers, device, dtype, channels_last=False):
    """The reference's arithmetic (oracle restatement of diffusers, `kind: port`) with the bench's weights (seed 1234)."""
    from dove_b200.weights import dit_param_spec, init_state_dict, vae_param_spec
    from oracle.dit import OracleCogVideoXTransformer3DModel
    from oracle.pipeline import OraclePipe
    from oracle.vae import OracleAutoencoderKLCogVideoX
    gen_dev = "cuda" if torch.cuda.is_available() else "cpu"
    vae = OracleAutoencoderKLCogVideoX().to(dtype)
    dit = OracleCogVideoXTransformer3DModel(num_layers=layers).to(dtype)
    for mod, spec in ((vae, vae_param_spec()), (dit, dit_param_spec(dict(num_layers=layers)))):
        for name, shape, kind in spec:                                  # stream tensor by tensor (23 GB in fp32)
            t = init_state_dict([(name, shape, kind)], 1234, gen_dev, torch.bfloat16)[name]
            owner, attr = name.rsplit(".", 1)
            getattr(mod.get_submodule(owner), attr).data.copy_(t.to(dtype).cpu())
    vae, dit = vae.to(device).eval(), dit.to(device).eval()
    if channels_last:       # only rank-5 tensors have a channels_last_3d format: convert the Conv3d weights one by one
        for m in vae.modules():
            if isinstance(m, torch.nn.Conv3d):
                m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last_3d)
    return OraclePipe(vae, dit)


def cpu_reference_cfg1(seconds_budget, layers, steps=1, warmup=0, debug_small=False):
    """BASELINE cfg-1 exactly: ONE 8-frame 256x256 clip through the fp32 oracle (full `layers`-layer DiT + VAE) on the
    host cores.  Returns (cpu_baseline dict, seconds per clip).  `value` is in the headline metric's unit: the cfg-2 clip
    rate the CPU would reach at the same FLOP rate (cfg-2 / cfg-1 work = 1271.1 / 19.0 TFLOP, dove_b200.workmodel), because
    a 33x768x1280 clip is ~20 minutes of CPU work; the measured cfg-1 clip rate itself is reported next to it."""
    from dove_b200.workmodel import clip_macs
    from oracle.pipeline import oracle_process_video
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    try:                                      # never drive the host out of memory: fp32 weights are 0.53 GB / layer
        import psutil
        avail_gb = psutil.virtual_memory().available / 2 ** 30
        layers_fit = max(1, min(layers, int((avail_gb * 0.6 - 4) / 0.53)))
    except Exception:
        layers_fit = layers
    t0 = time.time()
    pipe = build_oracle_pipe(layers_fit, "cpu", torch.float32)
    load_s = time.time() - t0
    video = cfg1_clip()
    if debug_small:
        video = video[..., :32, :32].contiguous()
    emb, emb_src = prompt_embedding()
    times = []
    for i in range(warmup + steps):
        torch.manual_seed(42)
        t0 = time.time()
        oracle_process_video(pipe, video, emb.float())
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if times and sum(times) + dt > seconds_budget:
            break
    sec = sum(times) / len(times)
    w = WORKLOADS["cfg2"]
    fl_cfg1 = 2.0 * clip_macs(video.shape[2], video.shape[3], video.shape[4], dict(num_layers=layers_fit))["total"]
    fl_cfg2 = 2.0 * clip_macs(w["frames"], w["height"], w["width"], dict(num_layers=layers))["total"]
    cpu_tflops = fl_cfg1 / sec / 1e12
    value = w["frames"] / (fl_cfg2 / (cpu_tflops * 1e12))
    return dict(value=value, unit="frames/s", cores=cores, kind="port",
                sample=("DEBUG 8x32x32 crop (contract test), " if debug_small else "") +
                       f"cfg-1 exactly: one 8x256x256 clip, fp32, full {layers_fit}-layer DiT + VAE (oracle = diffusers "
                       f"restatement), {len(times)} timed run(s) of {sec:.1f} s = {cpu_tflops:.3f} TFLOP/s on {cores} threads; "
                       f"value = cfg-2 clip rate at that FLOP rate ({fl_cfg2 / 1e12:.0f} / {fl_cfg1 / 1e12:.1f} TFLOP); "
                       f"weights load {load_s:.0f} s untimed; prompt = {emb_src}",
                cfg1_frames_per_s=CFG1["frames"] / sec, cfg1_seconds_per_clip=sec, cpu_tflops=cpu_tflops), sec


def run_reference(args, rank):
    if rank != 0:
        return
    # bounded: at most --steps timed runs, stop once ~100 s of timed CPU work has accumulated
    cb, sec = cpu_reference_cfg1(100.0, args.layers, steps=max(1, args.steps), warmup=min(args.warmup, 1),
                                 debug_small=args.debug_small_cpu)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, args.gpus),
            "cpu_baseline": cb,
            "cfg1": {"frames": 8, "height": 256, "width": 256, "cpu_frames_per_s": cb["cfg1_frames_per_s"]},
            "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_gpu_library(args, rank):
    """The reference's GPU path as it runs with stock libraries: the oracle's torch ops in bf16 (cuDNN conv3d, cuBLAS,
    SDPA) on one B200 — the stand-in for `pipe.to("cuda")` + diffusers (absent here) and the denominator of the
    north-star ">= 10x the reference GPU diffusers path" (BASELINE.md section 3)."""
    if rank != 0:
        return
    from oracle.pipeline import oracle_process_video
    w = WORKLOADS[args.workload]
    F, H, W = w["frames"], w["height"], w["width"]
    dev = torch.device("cuda", 0)
    pipe = build_oracle_pipe(args.layers, dev, torch.bfloat16, channels_last=args.channels_last)
    emb, emb_src = prompt_embedding()
    clip = synthetic_clip(F, H, W).to(dev)
    if args.channels_last:
        clip = clip.contiguous(memory_format=torch.channels_last_3d)

    def step():
        torch.manual_seed(42)
        return oracle_process_video(pipe, clip, emb.to(dev))
    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(1, min(args.steps, 3))
    s.record()
    for _ in range(steps):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    line = {"impl": "gpu_library", "metric": METRIC, "value": F / (ms / 1e3), "unit": "frames/s", "n_gpus": 1,
            "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(args, 1),
            "library_path": "oracle torch ops bf16 (cuDNN conv3d / cuBLAS / F.scaled_dot_product_attention), "
                            + ("channels_last_3d" if args.channels_last else "NCDHW") + "; prompt = " + emb_src,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "clocks": sampler.stop(), "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.impl == "gpu_library":
        return run_gpu_library(args, rank)

    import torch.distributed as dist
    from dove_b200 import _lib as L
    from dove_b200.pipeline import CogVideoXPipeline, process_video
    from dove_b200.runner import make_process_fn, super_resolve
    from dove_b200.workmodel import clip_macs

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.init(local_rank)
    if os.environ.get("DOVE_ATTN_VARIANT"):     # kernel A/B runs only (profiles/run_gpu_*.sh); default = automatic choice
        L.set_option("attn_variant", int(os.environ["DOVE_ATTN_VARIANT"]))
    pipe = CogVideoXPipeline.from_random(seed=1234, device=dev, dit_config=dict(num_layers=args.layers))
    emb, emb_src = prompt_embedding()
    wl = WORKLOADS[args.workload]
    F, H, W = wl["frames"], wl["height"], wl["width"]
    units_mode = world > 1 or args.tiled or args.workload != "cfg2"
    host_clip = synthetic_clip(F, H, W).pin_memory()
    dev_clip = host_clip.to(dev)
    out_dtype = torch.uint8 if units_mode else torch.bfloat16
    host_out = torch.empty((1, 3, F, H, W), dtype=out_dtype).pin_memory() if rank == 0 else None
    fn = make_process_fn(pipe, emb, output="uint8")
    # Unit runs (N > 1, --tiled, cfg-4) replay every unit after the first of a shape from a CUDA graph — the product path
    # of the runner for same-shape units.  A replay cannot host per-launch CUDA events, so in unit mode `families` /
    # `roofline` come from an instrumented EAGER pass of the same K steps right after the timed region (`families_source`);
    # at N = 1 untiled (the headline line) every launch of the timed region itself is bracketed.
    use_graph = units_mode and not os.environ.get("DOVE_BENCH_NO_GRAPH") and not args.profile
    fn_graph = make_process_fn(pipe, emb, output="uint8", use_graph=True) if use_graph else fn
    unit_kw = dict(chunk_len=wl["chunk_len"], overlap_t=wl["overlap_t"], noise_mode="per_unit", **wl["tiled"])
    rank_timings = {}

    def step_resident():
        if not units_mode:
            torch.manual_seed(42)
            return pipe.one_step_sr(dev_clip, emb)
        return super_resolve(dev_clip, fn_graph, timings=rank_timings, **unit_kw)

    def step_resident_eager():
        return super_resolve(dev_clip, fn, timings=rank_timings, **unit_kw)

    def step_e2e():
        if not units_mode:
            torch.manual_seed(42)
            out = process_video(pipe, host_clip, empty_prompt_embedding=emb)       # H2D inside (ref :407)
        else:
            out = super_resolve(host_clip, fn_graph, **unit_kw)                     # each rank copies only ITS units
        if rank == 0:
            host_out.copy_(out, non_blocking=True)                                  # D2H inside, on the consumer
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn_, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn_()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    n_warm = args.warmup if args.profile else max(args.warmup, 3)
    for _ in range(n_warm):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = L.launch_count
    pk = peaks()
    families_source = "every launch of the timed region bracketed by CUDA events"
    if use_graph:
        ms_total = timed(step_resident, args.steps)
        rank_timings_timed = dict(rank_timings)        # per-rank times of the timed (graph) pass, not of the eager one
        l0 = L.launch_count
        with CallTimer(L, pipe.vae) as ct:
            ms_inst = timed(step_resident_eager, args.steps)
            table, classes = ct.result(ms_inst, args.steps, pk)
        families_source = (f"separate instrumented eager pass of the same {args.steps} steps ({ms_inst / args.steps:.1f} ms "
                           "per step); the timed region replays CUDA graphs, which cannot host per-launch events")
    else:
        with CallTimer(L, pipe.vae) as ct:
            ms_total = timed(step_resident, args.steps)
            table, classes = ct.result(ms_total, args.steps, pk)
    launches = L.launch_count - l0
    per_rank = None
    if world > 1:             # last timed step's per-rank device times (CUDA events inside super_resolve)
        rt = rank_timings_timed if use_graph else rank_timings
        t = torch.tensor([rt.get("compute_ms", 0.0), rt.get("collective_ms", 0.0), float(rt.get("units", 0))], device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = {"compute_ms": [round(x[0].item(), 2) for x in allt],
                    "collective_ms": [round(x[1].item(), 2) for x in allt], "units": [int(x[2].item()) for x in allt]}
    ms_e2e = timed(step_e2e, args.steps) if not args.profile else ms_total
    clocks = sampler.stop() if rank == 0 else None

    # cfg-1 (BASELINE configs[0]) on the GPU: the SAME input the CPU arm times, through process_video with host buffers
    cfg1 = None
    if rank == 0 and not args.profile and args.workload == "cfg2":
        c1 = cfg1_clip().pin_memory()
        c1_out = torch.empty((1, 3, 8, 256, 256), dtype=torch.bfloat16).pin_memory()

        def step_cfg1():
            torch.manual_seed(42)
            c1_out.copy_(process_video(pipe, c1, empty_prompt_embedding=emb), non_blocking=True)
        for _ in range(3):
            step_cfg1()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            step_cfg1()
        e.record()
        torch.cuda.synchronize()
        ms1 = s.elapsed_time(e) / 5
        cfg1 = {"frames": 8, "height": 256, "width": 256, "gpu_ms_per_clip": ms1, "gpu_frames_per_s": 8 / (ms1 / 1e3),
                "path": "process_video, pinned host fp32 in / bf16 out (e2e)", "same_config_as_cpu_baseline": True}

    if rank == 0:
        ms_step = ms_total / args.steps
        value = F / (ms_step / 1e3)
        e2e_value = F / (ms_e2e / args.steps / 1e3)
        full = args.layers == 42
        flops_clip = 2.0 * clip_macs(F, H, W)["total"] if args.workload == "cfg2" else None
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": value / 2.21 if full and args.workload == "cfg2" else None,
            "dtype": "bf16", "data": "synthetic",
            "config": config_dict(args, world),
            "prompt": emb_src,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s",
                    "h2d_bytes_per_step": host_clip.numel() * 4, "d2h_bytes_per_step": host_out.numel() * host_out.element_size(),
                    "api": "process_video (bf16 result)" if not units_mode else
                           "runner.super_resolve (per-rank unit H2D by strided DMA, units replayed from a CUDA graph: "
                           + ("yes" if fn_graph.uses_graph() else "no" + (f" ({fn_graph.graph_error})" if fn_graph.graph_error else ""))
                           + ", uint8 result, D2H on rank 0)"},
            "gpu_launches": launches,
            "families": table,
            "families_source": families_source,
            "roofline": roofline(pk, table, classes, ms_inst if use_graph else ms_total, args.steps),
        }
        if flops_clip and not units_mode:
            line["step_tflops"] = flops_clip / (ms_step / 1e3) / 1e12
            line["step_frac_of_peak"] = line["step_tflops"] / pk["tflops"]
        if units_mode:          # work actually executed: every unit (chunk x tile) with its overlap, by the work model
            from dove_b200.bookkeeping import enumerate_units
            units = enumerate_units((1, 3, F, H, W), unit_kw["chunk_len"], unit_kw["overlap_t"], unit_kw["tile_size_hw"],
                                    unit_kw["overlap_hw"])
            uf = sum(2.0 * clip_macs(t1 - t0, h1 - h0, w1 - w0)["total"] for (t0, t1), (h0, h1, w0, w1) in units)
            line["units"] = {"count": len(units), "executed_tflop": uf / 1e12,
                             "unit_shape": [units[0][0][1] - units[0][0][0], units[0][1][1] - units[0][1][0],
                                            units[0][1][3] - units[0][1][2]]}
            line["step_tflops"] = uf / (ms_step / 1e3) / 1e12
            line["step_frac_of_peak"] = line["step_tflops"] / (pk["tflops"] * world)
        line["peak_mem_gb_rank0"] = torch.cuda.max_memory_allocated() / 2 ** 30
        top = sorted(((k, v) for k, v in classes.items() if v[1] > 0), key=lambda kv: -kv[1][1])[:8]
        line["top_classes"] = [{"class": " ".join(str(x) for x in k), "launches_per_step": v[0] / args.steps,
                                "ms_per_step": v[1] / args.steps,
                                "tflops": (v[2] / (v[1] / 1e3) / 1e12) if v[2] else None} for k, v in top]
        dump = os.environ.get("DOVE_BENCH_CLASSES")          # full per-class table (every conv / gemm / attn shape) to a file
        if dump:
            rows = [{"class": " ".join(str(x) for x in k), "launches_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps,
                     "tflops": (v[2] / (v[1] / 1e3) / 1e12) if v[2] and v[1] > 0 else None}
                    for k, v in sorted(classes.items(), key=lambda kv: -kv[1][1])]
            Path(dump).write_text(json.dumps(rows, indent=1))
        if per_rank:
            line["ranks"] = per_rank
        if cfg1:
            line["cfg1"] = cfg1
        if not args.no_cpu_baseline and not args.profile and world == 1:
            try:
                cb, _ = cpu_reference_cfg1(args.cpu_seconds, args.layers)
                line["cpu_baseline"] = cb
                if cfg1:
                    cfg1["cpu_frames_per_s"] = cb["cfg1_frames_per_s"]
            except Exception as ex:      # the baseline is reporting only; never lose the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(ex).__name__}: {ex}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
