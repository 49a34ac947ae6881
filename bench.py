#!/usr/bin/env python
"""bench.py — 720p frames/s of DOVE's one-step VSR hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 3 --warmup 3            # this repo (libdove_b200 kernels)
    python bench.py --impl reference --gpus 1 ...            # the reference's arithmetic on the host CPU cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (VAE encode -> 42-layer DiT at t=399 -> VAE decode) over one synthetic
33-frame 768x1280 clip (cfg-2: LR 33x180x320 -> rows padded to 192 -> x4).  N = 1 runs the clip as one
unit; N > 1 shards the same clip into the 8 script-level spatial tiles (`--tile_size_hw 416 352 --overlap_hw
64 64`, ref inference_script.py:282-361) round-robin over the ranks, with one NCCL all-gather of the valid pixels
(strong scaling of one clip).  Weights: random-init CogVideoX-1.5-5B DiT + CogVideoX VAE (no network), bf16.

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.  Every step's
working set (>= 2.3 GB activations per frame batch) exceeds the 126 MB L2, so no explicit flush is needed.
`value`  : inputs resident in HBM.  `e2e`: host pinned fp32 clip -> process_video -> host pinned result, H2D and
D2H inside the timed region.  `roofline`: the dominant kernel (tcgen05 implicit-GEMM conv, 70 % of the FLOPs)
timed per launch with CUDA events inside the timed region: algorithmic FLOPs (2*taps*Cin*Cout*voxels, real
channels) / summed launch time, against MEASURED_PEAKS.json.
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

CFG2 = dict(frames=33, height=768, width=1280)
TILE8 = dict(tile_size_hw=(416, 352), overlap_hw=(64, 64))
METRIC = "720p_frames_per_sec_one_step_vsr_33f"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=42, help="DiT depth (42 = CogVideoX-1.5-5B; smaller = debug only)")
    ap.add_argument("--frames", type=int, default=CFG2["frames"])
    ap.add_argument("--height", type=int, default=CFG2["height"])
    ap.add_argument("--width", type=int, default=CFG2["width"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--profile", action="store_true", help="for ncu: honour --warmup < 3, skip e2e + CPU legs")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
def synthetic_clip(F, H, W, seed=0):
    """cfg-2 style input: uint8 noise LR clip -> bilinear x4 on 0..255 floats -> /255*2-1 (ref :670-679)."""
    g = torch.Generator().manual_seed(seed)
    lr = torch.randint(0, 256, (F, 3, H // 4, W // 4), generator=g, dtype=torch.uint8).float()
    up = torch.nn.functional.interpolate(lr, scale_factor=4, mode="bilinear")
    return (up / 255.0 * 2.0 - 1.0).permute(1, 0, 2, 3)[None].contiguous()        # [1,3,F,H,W] fp32


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


class ConvTimer:
    """CUDA-event timing of every implicit-GEMM conv launch (the dominant kernel family) inside the timed region,
    grouped by problem class (Cin, Cout, kt, T, H, W)."""

    def __init__(self, L):
        self.L, self.ev, self.orig = L, [], L.conv_cl

    def __enter__(self):
        def timed(x, w, bias, y, Tout, kt, kh, kw, stride, pad, Ho, Wo, cout_valid, **kw_):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = self.orig(x, w, bias, y, Tout, kt, kh, kw, stride, pad, Ho, Wo, cout_valid, **kw_)
            e.record()
            cin_real = self.cin_real.get(id(w), x.shape[-1])
            flops = 2.0 * kt * kh * kw * cin_real * cout_valid * Tout * Ho * Wo
            nbytes = 2.0 * (x.shape[-1] * (Tout + kt - 1) * x.shape[1] * x.shape[2] + cout_valid * Tout * Ho * Wo) \
                + 2.0 * kt * kh * kw * x.shape[-1] * cout_valid
            self.ev.append((s, e, flops, nbytes, (cin_real, cout_valid, kt, Tout, Ho, Wo, stride)))
            return r
        self.cin_real = {}
        self.L.conv_cl = timed
        return self

    def register_real_cin(self, vae):
        for c in _all_convs(vae):
            self.cin_real[id(c.w)] = c.cin

    def __exit__(self, *a):
        self.L.conv_cl = self.orig

    def result(self):
        torch.cuda.synchronize()
        tot_f = tot_ms = 0.0
        classes = {}
        for s, e, f, b, key in self.ev:
            ms = s.elapsed_time(e)
            tot_f += f
            tot_ms += ms
            c = classes.setdefault(key, [0, 0.0, 0.0, 0.0])
            c[0] += 1
            c[1] += ms
            c[2] += f
            c[3] += b
        top = max(classes.items(), key=lambda kv: kv[1][1]) if classes else None
        return tot_f, tot_ms, len(self.ev), top


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


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(seconds_budget, layers, steps=1, warmup=0, weights_from=None):
    """Times the reference arithmetic (oracle fp32 restatement of diffusers, `kind: port`) on the host cores on a
    bounded crop of the workload; returns frames/s in 720p-equivalent units (sample pixel-frames / 768x1280)."""
    from dove_b200.pipeline import synthetic_prompt_embedding
    from dove_b200.weights import dit_param_spec, init_state_dict, vae_param_spec
    from oracle.dit import OracleCogVideoXTransformer3DModel
    from oracle.pipeline import OraclePipe, oracle_process_video
    from oracle.vae import OracleAutoencoderKLCogVideoX
    # intra-op threads: the sample's ops are small, so more than 32 threads only adds barrier overhead (measured on the
    # 128-core host: 44-157 s per run with 128 threads); `cores` reports the threads actually used
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    try:                                      # never drive the host out of memory: fp32 weights are 0.53 GB / layer
        import psutil
        avail_gb = psutil.virtual_memory().available / 2 ** 30
        layers = max(1, min(layers, int((avail_gb * 0.5 - 3) / 0.53)))
    except Exception:
        pass
    F, H, W = 9, 64, 96                       # 9-frame 64x96 crop (one VAE frame batch), ~10-20 s on 32 threads
    vae = OracleAutoencoderKLCogVideoX()
    dit = OracleCogVideoXTransformer3DModel(num_layers=layers)
    t0 = time.time()
    if weights_from is not None:
        vsd, dsd = weights_from
        vae.load_state_dict({k: v.float().cpu() for k, v in vsd.items()})
        dit.load_state_dict({k: v.float().cpu() for k, v in dsd.items()})
    else:
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        vae.load_state_dict({k: v.float().cpu() for k, v in
                             init_state_dict(vae_param_spec(), 1234, dev, torch.bfloat16).items()})
        for name, shape, kind in dit_param_spec(dict(num_layers=layers)):     # stream tensor by tensor (23 GB fp32)
            t = init_state_dict([(name, shape, kind)], 1234, dev, torch.bfloat16)[name]
            mod, attr = name.rsplit(".", 1)
            getattr(dit.get_submodule(mod), attr).data.copy_(t.float().cpu())
    load_s = time.time() - t0
    pipe = OraclePipe(vae.eval(), dit.eval())
    video = synthetic_clip(F, H, W, seed=1)
    emb = synthetic_prompt_embedding()
    times = []
    for i in range(warmup + steps):
        torch.manual_seed(42)
        t0 = time.time()
        oracle_process_video(pipe, video, emb.float())
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if sum(times) > seconds_budget and times:
            break
    sec = sum(times) / len(times)
    eq_frames = F * H * W / (CFG2["height"] * CFG2["width"])
    return dict(value=eq_frames / sec, unit="frames/s", cores=cores, kind="port",
                sample=f"{F}x{H}x{W} crop, full {layers}-layer fp32 CPU oracle (diffusers restatement), "
                       f"{len(times)} run(s) of {sec:.1f}s, 720p-equivalent frames = pixel-frames/(768*1280); "
                       f"weights load {load_s:.0f}s untimed"), sec


def run_reference(args, rank):
    if rank != 0:
        return
    # bounded: at most --steps timed runs, stop early once ~90 s of timed CPU work has accumulated
    cb, sec = cpu_reference_sample(90.0, args.layers, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg-2 33x768x1280 one-step VSR; CPU arm timed on a 9x64x96 crop (bounded sample)",
                       "layers": args.layers},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def roofline(pk, conv_flops, conv_ms, conv_launches, top, ms_total, steps):
    """Dominant kernel = the conv launch class with the largest summed time in the step (the 128->128 3x3x3 causal
    convs at full resolution).  achieved = algorithmic FLOPs of ONE such launch / its mean CUDA-event duration.
    `traffic`: dram read+write bytes of the same launch from the committed ncu --set full capture
    (profiles/r01_ncu_conv_trans128.txt), next to the algorithmic bytes (activations in+out once, weights once)."""
    (cin, cout, kt, T, Ho, Wo, stride), (n, ms, fl, by) = top
    ach = fl / (ms / 1e3) / 1e12
    fam = conv_flops / (conv_ms / 1e3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full (2 launches captured: 4.98 / 7.49 GB)
    NCU_TRAFFIC = {(128, 128, 3, 8, 768, 1280): 4.98e9}
    return {"bound": "tensor",
            "kernel": f"tcgen05 implicit-GEMM conv, class Cin{cin} Cout{cout} kt{kt} T{T} {Ho}x{Wo}"
                      + (" [umma_gemm_kernel<256,conv,trans>]" if cout == 128 else " [conv2cta_kernel<256>]"),
            "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
            "traffic": NCU_TRAFFIC.get((cin, cout, kt, T, Ho, Wo)), "algorithmic_bytes": by / n,
            "flops_per_launch": fl / n, "ms_per_launch": ms / n, "launches_per_step": n / steps,
            "peak_source": pk["source"] + " (sustained cuBLAS bf16; burst " + str(pk["tflops_burst"]) + ")",
            "class_share_of_step": ms / ms_total,
            "conv_family": {"achieved": fam, "frac": fam / pk["tflops"], "launches_per_step": conv_launches / steps,
                            "share_of_step": conv_ms / ms_total}}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch.distributed as dist
    from dove_b200 import _lib as L
    from dove_b200.pipeline import CogVideoXPipeline, process_video, synthetic_prompt_embedding
    from dove_b200.runner import make_process_fn, super_resolve

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.init(local_rank)
    pipe = CogVideoXPipeline.from_random(seed=1234, device=dev, dit_config=dict(num_layers=args.layers))
    emb = synthetic_prompt_embedding()
    F, H, W = args.frames, args.height, args.width
    host_clip = synthetic_clip(F, H, W).pin_memory()
    dev_clip = host_clip.to(dev)
    host_out = torch.empty((1, 3, F, H, W), dtype=torch.bfloat16).pin_memory()
    fn = make_process_fn(pipe, emb)
    tile_kw = TILE8 if world > 1 else dict(tile_size_hw=(0, 0), overlap_hw=(32, 32))

    def step_resident():
        if world == 1:
            torch.manual_seed(42)
            return pipe.one_step_sr(dev_clip, emb)
        return super_resolve(dev_clip, fn, **tile_kw)

    def step_e2e():
        if world == 1:
            torch.manual_seed(42)
            out = process_video(pipe, host_clip, empty_prompt_embedding=emb)       # H2D inside
        else:
            out = super_resolve(host_clip.to(dev, non_blocking=True), fn, **tile_kw)
        host_out.copy_(out, non_blocking=True)                                      # D2H inside
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
    with ConvTimer(L) as ct:
        ct.register_real_cin(pipe.vae)
        ms_total = timed(step_resident, args.steps)
        conv_flops, conv_ms, conv_launches, top = ct.result()
    launches = L.launch_count - l0
    ms_e2e = timed(step_e2e, args.steps) if not args.profile else ms_total
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        ms_step = ms_total / args.steps
        value = F / (ms_step / 1e3)
        e2e_value = F / (ms_e2e / args.steps / 1e3)
        pk = peaks()
        ach = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": value / 2.21 if args.layers == 42 and (F, H, W) == (33, 768, 1280) else None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"cfg-2: {F}x{H}x{W} one-step VSR (VAE encode -> {args.layers}-layer DiT t=399 -> VAE "
                                   f"decode), random-init CogVideoX-1.5-5B + VAE, untiled VAE"
                                   + ("" if world == 1 else ", 8 spatial tiles 416x352/64 sharded over ranks + 1 all-gather"),
                       "l2": "working set >> L2 (>= 2.3 GB per activation), no flush needed",
                       "baseline_note": "vs_baseline = value / 2.21 f/s (published 14.90 s per 33x720x1280 clip, 1xA100)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": host_clip.numel() * 4,
                    "d2h_bytes_per_step": host_out.numel() * 2},
            "gpu_launches": launches,
            "roofline": roofline(pk, conv_flops, conv_ms, conv_launches, top, ms_total, args.steps),
        }
        if not args.no_cpu_baseline and not args.profile:
            try:
                cb, _ = cpu_reference_sample(args.cpu_seconds, args.layers)
                line["cpu_baseline"] = cb
            except Exception as ex:      # the baseline is reporting only; never lose the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(ex).__name__}: {ex}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
