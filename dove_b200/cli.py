"""Flag-compatible command line of the reference's `inference_script.py` (ref: /root/reference/inference_script.py:506-778)
on top of the B200 path:

    python -m dove_b200.cli --input_dir datasets/demo --model_path pretrained_models/DOVE --output_path results \
        --is_vae_st --save_format yuv420p [--tile_size_hw 416 368 --overlap_hw 64 64 --chunk_len 25 --overlap_t 12]
    torchrun --nproc-per-node 8 -m dove_b200.cli ...        # units (chunk x tile) sharded over the GPUs of one node

Same flags, defaults and per-video flow as the reference (:507-554, :664-751): read clip -> pad frames to 8k+1 and H/W
to x16 -> bilinear x`--upscale` on 0..255 floats -> /255*2-1 -> chunk x tile units through `process_video` -> stitch with
write-count check -> un-pad -> save.  Differences, all on the B200 side of the boundary: pre-processing runs on the GPU
(`runner.preprocess_frames`), the stitched clip stays in HBM until it is saved, units are sharded over ranks when launched
under torchrun.  Not available in this image (flags are accepted and fail with a clear message when used): `--eval_metrics`
(pyiqa), `--lora_path` (peft), `--is_cpu_offload`, libx264 via imageio (OpenCV's writer is used instead), decord (OpenCV's
reader is used instead).  `--random_init` (extra flag) builds random-init CogVideoX-1.5-5B weights when no checkpoint
directory exists (this container has no network).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
from pathlib import Path

import torch

VIDEO_EXTS = [".mp4", ".avi", ".mov", ".mkv"]                      # ref :40
EMPTY_PROMPT = ("pretrained_models/prompt_embeddings/"
                "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855.safetensors")   # ref :582


def build_parser() -> argparse.ArgumentParser:
    """The reference's 22 flags with the reference's defaults (ref :507-554) + --random_init."""
    p = argparse.ArgumentParser(description="VSR using DOVE (B200-native hot path)")
    p.add_argument("--input_dir", type=str)
    p.add_argument("--input_json", type=str, default=None)
    p.add_argument("--gt_dir", type=str, default=None)
    p.add_argument("--eval_metrics", type=str, default="")
    p.add_argument("--model_path", type=str)
    p.add_argument("--lora_path", type=str, default=None, help="The path of the LoRA weights to be used")
    p.add_argument("--output_path", type=str, default="./results", help="The path save generated video")
    p.add_argument("--fps", type=int, default=16, help="The frames per second for the generated video")
    p.add_argument("--dtype", type=str, default="bfloat16", help="The data type for computation")
    p.add_argument("--seed", type=int, default=42, help="The seed for reproducibility")
    p.add_argument("--upscale_mode", type=str, default="bilinear")
    p.add_argument("--upscale", type=int, default=4)
    p.add_argument("--noise_step", type=int, default=0)
    p.add_argument("--sr_noise_step", type=int, default=399)
    p.add_argument("--is_cpu_offload", action="store_true", help="Enable CPU offload for the model")
    p.add_argument("--is_vae_st", action="store_true", help="Enable VAE slicing and tiling")
    p.add_argument("--png_save", action="store_true", help="Save output as PNG sequence")
    p.add_argument("--save_format", type=str, default="yuv444p", help="Save output as PNG sequence")
    p.add_argument("--tile_size_hw", type=int, nargs=2, default=(0, 0), help="Tile size for spatial tiling (height, width)")
    p.add_argument("--overlap_hw", type=int, nargs=2, default=(32, 32))
    p.add_argument("--chunk_len", type=int, default=0, help="Chunk length for temporal chunking")
    p.add_argument("--overlap_t", type=int, default=8)
    p.add_argument("--random_init", action="store_true", help="random-init CogVideoX-1.5-5B weights (no checkpoint)")
    return p


def effective_overlaps(args):
    """ref :565-576: overlaps only apply when the corresponding chunking / tiling is on."""
    overlap_t = args.overlap_t if args.chunk_len > 0 else 0
    overlap_hw = tuple(args.overlap_hw) if tuple(args.tile_size_hw) != (0, 0) else (0, 0)
    return overlap_t, overlap_hw


def read_video_frames(path) -> torch.Tensor:
    """[F, H, W, 3] uint8 RGB (the reference uses decord, ref :211-214; OpenCV is what this image has)."""
    import cv2
    cap = cv2.VideoCapture(str(path))
    frames = []
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        frames.append(torch.from_numpy(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)))
    cap.release()
    if not frames:
        raise ValueError(f"could not decode any frame from {path}")
    return torch.stack(frames)


def save_frames_u8(frames_u8, output_path, fps, png):
    """frames_u8 [F, H, W, 3] uint8 RGB on the host (already quantised as ref :124/:143/:168 do)."""
    import cv2
    if png:
        out_dir = str(output_path).rsplit(".", 1)[0]                                   # ref :747
        os.makedirs(out_dir, exist_ok=True)
        for i, f in enumerate(frames_u8.numpy()):
            cv2.imwrite(os.path.join(out_dir, f"{i:03d}.png"), cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
        return out_dir
    output_path = str(output_path).replace(".mkv", ".mp4")                             # ref :750
    F, H, W, _ = frames_u8.shape
    vw = cv2.VideoWriter(output_path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (W, H))
    for f in frames_u8.numpy():
        vw.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    vw.release()
    return output_path


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.dtype != "bfloat16":
        raise NotImplementedError("dove_b200 computes in bfloat16 (the reference default, ref :523)")
    if args.upscale_mode != "bilinear":
        raise NotImplementedError("only --upscale_mode bilinear (the reference default) runs on the GPU pre-processing")
    if args.eval_metrics:
        raise NotImplementedError("--eval_metrics needs pyiqa and its pretrained metric networks (not in this image)")
    if args.is_cpu_offload:
        raise NotImplementedError("--is_cpu_offload: 180 GB of HBM per B200 holds the whole model; offload is not implemented")
    import torch.distributed as dist
    from .pipeline import CogVideoXPipeline, load_prompt_embedding
    from .runner import (make_process_fn, preprocess_frames, remove_padding_and_extra_frames, super_resolve)
    from .scheduler import CogVideoXDPMScheduler
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    overlap_t, overlap_hw = effective_overlaps(args)
    torch.manual_seed(args.seed)                                                        # set_seed, ref :577
    torch.cuda.manual_seed_all(args.seed)
    if not Path(EMPTY_PROMPT).exists():
        raise FileNotFoundError(f"{EMPTY_PROMPT} not found (ref :580-590); DOVE runs with the pre-computed empty prompt")
    emb = load_prompt_embedding(EMPTY_PROMPT)
    video_prompt = json.load(open(args.input_json)) if args.input_json else {}
    files = sorted(f for ext in VIDEO_EXTS for f in glob.glob(os.path.join(args.input_dir, f"*{ext}")))
    if not files:
        raise ValueError(f"No video files found in {args.input_dir}")
    os.makedirs(args.output_path, exist_ok=True)
    if args.random_init:
        pipe = CogVideoXPipeline.from_random(device=dev)
    else:
        pipe = CogVideoXPipeline.from_pretrained(args.model_path, torch_dtype=torch.bfloat16, device=dev)   # ref :613
    if args.lora_path:
        pipe.load_lora_weights(args.lora_path, weight_name="pytorch_lora_weights.safetensors", adapter_name="test_1")
        pipe.fuse_lora(components=["transformer"], lora_scale=1.0)
    pipe.scheduler = CogVideoXDPMScheduler.from_config(pipe.scheduler.config, timestep_spacing="trailing")   # ref :629
    pipe.to("cuda")
    if args.is_vae_st:                                                                  # ref :643-645
        pipe.vae.enable_slicing()
        pipe.vae.enable_tiling()
    # several units per clip (chunks / tiles) share a shape: replay them from one CUDA graph (runner.make_process_fn)
    multi_unit = args.chunk_len > 0 or tuple(args.tile_size_hw) != (0, 0)
    fn = make_process_fn(pipe, emb, sr_noise_step=args.sr_noise_step, noise_step=args.noise_step, output="uint8",
                         use_graph=multi_unit)
    for path in files:
        name = os.path.basename(path)
        if video_prompt.get(name, "") != "":
            raise NotImplementedError("non-empty prompts need the T5 encoder (DOVE always uses \"\", README.md:235)")
        frames = read_video_frames(path)
        video, pad_f, pad_h, pad_w = preprocess_frames(frames, args.upscale, dev)      # ref :670-679 on the GPU
        out = super_resolve(video, fn, chunk_len=args.chunk_len, overlap_t=overlap_t, tile_size_hw=tuple(args.tile_size_hw),
                            overlap_hw=overlap_hw, noise_mode="global" if world == 1 else "per_unit", seed=args.seed)
        out = remove_padding_and_extra_frames(out, pad_f, pad_h, pad_w, 4)              # x4 hard-coded, ref :731
        if rank == 0:
            u8 = out[0].permute(1, 2, 3, 0).contiguous().cpu()                          # [F, H, W, 3] uint8
            where = save_frames_u8(u8, os.path.join(args.output_path, name), args.fps, args.png_save)
            print(f"Process video: {name} | Frame: {video.shape[2]} (ori: {frames.shape[0]}; pad: {pad_f}) | Target "
                  f"Resolution: {video.shape[3]}, {video.shape[4]} | saved {where}")
    if rank == 0:
        print("All videos processed.")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
