"""Oracle pipe + the one-step SR step.  TEST INFRASTRUCTURE (see oracle/__init__.py).

`OraclePipe` exposes the attribute surface that /root/reference/inference_script.py:394-503
(`process_video`) touches on a diffusers `CogVideoXPipeline`, so the
reference's own `process_video` source runs on it unmodified (tests/test_reference_source.py).
`oracle_process_video` restates that function step by step with the file:line of each step.
`decode_latents` restates diffusers `pipelines/cogvideo/pipeline_cogvideox.py::decode_latents`.
"""
from __future__ import annotations

import torch

from .dit import OracleCogVideoXTransformer3DModel, get_3d_rotary_pos_embed
from .scheduler import OracleCogVideoXDPMScheduler
from .vae import OracleAutoencoderKLCogVideoX


class OraclePipe:
    def __init__(self, vae: OracleAutoencoderKLCogVideoX, transformer: OracleCogVideoXTransformer3DModel,
                 scheduler: OracleCogVideoXDPMScheduler | None = None):
        self.vae = vae
        self.transformer = transformer
        self.scheduler = scheduler or OracleCogVideoXDPMScheduler()
        self.vae_scaling_factor_image = vae.config.scaling_factor

    @property
    def device(self):
        return self.vae.device

    def to(self, device):
        self.vae.to(device)
        self.transformer.to(device)
        return self

    def decode_latents(self, latents):
        latents = latents.permute(0, 2, 1, 3, 4)
        latents = 1 / self.vae_scaling_factor_image * latents
        return self.vae.decode(latents).sample


def prepare_rope(height, width, num_frames, tcfg, vae_sf, device):
    """ref: inference_script.py:364-392"""
    gh = height // (vae_sf * tcfg.patch_size)
    gw = width // (vae_sf * tcfg.patch_size)
    base = num_frames if tcfg.patch_size_t is None else (num_frames + tcfg.patch_size_t - 1) // tcfg.patch_size_t
    return get_3d_rotary_pos_embed(tcfg.attention_head_dim, None, (gh, gw), base, grid_type="slice",
                                   max_size=(gh, gw), device=device)


@torch.no_grad()
def oracle_process_video(pipe, video, empty_prompt_embedding, noise_step=0, sr_noise_step=399, noise=None,
                         return_intermediates=False):
    """Restatement of ref: inference_script.py:394-503 for prompt == "" (the only mode DOVE uses).
    `noise`: optional explicit latent noise (tests inject it so GPU and CPU paths share the draw)."""
    inter = {}
    video = video.to(pipe.vae.device, dtype=pipe.vae.dtype)                          # :407
    dist = pipe.vae.encode(video).latent_dist                                        # :408
    inter["moments"] = dist.parameters
    latent = dist.sample(noise=noise) * pipe.vae.config.scaling_factor               # :409
    pt = pipe.transformer.config.patch_size_t
    ncopy = 0
    if pt is not None:                                                               # :411-418
        ncopy = latent.shape[2] % pt
        latent = torch.cat([latent[:, :, :1].repeat(1, 1, ncopy, 1, 1), latent], dim=2)
        assert latent.shape[2] % pt == 0
    b, c, f, h, w = latent.shape
    emb = empty_prompt_embedding.to(latent.device, dtype=latent.dtype)               # :423-428
    if emb.shape[0] != b:
        emb = emb.repeat(b, 1, 1)
    latent = latent.permute(0, 2, 1, 3, 4)                                           # :446
    if noise_step != 0:                                                              # :449-457
        n = torch.randn_like(latent)
        latent = pipe.scheduler.add_noise(latent, n, torch.full((b,), noise_step, dtype=torch.long,
                                                                device=latent.device))
    t = torch.full((b,), sr_noise_step, dtype=torch.long, device=latent.device)      # :459-464
    sf = 2 ** (len(pipe.vae.config.block_out_channels) - 1)                          # :467
    rope = prepare_rope(h * sf, w * sf, f, pipe.transformer.config, sf, latent.device)   # :468-480
    inter["latent"] = latent
    pred = pipe.transformer(hidden_states=latent, encoder_hidden_states=emb, timestep=t,
                            image_rotary_emb=rope, return_dict=False)[0]             # :483-489
    inter["pred"] = pred
    x0 = pipe.scheduler.get_velocity(pred, latent, t)                                # :491-493
    if pt is not None and ncopy > 0:                                                 # :496-497
        x0 = x0[:, ncopy:]
    inter["x0"] = x0
    out = pipe.decode_latents(x0)                                                    # :500
    inter["decoded"] = out
    out = (out * 0.5 + 0.5).clamp(0.0, 1.0)                                          # :501
    return (out, inter) if return_intermediates else out
