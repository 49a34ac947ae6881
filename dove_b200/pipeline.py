"""Drop-in pipeline object for DOVE's one-step inference hot path.

`CogVideoXPipeline` mirrors the attribute surface of diffusers' `CogVideoXPipeline` that
/root/reference/inference_script.py touches:

    from_pretrained(path, torch_dtype)            ref :613      .to("cuda") / .device               :641, :656
    .vae .transformer .scheduler (assignable)     :629-631      .decode_latents(latents)            :500
    .enable_sequential_cpu_offload / load_lora_weights / fuse_lora  (flags default off -> NotImplementedError)
    .tokenizer / .text_encoder                    :431-444      (non-empty prompts only -> NotImplementedError)

`process_video` is this package's implementation of ref :394-503 (same signature and semantics); it runs the
fused device path (`one_step_sr`): pixels -> encoder -> sample*0.7 -> first-frame copy -> DiT -> x0 ->
decoder -> *0.5+0.5 clamp, entirely through libdove_b200 kernels.  The reference's own `process_video` source
is run unmodified against this object on the GPU by tests/test_pipeline_gpu.py (fixture tests/golden/reference_process_video.py.txt,
extracted by tests/golden/make_reference_source_fixture.py) and against the CPU oracle pipe by tests/test_reference_source.py.
"""
from __future__ import annotations

import json
from pathlib import Path

import torch

from . import _lib as L
from .embeddings import get_3d_rotary_pos_embed
from .scheduler import CogVideoXDPMScheduler
from .transformer import CogVideoXTransformer3DModel
from .vae import AutoencoderKLCogVideoX
from .weights import (DIT_CONFIG, SCHED_CONFIG, VAE_CONFIG, dit_param_spec, init_state_dict, load_safetensors_dir,
                      vae_param_spec)

BF = torch.bfloat16


class CogVideoXPipeline:
    def __init__(self, vae: AutoencoderKLCogVideoX, transformer: CogVideoXTransformer3DModel,
                 scheduler: CogVideoXDPMScheduler | None = None):
        self.vae = vae
        self.transformer = transformer
        self.scheduler = scheduler or CogVideoXDPMScheduler()
        self.vae_scaling_factor_image = vae.config.scaling_factor
        self.vae_scale_factor_spatial = 2 ** (len(vae.config.block_out_channels) - 1)

    # ---- construction -------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, path, torch_dtype=torch.bfloat16, device="cuda"):
        """Read a diffusers-layout directory (`transformer/`, `vae/`, `scheduler/` with config.json +
        safetensors).  Only bf16 compute is implemented (BASELINE config)."""
        if torch_dtype not in (torch.bfloat16, None):
            raise NotImplementedError("dove_b200 computes in bfloat16 (pass --dtype bfloat16)")
        path = Path(path)

        def cfg(sub, name="config.json"):
            f = path / sub / name
            return {k: v for k, v in json.loads(f.read_text()).items() if not k.startswith("_")} if f.exists() else {}

        known_v, known_d = set(VAE_CONFIG), set(DIT_CONFIG)
        vcfg = {k: v for k, v in cfg("vae").items() if k in known_v}
        dcfg = {k: v for k, v in cfg("transformer").items() if k in known_d}
        scfg = {k: v for k, v in cfg("scheduler", "scheduler_config.json").items() if k in SCHED_CONFIG}
        vae = AutoencoderKLCogVideoX(load_safetensors_dir(path / "vae"), vcfg, device)
        dit = CogVideoXTransformer3DModel(load_safetensors_dir(path / "transformer"), dcfg, device)
        return cls(vae, dit, CogVideoXDPMScheduler(**scfg))

    @classmethod
    def from_random(cls, seed=1234, device="cuda", vae_config=None, dit_config=None, init_device=None):
        """Random-init weights of the CogVideoX-1.5-5B architecture (no network / released weights here)."""
        init_device = init_device or device
        vsd = init_state_dict(vae_param_spec(vae_config), seed, init_device, BF)
        vae = AutoencoderKLCogVideoX(vsd, vae_config, device)
        del vsd
        dsd = init_state_dict(dit_param_spec(dit_config), seed, init_device, BF)
        dit = CogVideoXTransformer3DModel(dsd, dit_config, device)
        del dsd
        return cls(vae, dit)

    # ---- surface ----------------------------------------------------------------------------------
    @property
    def device(self):
        return self.vae.device

    def to(self, device=None, *a, **k):
        if device is not None and torch.device(device).type != "cuda":
            raise L.DoveError("dove_b200 has no CPU path: weights live on one sm_100a GPU")
        return self

    def enable_sequential_cpu_offload(self, *a, **k):
        raise NotImplementedError("CPU offload is pointless with 180 GB of HBM per B200 and is not implemented")

    def load_lora_weights(self, *a, **k):
        raise NotImplementedError("LoRA loading/fusing is out of scope (SURVEY.md section 8 f-3)")

    fuse_lora = load_lora_weights

    def tokenizer(self, *a, **k):
        raise NotImplementedError("DOVE always uses the empty prompt (pre-computed embedding, ref :580-590)")

    text_encoder = tokenizer

    def decode_latents(self, latents):
        """[B,F,16,h,w] -> [B,3,F',8h,8w]: permute, * 1/scaling_factor, vae.decode (diffusers decode_latents)."""
        z = latents.permute(0, 2, 1, 3, 4).contiguous()
        return self.vae.decode_scaled(z, 1.0 / self.vae_scaling_factor_image)

    def _prompt_on_device(self, emb):
        """[n_text, 4096] bf16 on the GPU.  The embedding is the same constant tensor for every unit of every clip
        (ref :580-590), so its device copy is kept — which also lets the transformer keep `text_proj(emb)` cached."""
        key = (emb.data_ptr(), emb._version, tuple(emb.shape), emb.dtype, str(emb.device))
        hit = getattr(self, "_prompt_cache", None)
        if hit is None or hit[0] != key:
            d = emb.to(self.device, dtype=BF)
            d = d.reshape(-1, d.shape[-1]).contiguous()
            self._prompt_cache = hit = (key, emb, d)       # holding `emb` keeps the pointer key valid
        return hit[2]

    # ---- fused one-step path ------------------------------------------------------------------------------
    @torch.no_grad()
    def one_step_sr(self, video, empty_prompt_embedding, sr_noise_step=399, noise=None, noise_step=0,
                    return_intermediates=False, output="unit", generator=None):
        """video [1,3,F,H,W] in [-1,1] (host or device, fp32/bf16) -> [1,3,F,H,W] on the GPU: bf16 in [0,1]
        (output="unit", what ref process_video returns) or uint8 = trunc(that*255) (output="uint8", what the
        reference's savers make of it, ref :124/:143/:168 — the runner gathers / copies 1 byte per element).
        `generator`: CUDA generator for every random draw of this unit (latent noise when `noise` is None, and the
        add_noise draw when noise_step != 0); None = the global generator (reference semantics, ref :409, :449-457)."""
        assert output in ("unit", "uint8")
        dev = self.device
        inter = {}
        video = video.to(dev, non_blocking=True)                                        # ref :407 (H2D)
        mom, (Tl, h, w) = self.vae.encode_cl(video if video.dtype in (torch.float32, BF) else video.to(BF))
        if noise is None:                                                               # ref :409 (global RNG)
            noise = torch.randn((1, 16, Tl, h, w), device=dev, dtype=BF, generator=generator)
        assert tuple(noise.shape) == (1, 16, Tl, h, w), (tuple(noise.shape), (1, 16, Tl, h, w))
        pt = self.transformer.config.patch_size_t
        ncopy = Tl % pt if pt is not None else 0                                        # ref :411-418
        if pt is not None:
            assert (Tl + ncopy) % pt == 0, "latent frames must be divisible by patch_size_t (ref :418)"
        z = torch.empty(16, Tl, h, w, dtype=BF, device=dev)
        L.gaussian_sample(mom, noise.contiguous(), z, Tl * h * w, self.vae_scaling_factor_image)
        latent = torch.empty(ncopy + Tl, 16, h, w, dtype=BF, device=dev)                # [F,16,h,w]  (ref :446)
        latent[ncopy:] = z.permute(1, 0, 2, 3)
        if ncopy:
            latent[:ncopy] = z[:, :1].permute(1, 0, 2, 3)
        Fl = latent.shape[0]
        if noise_step != 0:                                                             # ref :449-457
            # the reference draws randn_like(latent) on the PERMUTED view of the [B,C,F,H,W] tensor (ref :446, :451), i.e.
            # in (C, F, h, w) memory order: draw in that order, then view it frame-major like `latent`
            n = torch.randn((16, Fl, h, w), device=dev, dtype=latent.dtype, generator=generator).permute(1, 0, 2, 3)
            latent = self.scheduler.add_noise(latent, n, torch.tensor([noise_step]))
        emb = self._prompt_on_device(empty_prompt_embedding)                            # ref :423-428
        tc = self.transformer.config
        rope = None
        if tc.use_rotary_positional_embeddings:                                         # ref :467-480
            gh, gw = h // tc.patch_size, w // tc.patch_size
            rope = get_3d_rotary_pos_embed(tc.attention_head_dim, None, (gh, gw), (Fl + pt - 1) // pt,
                                           grid_type="slice", max_size=(gh, gw), device=dev)
        tok = self.transformer.forward_tokens(latent, emb, int(sr_noise_step), rope)     # ref :483-489
        a, b = self.scheduler.coefficients(int(sr_noise_step), BF)
        x0 = torch.empty_like(latent)
        pred = torch.empty_like(latent) if return_intermediates else None
        L.unpatchify_velocity(tok, latent, x0, pred, Fl, 16, h, w, a, b)                 # ref :491-493
        x0 = x0[ncopy:]                                                                 # ref :496-497
        zdec = x0.permute(1, 0, 2, 3).contiguous()[None]                                # [1,16,Tl,h,w]
        inv = 1.0 / self.vae_scaling_factor_image
        if return_intermediates:                                                        # raw decode kept for the tests
            dec = self.vae.decode_scaled(zdec, inv)                                     # ref :500
            out = torch.empty_like(dec)
            L.post_scale(dec, out)                                                      # ref :501
            if output == "uint8":
                out = (out.float() * 255.0).to(torch.uint8)
            inter.update(moments=LazyMoments(mom, (Tl, h, w)), latent=latent[None], pred=pred[None], x0=x0[None],
                         decoded=dec)
            return out, inter
        # ref :500-501 in one pass: `*0.5+0.5, clamp` (and the savers' uint8 quantisation) run in the last conv's epilogue
        return self.vae.decode_scaled(zdec, inv, post=output)


class LazyMoments:
    def __init__(self, mom, shape):
        self.mom, self.shape = mom, shape

    def ncthw(self):
        T, h, w = self.shape
        out = torch.empty(1, 32, T, h, w, dtype=BF, device=self.mom.device)
        L.cl_to_ncthw(self.mom, out, 32, T, h, w, 32)
        return out


@torch.no_grad()
def process_video(pipe: CogVideoXPipeline, video: torch.Tensor, prompt: str = "", noise_step: int = 0,
                  sr_noise_step: int = 399, empty_prompt_embedding: torch.Tensor = None):
    """Same signature/semantics as ref: /root/reference/inference_script.py:394-503."""
    if prompt != "" or empty_prompt_embedding is None:
        raise NotImplementedError("DOVE runs with the empty prompt and its pre-computed embedding (README.md:235)")
    return pipe.one_step_sr(video, empty_prompt_embedding, sr_noise_step=sr_noise_step, noise_step=noise_step)


def synthetic_prompt_embedding(n_tokens=226, dim=4096, seed=7, device="cpu"):
    """Stand-in for pretrained_models/prompt_embeddings/e3b0...b855.safetensors ([226,4096] bf16,
    mean 0.00175, std 0.1485 — statistics of the shipped file) for runs without /root/reference."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(n_tokens, dim, generator=g) * 0.1485 + 0.00175).to(BF).to(device)


def load_prompt_embedding(path):
    from safetensors.torch import load_file
    return load_file(str(path))["prompt_embedding"]
