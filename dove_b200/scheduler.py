"""CogVideoXDPMScheduler pieces on the DOVE hot path (mirror of the diffusers scheduler surface the reference
touches: construction via `from_config(pipe.scheduler.config, timestep_spacing="trailing")`
ref: /root/reference/inference_script.py:629-631; `add_noise` :457; `get_velocity(pred, latent, t)` :491-493).

alpha_bar table: beta = linspace(sqrt(b0), sqrt(b1), 1000, float64)^2, cumprod, SNR shift, zero-terminal-SNR
rescale — evaluated on the host in float64.  PARITY-CRITICAL: the reference casts alpha_bar to the sample dtype
BEFORE indexing and taking the square roots, so in bf16 the two coefficients at t = 399 are 0.625 / 0.78125
(not 0.62733 / 0.77875); `coefficients()` reproduces that rounding order.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import _lib as L
from .weights import SCHED_CONFIG


class CogVideoXDPMScheduler:
    def __init__(self, **kw):
        cfg = dict(SCHED_CONFIG)
        cfg.update(kw)
        self.config = SimpleNamespace(**cfg)
        c = self.config
        if c.beta_schedule != "scaled_linear":
            raise NotImplementedError(c.beta_schedule)
        betas = torch.linspace(c.beta_start ** 0.5, c.beta_end ** 0.5, c.num_train_timesteps, dtype=torch.float64) ** 2
        ac = torch.cumprod(1.0 - betas, dim=0)
        ac = ac / (c.snr_shift_scale + (1 - c.snr_shift_scale) * ac)
        if c.rescale_betas_zero_snr:
            s = ac.sqrt()
            s0, sT = s[0].clone(), s[-1].clone()
            s = (s - sT) * (s0 / (s0 - sT))
            ac = s ** 2
        self.alphas_cumprod = ac

    @classmethod
    def from_config(cls, config, **kw):
        d = dict(config) if isinstance(config, dict) else dict(vars(config))
        d.update(kw)
        return cls(**d)

    def coefficients(self, t: int, dtype=torch.bfloat16):
        """(sqrt(alpha_bar_t), sqrt(1 - alpha_bar_t)) with the table cast to `dtype` first."""
        a = self.alphas_cumprod.to(dtype)[t]
        return float(a ** 0.5), float((1 - a) ** 0.5)

    @staticmethod
    def _t(timesteps):
        t = timesteps.reshape(-1) if torch.is_tensor(timesteps) else torch.tensor([int(timesteps)])
        assert t.numel() == 1, "batch 1"
        return int(t[0].item())

    def get_velocity(self, sample, noise, timesteps):
        """sqrt(ab)*noise - sqrt(1-ab)*sample, bf16 with the reference's rounding points."""
        a, b = self.coefficients(self._t(timesteps), sample.dtype)
        sample, noise = sample.contiguous(), noise.contiguous()      # the reference passes permuted views (ref :446)
        out = torch.empty_like(sample)
        L.velocity(sample, noise, out, a, b)
        return out

    def add_noise(self, original_samples, noise, timesteps):
        """sqrt(ab)*x + sqrt(1-ab)*noise (only reached with --noise_step != 0, default 0)."""
        a, b = self.coefficients(self._t(timesteps), original_samples.dtype)
        original_samples, noise = original_samples.contiguous(), noise.contiguous()
        out = torch.empty_like(original_samples)
        L.velocity(noise, original_samples, out, a, -b)
        return out
