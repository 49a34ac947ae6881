"""Oracle: CogVideoXDPMScheduler pieces on the hot path (restates diffusers
`schedulers/scheduling_dpm_cogvideox.py::{__init__, rescale_zero_terminal_snr, get_velocity, add_noise}`).
TEST INFRASTRUCTURE.  Reference call sites: /root/reference/inference_script.py:629-631 (construction),
:457 (add_noise, only if --noise_step != 0), :491-493 (get_velocity(pred, latent, t))."""
from types import SimpleNamespace

import torch

SCHED_CONFIG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                    beta_schedule="scaled_linear", prediction_type="v_prediction",
                    rescale_betas_zero_snr=True, snr_shift_scale=1.0, timestep_spacing="trailing")


def rescale_zero_terminal_snr(alphas_cumprod):
    s = alphas_cumprod.sqrt()
    s0, sT = s[0].clone(), s[-1].clone()
    s = s - sT
    s = s * (s0 / (s0 - sT))
    return s ** 2


class OracleCogVideoXDPMScheduler:
    def __init__(self, **kw):
        cfg = dict(SCHED_CONFIG)
        cfg.update(kw)
        self.config = SimpleNamespace(**cfg)
        c = self.config
        betas = torch.linspace(c.beta_start ** 0.5, c.beta_end ** 0.5, c.num_train_timesteps,
                               dtype=torch.float64) ** 2
        ac = torch.cumprod(1.0 - betas, dim=0)
        ac = ac / (c.snr_shift_scale + (1 - c.snr_shift_scale) * ac)
        if c.rescale_betas_zero_snr:
            ac = rescale_zero_terminal_snr(ac)
        self.alphas_cumprod = ac

    @classmethod
    def from_config(cls, config, **kw):
        d = dict(vars(config)) if not isinstance(config, dict) else dict(config)
        d.update(kw)
        return cls(**d)

    def _coeffs(self, like, timesteps):
        ac = self.alphas_cumprod.to(device=like.device).to(dtype=like.dtype)   # cast BEFORE index/sqrt
        t = timesteps.to(like.device)
        a = (ac[t] ** 0.5).flatten()
        b = ((1 - ac[t]) ** 0.5).flatten()
        while a.dim() < like.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a, b

    def get_velocity(self, sample, noise, timesteps):
        a, b = self._coeffs(sample, timesteps)
        return a * noise - b * sample

    def add_noise(self, original_samples, noise, timesteps):
        a, b = self._coeffs(original_samples, timesteps)
        return a * original_samples + b * noise
