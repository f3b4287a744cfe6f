"""Diffusion schedule holder — same constructor / attributes as the reference's
diffbir.model.Diffusion (model/gaussian_diffusion.py:77-129); host-side numpy fp64 tables."""
from __future__ import annotations

import numpy as np
import torch


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    if schedule != "linear":
        raise ValueError(f"schedule '{schedule}' is not on the inference hot path")
    return np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2


def enforce_zero_terminal_snr(betas: np.ndarray) -> np.ndarray:
    """Zero-terminal-SNR rescale (arXiv 2305.08891), gaussian_diffusion.py:49-72."""
    ab_sqrt = np.sqrt(np.cumprod(1.0 - betas))
    first, last = ab_sqrt[0].copy(), ab_sqrt[-1].copy()
    ab_sqrt = (ab_sqrt - last) * (first / (first - last))
    ab = ab_sqrt ** 2
    alphas = np.concatenate([ab[:1], ab[1:] / ab[:-1]])
    return 1.0 - alphas


class Diffusion:
    def __init__(self, timesteps=1000, beta_schedule="linear", loss_type="l2", linear_start=1e-4,
                 linear_end=2e-2, cosine_s=8e-3, parameterization="eps", zero_snr=False):
        assert parameterization in ("eps", "x0", "v")
        self.num_timesteps = timesteps
        self.parameterization = parameterization
        self.zero_snr = zero_snr
        betas = make_beta_schedule(beta_schedule, timesteps, linear_start, linear_end, cosine_s)
        if zero_snr:
            betas = enforce_zero_terminal_snr(betas)
        self.betas = betas
        ac = np.cumprod(1.0 - betas, axis=0)
        self.sqrt_alphas_cumprod = torch.tensor(np.sqrt(ac), dtype=torch.float32)
        self.sqrt_one_minus_alphas_cumprod = torch.tensor(np.sqrt(1.0 - ac), dtype=torch.float32)

    def to(self, device):
        self.sqrt_alphas_cumprod = self.sqrt_alphas_cumprod.to(device)
        self.sqrt_one_minus_alphas_cumprod = self.sqrt_one_minus_alphas_cumprod.to(device)
        return self

    def eval(self):
        return self

    def q_sample(self, x_start, t, noise):
        a = self.sqrt_alphas_cumprod.to(x_start.device)[t].view(-1, 1, 1, 1)
        s = self.sqrt_one_minus_alphas_cumprod.to(x_start.device)[t].view(-1, 1, 1, 1)
        return a * x_start + s * noise
