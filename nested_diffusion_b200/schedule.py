"""Diffusion schedule helpers of the sampler (host side, PyTorch).

``make_beta_schedule`` mirrors diffusion_utils.py:5-28; ``schedule_tensors`` the runner's derived
vectors (classification_train_separately.py:215-226); ``coef_table`` evaluates the per-step scalars
of ``p_sample`` / ``p_sample_t_1to0`` (diffusion_utils.py:69-78, :90, :100-101) once per schedule,
with the same FP32 torch expressions the reference evaluates every step, so the kernels consume
bit-identical coefficients.
"""
from __future__ import annotations

import math

import torch


def make_beta_schedule(schedule="linear", num_timesteps=1000, start=1e-5, end=1e-2):
    if schedule == "linear":
        betas = torch.linspace(start, end, num_timesteps)
    elif schedule == "const":
        betas = end * torch.ones(num_timesteps)
    elif schedule == "quad":
        betas = torch.linspace(start ** 0.5, end ** 0.5, num_timesteps) ** 2
    elif schedule == "jsd":
        betas = 1.0 / torch.linspace(num_timesteps, 1, num_timesteps)
    elif schedule == "sigmoid":
        betas = torch.sigmoid(torch.linspace(-6, 6, num_timesteps)) * (end - start) + start
    elif schedule in ("cosine", "cosine_reverse"):
        offset = 0.008

        def alpha_bar(i):
            return math.cos((i / num_timesteps + offset) / (1 + offset) * math.pi / 2) ** 2

        betas = torch.tensor([min(1 - alpha_bar(i + 1) / alpha_bar(i), 0.999) for i in range(num_timesteps)])
    elif schedule == "cosine_anneal":
        betas = torch.tensor([start + 0.5 * (end - start) * (1 - math.cos(t / (num_timesteps - 1) * math.pi))
                              for t in range(num_timesteps)])
    else:
        raise ValueError(f"unknown beta schedule {schedule!r}")
    return betas


def schedule_tensors(betas: torch.Tensor, schedule: str = "linear"):
    """-> (alphas, one_minus_alphas_bar_sqrt) as the runner builds them."""
    betas = betas.float()
    alphas = 1.0 - betas
    one_minus_alphas_bar_sqrt = torch.sqrt(1 - alphas.cumprod(dim=0))
    if schedule == "cosine":
        one_minus_alphas_bar_sqrt = one_minus_alphas_bar_sqrt * 0.9999
    return alphas, one_minus_alphas_bar_sqrt


def coef_table(alphas: torch.Tensor, one_minus_alphas_bar_sqrt: torch.Tensor, n_steps: int) -> torch.Tensor:
    """HOST FP32 [n_steps, 8]: inv_q, 1-q, s, gamma_0, gamma_1, gamma_2, sqrt(beta_hat), 0 per table index t.

    Row 0 is the noiseless last step (only the first three entries are used)."""
    if alphas.shape[0] < n_steps or one_minus_alphas_bar_sqrt.shape[0] < n_steps:
        raise ValueError("schedule vectors are shorter than n_steps")
    a = alphas.detach()[:n_steps].to("cpu", torch.float32)
    s = one_minus_alphas_bar_sqrt.detach()[:n_steps].to("cpu", torch.float32)
    s_prev = torch.cat([s[:1], s[:-1]])
    q = (1 - s.square()).sqrt()
    q_prev = (1 - s_prev.square()).sqrt()
    gamma_0 = (1 - a) * q_prev / (s.square())
    gamma_1 = (s_prev.square()) * (a.sqrt()) / (s.square())
    gamma_2 = 1 + (q - 1) * (a.sqrt() + q_prev) / (s.square())
    sigma = ((s_prev.square()) / (s.square()) * (1 - a)).sqrt()
    table = torch.stack([1 / q, 1 - q, s, gamma_0, gamma_1, gamma_2, sigma, torch.zeros_like(s)], dim=1)
    table[0, 3:] = 0
    return table.contiguous()
