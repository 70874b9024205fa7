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


_COEF_CACHE: list = []   # [(key, (alphas, omabs) strong refs, table)], most recent first
_COEF_CACHE_SIZE = 8


def _vec_key(t: torch.Tensor) -> tuple:
    return (t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape), tuple(t.stride()), t._version, t.dtype,
            str(t.device))


def coef_table(alphas: torch.Tensor, one_minus_alphas_bar_sqrt: torch.Tensor, n_steps: int) -> torch.Tensor:
    """HOST FP32 [n_steps, 8]: inv_q, 1-q, s, gamma_0, gamma_1, gamma_2, sqrt(beta_hat), 0 per table index t.

    Row 0 is the noiseless last step (only the first three entries are used).  The table (READ-ONLY for callers) is
    remembered for the last few (alphas, one_minus_alphas_bar_sqrt, n_steps): the runner passes the same two device
    vectors to every one of its K x 20 ``p_sample_loop`` calls per batch, and rebuilding the table means two
    device->host copies, i.e. a stream synchronisation that would serialise back-to-back calls.  A hit needs the same
    storages / views / in-place version counters (strong references are held, so an address cannot be recycled)."""
    if alphas.shape[0] < n_steps or one_minus_alphas_bar_sqrt.shape[0] < n_steps:
        raise ValueError("schedule vectors are shorter than n_steps")
    key = (_vec_key(alphas), _vec_key(one_minus_alphas_bar_sqrt), int(n_steps))
    for i, (k, _, table) in enumerate(_COEF_CACHE):
        if k == key:
            if i:
                _COEF_CACHE.insert(0, _COEF_CACHE.pop(i))
            return table
    table = _coef_table(alphas, one_minus_alphas_bar_sqrt, int(n_steps))
    _COEF_CACHE.insert(0, (key, (alphas, one_minus_alphas_bar_sqrt), table))
    del _COEF_CACHE[_COEF_CACHE_SIZE:]
    return table


def _coef_table(alphas: torch.Tensor, one_minus_alphas_bar_sqrt: torch.Tensor, n_steps: int) -> torch.Tensor:
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
