"""Drop-in for the reference's ``diffusion_utils`` module (star-imported by the runner,
classification_train_separately.py:20).  Same names, positional order, defaults and return types;
the reverse process (``p_sample_loop`` / ``p_sample`` / ``p_sample_t_1to0``) runs in the fused
sm_100a kernels behind include/ladine.h instead of ~60 eager ops per step.

Differences a caller can observe:
  * the model must be in ``eval()`` mode, on a CUDA device, and the call must not need gradients
    (``output_detach=True`` under ``torch.no_grad()`` is how the reference always calls it,
    classification_train_separately.py:692-694, :768); anything else raises ``NotImplementedError``
    -- there is no CPU / eager fallback;
  * keyword-only extras: ``noise=`` injects the N(0,1) draws the reference would take from
    ``torch.randn_like`` (``noise[0]`` -> y_T, ``noise[k]`` -> step t = n_steps - k), ``seed=``
    fixes the counter-based Philox stream, ``precision=`` picks the arithmetic of the packed member.
    Without either, the seed is drawn from torch's global CPU generator (``torch.manual_seed``);
  * opt-in ``set_draws_ahead(D)`` / ``LADINE_DRAWS_AHEAD=D``: the runner asks for its posterior draws one call at a
    time -- ``for trial in range(20): p_sample_loop(same member, same images, ...)``
    (classification_train_separately.py:770-777).  With draws-ahead the FIRST such call samples D independent chains
    per row in one batched launch and the next D-1 calls with the very same inputs are served from that batch
    (see ``p_sample_loop``); the samples are i.i.d. draws of the same posterior either way.
"""
from __future__ import annotations

import os
import weakref

import torch

from . import engine
from .schedule import coef_table, make_beta_schedule  # noqa: F401  (re-exported, diffusion_utils.py:5)

__all__ = ["make_beta_schedule", "extract", "q_sample", "p_sample", "p_sample_t_1to0", "y_0_reparam",
           "p_sample_loop"]

# ------------------------------------------------------------------------------------------------
# draws-ahead (opt-in): serve the runner's D sequential p_sample_loop calls from one batched launch
# ------------------------------------------------------------------------------------------------
_DRAWS_AHEAD = max(0, int(os.environ.get("LADINE_DRAWS_AHEAD", "0") or 0))
_AHEAD: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()   # model -> _Ahead


def set_draws_ahead(draws: int) -> int:
    """Enable (``draws`` >= 2) or disable (0 / 1) draws-ahead for ``p_sample_loop``; returns the previous setting.
    Set it to the runner's ``mc_trials`` (20, classification_train_separately.py:770)."""
    global _DRAWS_AHEAD
    prev, _DRAWS_AHEAD = _DRAWS_AHEAD, max(0, int(draws))
    _AHEAD.clear()
    return prev


def _tensor_key(t: torch.Tensor) -> tuple:
    return (t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape), tuple(t.stride()), t._version, t.dtype,
            str(t.device))


class _Ahead:
    """D chains per row sampled by one call, handed out one draw per call while the inputs are provably the same:
    same storages / views / in-place version counters (strong references are kept, so no address can be recycled for
    other values), same packed member and same encoder features (both are re-made when a parameter changes)."""
    __slots__ = ("key", "refs", "samples", "next")

    def __init__(self, key, refs, samples):
        self.key, self.refs, self.samples, self.next = key, refs, samples, 1


def extract(input, t, x):
    """diffusion_utils.py:31-35."""
    out = torch.gather(input, 0, t.to(input.device))
    return out.reshape(t.shape[0], *([1] * (len(x.shape) - 1)))


def q_sample(y, y_0_hat, alphas_bar_sqrt, one_minus_alphas_bar_sqrt, t, noise=None):
    """Forward noising q(y_t | y_0, x), diffusion_utils.py:39-50 (training only; plain PyTorch)."""
    if noise is None:
        noise = torch.randn_like(y)
    sqrt_alpha_bar_t = extract(alphas_bar_sqrt, t, y)
    sqrt_one_minus_alpha_bar_t = extract(one_minus_alphas_bar_sqrt, t, y)
    return sqrt_alpha_bar_t * y + (1 - sqrt_alpha_bar_t) * y_0_hat + sqrt_one_minus_alpha_bar_t * noise


def y_0_reparam(model, x, y, y_0_hat, y_T_mean, t, one_minus_alphas_bar_sqrt, output_detach=True):
    """diffusion_utils.py:114-130: y_0 from q(y_t | y_0) with per-row ``t``.  Not on the test path
    (no caller in the reference); kept as PyTorch code for API parity."""
    device = next(model.parameters()).device
    sqrt_one_minus_alpha_bar_t = extract(one_minus_alphas_bar_sqrt, t, y)
    sqrt_alpha_bar_t = (1 - sqrt_one_minus_alpha_bar_t.square()).sqrt()
    eps_theta = model(x, y, t, y_0_hat).to(device)
    if output_detach:
        eps_theta = eps_theta.detach()
    return (1 / sqrt_alpha_bar_t * (y - (1 - sqrt_alpha_bar_t) * y_T_mean - eps_theta * sqrt_one_minus_alpha_bar_t)).to(device)


# ------------------------------------------------------------------------------------------------
def _prepare(model, x, y_0_hat, y_T_mean, output_detach, precision):
    if not output_detach or (torch.is_grad_enabled() and any(
            t.requires_grad for t in (x, y_0_hat, y_T_mean) if torch.is_tensor(t))):
        raise NotImplementedError("the fused sampler does not build an autograd graph (output_detach=False / "
                                  "inputs requiring grad); the reference only samples under torch.no_grad()")
    if getattr(model, "training", False):
        raise NotImplementedError("model.train() (batch-statistics BatchNorm) is not accelerated: call model.eval()")
    device = model.device if isinstance(model, engine.PackedModel) else next(model.parameters()).device
    pm = engine.packed_member_of(model, precision)
    for name, t in (("x", x), ("y_0_hat", y_0_hat), ("y_T_mean", y_T_mean)):
        if t.device != device:
            raise ValueError(f"{name} is on {t.device} but the model is on {device}")
    if y_0_hat.dim() != 2 or y_0_hat.shape != y_T_mean.shape or y_0_hat.shape[0] != x.shape[0]:
        raise ValueError("expected x [B, ...], y_0_hat [B, C], y_T_mean [B, C]")
    xf = engine.features_of(model, x).unsqueeze(0)   # once per (model, image tensor): see engine.features_of
    return pm, xf, y_0_hat.unsqueeze(0), y_T_mean.unsqueeze(0)


def p_sample_loop(model, x, y_0_hat, y_T_mean, n_steps, alphas, one_minus_alphas_bar_sqrt, only_last_sample=False,
                  input_model_original_version=True, output_detach=True, *, noise=None, seed=None, precision="auto",
                  draws=None, persistent=True, draws_ahead=None):
    """Full reverse chain y_T -> y_0, diffusion_utils.py:133-163.

    Returns y_0 ``[B, C]`` when ``only_last_sample`` else the list ``[y_T, ..., y_1, y_0]`` of
    ``n_steps + 1`` tensors, on the model's device, FP32, detached.

    ``draws=D`` (keyword-only extension) runs D independent chains per input row in ONE launch -- what the runner
    obtains with D sequential calls (classification_train_separately.py:770-777) -- and returns ``[D, B, C]``
    (needs ``only_last_sample=True``; ``noise``, if given, is ``[D, n_steps, B, C]``).

    ``draws_ahead`` (default: the module setting, ``set_draws_ahead`` / ``LADINE_DRAWS_AHEAD``; off unless set): with
    D >= 2, a plain call (``only_last_sample=True``, no ``noise`` / ``seed`` / ``draws``) samples D chains per row in one
    launch, returns the first, and the following D-1 calls with the same model and the very same input tensors
    (storage, view, in-place version, packed weights, encoder features all unchanged) return the remaining draws without
    touching the GPU again -- the runner's ``for trial in range(20)`` loop then costs one batched launch per member
    instead of 20 chains one after the other.  Any change of an input starts a new batch.

    ``persistent`` (default True): a call of at most 128 chains at a tensor-core width runs as ONE cooperative launch
    for the whole chain (``engine.sample_chains(persistent=True)``) -- the shape of the runner's own calls (70 images,
    one draw); larger calls use the tile kernels either way."""
    if not input_model_original_version:
        model = model.conditional_model
    pm, xf, yh, mu = _prepare(model, x, y_0_hat, y_T_mean, output_detach, precision)
    ahead = _DRAWS_AHEAD if draws_ahead is None else max(0, int(draws_ahead))
    if ahead >= 2 and only_last_sample and noise is None and seed is None and draws is None:
        key = (id(pm), xf.untyped_storage().data_ptr(), int(n_steps), ahead, bool(persistent)) + tuple(
            _tensor_key(t) for t in (x, y_0_hat, y_T_mean, alphas, one_minus_alphas_bar_sqrt))
        st = _AHEAD.get(model)
        if st is not None and st.key == key and st.next < st.samples.shape[0]:
            st.next += 1
            return st.samples[st.next - 1].clone()
        samples = engine.sample_chains([pm], xf, yh, mu, coef_table(alphas, one_minus_alphas_bar_sqrt, n_steps), ahead,
                                       seed=engine.fresh_seed(), persistent=persistent)["y"][0]          # [D, B, C]
        _AHEAD[model] = _Ahead(key, (pm, xf, x, y_0_hat, y_T_mean, alphas, one_minus_alphas_bar_sqrt), samples)
        return samples[0].clone()
    coef = coef_table(alphas, one_minus_alphas_bar_sqrt, n_steps)
    B, Cc = y_0_hat.shape
    if draws is not None:
        D = int(draws)
        if D < 1 or not only_last_sample:
            raise ValueError("draws= needs draws >= 1 and only_last_sample=True")
        if noise is not None:
            if tuple(noise.shape) != (D, n_steps, B, Cc):
                raise ValueError(f"noise must be [draws={D}, n_steps={n_steps}, {B}, {Cc}]")
            noise = noise.reshape(1, D, n_steps, B, Cc)
        elif seed is None:
            seed = engine.fresh_seed()
        return engine.sample_chains([pm], xf, yh, mu, coef, D, noise=noise, seed=seed or 0, persistent=persistent)["y"][0]
    if noise is not None:
        if noise.shape[0] < n_steps or tuple(noise.shape[1:]) != (B, Cc):
            raise ValueError(f"noise must be [n_steps={n_steps}, {B}, {Cc}]")
        noise = noise[:n_steps].reshape(1, 1, n_steps, B, Cc)
    elif seed is None:
        seed = engine.fresh_seed()
    out = engine.sample_chains([pm], xf, yh, mu, coef, 1, noise=noise, seed=seed or 0,
                               trajectory=not only_last_sample, persistent=persistent)
    if only_last_sample:
        return out["y"][0, 0]
    return list(out["traj"][0, 0].unbind(0))


def p_sample(model, x, y, y_0_hat, y_T_mean, t, alphas, one_minus_alphas_bar_sqrt, output_detach=True, *,
             noise=None, seed=None, precision="auto", persistent=True):
    """One reverse step y_t -> y_{t-1} at table index ``t`` (>= 1), diffusion_utils.py:54-92."""
    t = int(t)
    if t < 1:
        raise ValueError("p_sample needs t >= 1 (it reads the schedule at t-1); use p_sample_t_1to0 for the last step")
    pm, xf, yh, mu = _prepare(model, x, y_0_hat, y_T_mean, output_detach, precision)
    coef = coef_table(alphas, one_minus_alphas_bar_sqrt, alphas.shape[0])
    if noise is not None:
        noise = noise.reshape(1, 1, 1, *y.shape)
    elif seed is None:
        seed = engine.fresh_seed()
    out = engine.sample_chains([pm], xf, yh, mu, coef, 1, t_first=t, t_last=t, y_init=y.reshape(1, 1, *y.shape),
                               noise=noise, seed=seed or 0, persistent=persistent)
    return out["y"][0, 0]


def p_sample_t_1to0(model, x, y, y_0_hat, y_T_mean, one_minus_alphas_bar_sqrt, output_detach=True, *,
                    precision="auto", persistent=True):
    """Last reverse step y_1 -> y_0 (table index 0, no noise), diffusion_utils.py:96-111."""
    pm, xf, yh, mu = _prepare(model, x, y_0_hat, y_T_mean, output_detach, precision)
    # row 0 of the coefficient table needs only one_minus_alphas_bar_sqrt[0]
    s = one_minus_alphas_bar_sqrt.detach()[:1].to("cpu", torch.float32)
    q = (1 - s.square()).sqrt()
    coef = torch.zeros(1, 8)
    coef[0, 0], coef[0, 1], coef[0, 2] = (1 / q)[0], (1 - q)[0], s[0]
    out = engine.sample_chains([pm], xf, yh, mu, coef, 1, t_first=0, t_last=0, y_init=y.reshape(1, 1, *y.shape),
                               persistent=persistent)
    return out["y"][0, 0]
