"""CPU oracle for the LaDiNE nested-ensemble reverse-diffusion sampler.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the CPU baseline, never as the thing shipped.

It restates, on plain state-dicts and torch CPU tensors, the arithmetic of the
reference's hot path (all citations relative to ``/root/reference/diffusion``):

* ``diffusion_utils.py:5-28``    make_beta_schedule
* ``diffusion_utils.py:31-35``   extract
* ``diffusion_utils.py:39-50``   q_sample
* ``diffusion_utils.py:54-92``   p_sample
* ``diffusion_utils.py:96-111``  p_sample_t_1to0
* ``diffusion_utils.py:114-130`` y_0_reparam
* ``diffusion_utils.py:133-163`` p_sample_loop
* ``latent_model.py:93-105``     ConditionalLinear
* ``latent_model.py:108-184``    ConditionalModel ('linear' encoder arch)
* ``classification_train_separately.py:215-226``  schedule tensors
* ``classification_train_separately.py:764-784``  nested-ensemble loop

The arithmetic itself lives in PyTorch (pinned torch==1.10.0 in the reference's
requirements.txt:59; torch 2.11 here, same operator semantics for
Linear / Embedding / BatchNorm1d(eval) / softplus / gather), so the FP32
restatement calls the same ``torch.nn.functional`` operators in the same order.

Parity pin: the reference ships NO tests or golden vectors, so
this oracle is pinned against outputs of the reference itself, generated in the
build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference/diffusion``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every fixture.

The noise the reference draws with ``torch.randn_like`` is passed in explicitly:
``noise[0]`` is the y_T draw, ``noise[k]`` (k >= 1) the draw of step t = T - k;
the final step (table index 0) draws none (diffusion_utils.py:139, :67, :96-111).

A second section restates the *packed form* the CUDA kernels implement (encoder
hoisted, BatchNorm + bias + gamma folded into per-step scale/shift rows, optional
rounding of the GEMM operands to fp16/bf16) so a kernel bug can be told apart
from a precision effect.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]
BN_EPS = 1e-5  # nn.BatchNorm1d default, latent_model.py:129,132,155,162,164,166


# ---------------------------------------------------------------------------
# schedules
# ---------------------------------------------------------------------------
def make_beta_schedule(schedule: str = "linear", num_timesteps: int = 1000,
                       start: float = 1e-5, end: float = 1e-2) -> torch.Tensor:
    """diffusion_utils.py:5-28 (seven schedule kinds)."""
    n = num_timesteps
    if schedule == "linear":
        return torch.linspace(start, end, n)
    if schedule == "const":
        return end * torch.ones(n)
    if schedule == "quad":
        return torch.linspace(start ** 0.5, end ** 0.5, n) ** 2
    if schedule == "jsd":
        return 1.0 / torch.linspace(n, 1, n)
    if schedule == "sigmoid":
        return torch.sigmoid(torch.linspace(-6, 6, n)) * (end - start) + start
    if schedule in ("cosine", "cosine_reverse"):
        s = 0.008

        def abar(i):
            return math.cos((i / n + s) / (1 + s) * math.pi / 2) ** 2

        return torch.tensor([min(1 - abar(i + 1) / abar(i), 0.999) for i in range(n)])
    if schedule == "cosine_anneal":
        return torch.tensor([start + 0.5 * (end - start) * (1 - math.cos(t / (n - 1) * math.pi))
                             for t in range(n)])
    raise ValueError(f"unknown beta schedule {schedule!r}")


def schedule_tensors(betas: torch.Tensor, schedule: str = "linear"):
    """classification_train_separately.py:215-226 -> (alphas, one_minus_alphas_bar_sqrt)."""
    betas = betas.float()
    alphas = 1.0 - betas
    omabs = torch.sqrt(1 - alphas.cumprod(dim=0))
    if schedule == "cosine":
        omabs = omabs * 0.9999
    return alphas, omabs


def extract(table: torch.Tensor, t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """diffusion_utils.py:31-35."""
    picked = torch.gather(table, 0, t)
    return picked.reshape(t.shape[0], *([1] * (like.dim() - 1)))


def q_sample(y, y_0_hat, alphas_bar_sqrt, one_minus_alphas_bar_sqrt, t, noise):
    """diffusion_utils.py:39-50."""
    sa = extract(alphas_bar_sqrt, t, y)
    so = extract(one_minus_alphas_bar_sqrt, t, y)
    return sa * y + (1 - sa) * y_0_hat + so * noise


# ---------------------------------------------------------------------------
# denoiser (ConditionalModel, 'linear' arch) on a plain state-dict
# ---------------------------------------------------------------------------
def _bn_eval(sd: StateDict, prefix: str, h: torch.Tensor) -> torch.Tensor:
    return F.batch_norm(h, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, BN_EPS)


def encoder_features(sd: StateDict, x: torch.Tensor) -> torch.Tensor:
    """``norm(encoder_x(x))`` -- latent_model.py:126-135, :155, :170-171 (step-invariant)."""
    if "encoder_x.weight" in sd:  # dataset == 'toy': a single Linear (latent_model.py:119-120)
        h = F.linear(x, sd["encoder_x.weight"], sd["encoder_x.bias"])
    else:
        h = F.linear(x, sd["encoder_x.0.weight"], sd["encoder_x.0.bias"])
        h = F.softplus(_bn_eval(sd, "encoder_x.1", h))
        h = F.linear(h, sd["encoder_x.3.weight"], sd["encoder_x.3.bias"])
        h = F.softplus(_bn_eval(sd, "encoder_x.4", h))
        h = F.linear(h, sd["encoder_x.6.weight"], sd["encoder_x.6.bias"])
    return _bn_eval(sd, "norm", h)


def _cond_linear(sd: StateDict, name: str, h: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """ConditionalLinear.forward -- latent_model.py:101-105."""
    out = F.linear(h, sd[name + ".lin.weight"], sd[name + ".lin.bias"])
    gamma = F.embedding(t, sd[name + ".embed.weight"])
    return gamma.view(-1, out.shape[-1]) * out


def trunk_forward(sd: StateDict, xf: torch.Tensor, y: torch.Tensor, t: torch.Tensor,
                  yhat: Optional[torch.Tensor]) -> torch.Tensor:
    """latent_model.py:172-184 given xf = norm(encoder_x(x))."""
    guidance = sd["lin1.lin.weight"].shape[1] == 2 * y.shape[-1]
    h = torch.cat([y, yhat], dim=-1) if guidance else y
    h = F.softplus(_bn_eval(sd, "unetnorm1", _cond_linear(sd, "lin1", h, t)))
    h = xf * h
    h = F.softplus(_bn_eval(sd, "unetnorm2", _cond_linear(sd, "lin2", h, t)))
    h = F.softplus(_bn_eval(sd, "unetnorm3", _cond_linear(sd, "lin3", h, t)))
    return F.linear(h, sd["lin4.weight"], sd["lin4.bias"])


def denoiser_forward(sd: StateDict, x, y, t, yhat=None) -> torch.Tensor:
    """ConditionalModel.forward as written (encoder re-evaluated) -- latent_model.py:169-184."""
    return trunk_forward(sd, encoder_features(sd, x), y, t, yhat)


# ---------------------------------------------------------------------------
# reverse process
# ---------------------------------------------------------------------------
class _Eps:
    """eps_theta provider: 'as written' re-runs the encoder every call, 'hoisted' caches xf."""

    def __init__(self, sd: StateDict, x: torch.Tensor, hoist: bool):
        self.sd, self.x, self.hoist = sd, x, hoist
        self.xf = encoder_features(sd, x) if hoist else None

    @classmethod
    def from_features(cls, sd: StateDict, xf: torch.Tensor) -> "_Eps":
        """Hoisted provider for callers that already hold xf = norm(encoder_x(x)) (no encoder keys needed)."""
        self = cls.__new__(cls)
        self.sd, self.x, self.hoist, self.xf = sd, None, True, xf
        return self

    def __call__(self, y, t, yhat):
        if self.hoist:
            return trunk_forward(self.sd, self.xf, y, t, yhat)
        return denoiser_forward(self.sd, self.x, y, t, yhat)


def p_sample(eps_fn, y, y_0_hat, y_T_mean, t: int, alphas, omabs, z) -> torch.Tensor:
    """diffusion_utils.py:54-92, with the randn_like draw ``z`` passed in."""
    tt = torch.tensor([t])
    alpha_t = extract(alphas, tt, y)
    s_t = extract(omabs, tt, y)
    s_tm1 = extract(omabs, tt - 1, y)
    q_t = (1 - s_t.square()).sqrt()
    q_tm1 = (1 - s_tm1.square()).sqrt()
    gamma_0 = (1 - alpha_t) * q_tm1 / (s_t.square())
    gamma_1 = (s_tm1.square()) * (alpha_t.sqrt()) / (s_t.square())
    gamma_2 = 1 + (q_t - 1) * (alpha_t.sqrt() + q_tm1) / (s_t.square())
    eps = eps_fn(y, tt, y_0_hat)
    y0r = 1 / q_t * (y - (1 - q_t) * y_T_mean - eps * s_t)
    mean = gamma_0 * y0r + gamma_1 * y + gamma_2 * y_T_mean
    beta_hat = (s_tm1.square()) / (s_t.square()) * (1 - alpha_t)
    return mean + beta_hat.sqrt() * z


def p_sample_t_1to0(eps_fn, y, y_0_hat, y_T_mean, omabs) -> torch.Tensor:
    """diffusion_utils.py:96-111 (table index 0, no noise)."""
    tt = torch.tensor([0])
    s_t = extract(omabs, tt, y)
    q_t = (1 - s_t.square()).sqrt()
    eps = eps_fn(y, tt, y_0_hat)
    return 1 / q_t * (y - (1 - q_t) * y_T_mean - eps * s_t)


def y_0_reparam(eps_fn, y, y_0_hat, y_T_mean, t: torch.Tensor, omabs) -> torch.Tensor:
    """diffusion_utils.py:114-130 (per-row t)."""
    s_t = extract(omabs, t, y)
    q_t = (1 - s_t.square()).sqrt()
    eps = eps_fn(y, t, y_0_hat)
    return 1 / q_t * (y - (1 - q_t) * y_T_mean - eps * s_t)


def p_sample_loop(sd: StateDict, x, y_0_hat, y_T_mean, n_steps: int, alphas, omabs,
                  noise: torch.Tensor, only_last_sample: bool = False, hoist: bool = False):
    """diffusion_utils.py:133-163.  ``noise``: [n_steps, B, C]."""
    assert noise.shape[0] >= n_steps
    eps_fn = x if isinstance(x, _Eps) else _Eps(sd, x, hoist)   # an _Eps (e.g. _Eps.from_features) may be passed as x
    cur = noise[0] + y_T_mean
    seq = [cur]
    for k, t in enumerate(reversed(range(1, n_steps)), start=1):
        cur = p_sample(eps_fn, cur, y_0_hat, y_T_mean, t, alphas, omabs, noise[k])
        seq.append(cur)
    y0 = p_sample_t_1to0(eps_fn, cur, y_0_hat, y_T_mean, omabs)
    if only_last_sample:
        return y0
    seq.append(y0)
    return seq


def ensemble_loop(sds: Sequence[StateDict], x, y0hats: Sequence[torch.Tensor], draws: int,
                  n_steps: int, alphas, omabs, noise: torch.Tensor, hoist: bool = True) -> torch.Tensor:
    """classification_train_separately.py:764-784: for member ii, ``draws`` sequential chains
    with y_0_hat = y_T_mean = target_pred[ii].  ``noise``: [K, D, n_steps, N, C] -> [K, D, N, C]."""
    out = []
    for k, sd in enumerate(sds):
        per = [p_sample_loop(sd, x, y0hats[k], y0hats[k], n_steps, alphas, omabs, noise[k, d],
                             only_last_sample=True, hoist=hoist) for d in range(draws)]
        out.append(torch.stack(per))
    return torch.stack(out)


# ---------------------------------------------------------------------------
# packed form (what the CUDA kernels compute) -- SURVEY.md §8a "algebraic restatement"
# ---------------------------------------------------------------------------
def coef_table(alphas: torch.Tensor, omabs: torch.Tensor, n_steps: int) -> torch.Tensor:
    """[n_steps, 8] FP32 rows (inv_q, 1-q, s, g0, g1, g2, sigma, 0), built with the reference's own
    torch expressions (diffusion_utils.py:69-78, :90, :100-101) evaluated elementwise over t."""
    a, s = alphas[:n_steps].float(), omabs[:n_steps].float()
    s_prev = torch.cat([s[:1], s[:-1]])  # row 0 never uses t-1
    q = (1 - s.square()).sqrt()
    q_prev = (1 - s_prev.square()).sqrt()
    g0 = (1 - a) * q_prev / (s.square())
    g1 = (s_prev.square()) * (a.sqrt()) / (s.square())
    g2 = 1 + (q - 1) * (a.sqrt() + q_prev) / (s.square())
    sig = ((s_prev.square()) / (s.square()) * (1 - a)).sqrt()
    tab = torch.stack([1 / q, 1 - q, s, g0, g1, g2, sig, torch.zeros_like(s)], dim=1)
    tab[0, 3:] = 0
    return tab.contiguous()


def fold_member(sd: StateDict, n_steps: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """A_l[t] = E_l[t] * s_l ; C_l[t] = E_l[t] * (s_l * b_l) + (beta_l - s_l * mean_l)."""
    out = {}
    C = sd["lin4.weight"].shape[0]
    for l in (1, 2, 3):
        bn = f"unetnorm{l}"
        s = sd[bn + ".weight"].to(dtype) / torch.sqrt(sd[bn + ".running_var"].to(dtype) + BN_EPS)
        shift = sd[bn + ".bias"].to(dtype) - s * sd[bn + ".running_mean"].to(dtype)
        E = sd[f"lin{l}.embed.weight"][:n_steps].to(dtype)
        out[f"A{l}"] = E * s
        out[f"C{l}"] = E * (s * sd[f"lin{l}.lin.bias"].to(dtype)) + shift
    W1 = sd["lin1.lin.weight"].to(dtype)
    out["W1y"] = W1[:, :C].contiguous()
    out["W1g"] = W1[:, C:].contiguous() if W1.shape[1] == 2 * C else None
    out["W2"], out["W3"] = sd["lin2.lin.weight"].to(dtype), sd["lin3.lin.weight"].to(dtype)
    out["W4"], out["b4"] = sd["lin4.weight"].to(dtype), sd["lin4.bias"].to(dtype)
    return out


def _round_to(x: torch.Tensor, operand_dtype) -> torch.Tensor:
    """Round to the GEMM operand type the way the kernels do: round-to-nearest-even, SATURATING at the largest finite
    value (cvt.rn.satfinite), so an activation beyond 65504 in FP16 becomes 65504, not inf."""
    if operand_dtype is None:
        return x
    if operand_dtype == "fp16x2":
        # the FP32X path: operand = FP16 hi + FP16 lo (the sum is exact in FP32: 22 significant bits); the kernel also
        # pre-scales W by a power of two so its lo parts stay normal -- emulated here by splitting a scaled copy
        m = float(x.abs().max())
        scale = 2.0 ** (14 - math.frexp(m)[1]) if (m > 0 and math.isfinite(m) and x.dim() == 2 and x.shape[0] == x.shape[1]) else 1.0
        xs = x * scale
        hi = xs.clamp(-65504, 65504).to(torch.float16).to(x.dtype)
        lo = (xs - hi).clamp(-65504, 65504).to(torch.float16).to(x.dtype)
        return (hi + lo) / scale
    lim = torch.finfo(operand_dtype).max
    return x.clamp(-lim, lim).to(operand_dtype).to(x.dtype)


def packed_sample(sd: StateDict, xf, y_0_hat, y_T_mean, n_steps: int, alphas, omabs, noise,
                  operand_dtype=None, dtype=torch.float32, trajectory: bool = False):
    """The kernel's algebra on the CPU.  ``operand_dtype`` (torch.float16 / torch.bfloat16 / None)
    rounds W2, W3, h1 and h2 exactly where the tensor-core path does; accumulation stays ``dtype``."""
    p = fold_member(sd, n_steps, dtype)
    tab = coef_table(alphas, omabs, n_steps).to(dtype)
    xf, yh, mu = xf.to(dtype), y_0_hat.to(dtype), y_T_mean.to(dtype)
    W2, W3 = _round_to(p["W2"], operand_dtype), _round_to(p["W3"], operand_dtype)
    u = yh @ p["W1g"].T if p["W1g"] is not None else torch.zeros_like(xf)
    noise = noise.to(dtype)

    def eps_at(y, t):
        h1 = F.softplus(p["A1"][t] * (y @ p["W1y"].T + u) + p["C1"][t]) * xf
        h2 = F.softplus(p["A2"][t] * (_round_to(h1, operand_dtype) @ W2.T) + p["C2"][t])
        h3 = F.softplus(p["A3"][t] * (_round_to(h2, operand_dtype) @ W3.T) + p["C3"][t])
        return h3 @ p["W4"].T + p["b4"]

    y = noise[0] + mu
    seq = [y]
    for k, t in enumerate(reversed(range(1, n_steps)), start=1):
        inv_q, omq, s, g0, g1, g2, sig, _ = tab[t]
        y0r = inv_q * (y - omq * mu - eps_at(y, t) * s)
        y = (g0 * y0r + g1 * y + g2 * mu) + sig * noise[k]
        seq.append(y)
    inv_q, omq, s = tab[0, :3]
    y0 = inv_q * (y - omq * mu - eps_at(y, 0) * s)
    seq.append(y0)
    return seq if trajectory else y0


# ---------------------------------------------------------------------------
# ensemble statistics (consumers of the K*D samples) -- SURVEY.md §8f-1
# ---------------------------------------------------------------------------
def majority_vote(samples: torch.Tensor) -> torch.Tensor:
    """classification_train_separately.py:51-68.  samples [S, N, C] -> [N] int64; ties resolve to
    the smallest label (sorted torch.unique + first argmax)."""
    votes = samples.argmax(dim=-1).T  # [N, S]
    out = []
    for row in votes:
        labels, counts = torch.unique(row, return_counts=True)
        out.append(labels[counts.argmax()])
    return torch.stack(out)


def convert_to_prob(samples: torch.Tensor, temperature: float) -> torch.Tensor:
    """classification_train_separately.py:392-398."""
    return torch.softmax(((samples - 1.0) ** 2) * (-1.0) / temperature, dim=-1)


def ensemble_confidence(samples: torch.Tensor, temperature: float) -> torch.Tensor:
    """classification_train_separately.py:425-447: mean over the K*D axis of convert_to_prob."""
    return convert_to_prob(samples, temperature).mean(dim=0)


def ece_l1(probs: torch.Tensor, target: torch.Tensor, n_bins: int = 10) -> torch.Tensor:
    """torchmetrics 0.11.4 MulticlassCalibrationError(n_bins=10, norm='l1') as called at
    classification_train_separately.py:413-423: top-1 confidence/accuracy, uniform bins on [0,1],
    bucketize(right=True) - 1, sum_b |acc_b - conf_b| * prop_b.  torchmetrics is absent from this
    image, so this restates its published algorithm ("parity unpinned" for this one function)."""
    conf, pred = probs.max(dim=-1)
    acc = (pred == target).to(conf.dtype)
    edges = torch.linspace(0, 1, n_bins + 1, dtype=conf.dtype)
    idx = torch.bucketize(conf, edges, right=True) - 1
    idx = idx.clamp(0, n_bins - 1)
    count = torch.zeros(n_bins, dtype=conf.dtype).scatter_add_(0, idx, torch.ones_like(conf))
    conf_b = torch.zeros(n_bins, dtype=conf.dtype).scatter_add_(0, idx, conf)
    acc_b = torch.zeros(n_bins, dtype=conf.dtype).scatter_add_(0, idx, acc)
    nz = count > 0
    conf_b = torch.where(nz, conf_b / count.clamp(min=1), torch.zeros_like(conf_b))
    acc_b = torch.where(nz, acc_b / count.clamp(min=1), torch.zeros_like(acc_b))
    prop = count / count.sum()
    return torch.sum(torch.abs(acc_b - conf_b) * prop)


def mean_piw_per_class(samples: torch.Tensor, mv: torch.Tensor, label: torch.Tensor):
    """classification_train_separately.py:102-140.  samples [S, N, C]."""
    lo = torch.quantile(samples, q=0.025, dim=0)
    hi = torch.quantile(samples, q=0.975, dim=0)
    piw = (hi - lo)[torch.arange(samples.shape[1]), mv]
    C = samples.shape[2]
    good, bad = torch.zeros(C), torch.zeros(C)
    for c in range(C):
        sel = mv == c
        good[c] = piw[sel & (mv == label)].mean()
        bad[c] = piw[sel & (mv != label)].mean()
    return good, bad


def class_variances(samples: torch.Tensor, pred: torch.Tensor, truth: torch.Tensor):
    """classification_train_separately.py:143-174.  samples [S, N, C]."""
    C = samples.shape[2]
    good, bad = torch.zeros(C), torch.zeros(C)
    for c in range(C):
        ok = (pred == c) & (truth == c)
        ko = (pred == c) & (truth != c)
        if ok.any():
            good[c] = samples[:, ok, c].var(dim=0).mean()
        if ko.any():
            bad[c] = samples[:, ko, c].var(dim=0).mean()
    return good, bad


# ---------------------------------------------------------------------------
# deterministic synthetic members / inputs (shared by golden generation, tests and bench)
# ---------------------------------------------------------------------------
def synth_state_dict(seed: int, F_dim: int, H_dim: int, Dx: int, C: int, T: int,
                     guidance: bool = True, randomize_bn: bool = True,
                     eps_gain: float = 1.0) -> StateDict:
    """A ConditionalModel('linear') state-dict with nn.Linear-style U(-1/sqrt(in), 1/sqrt(in)) weights,
    U(0,1) gamma tables (latent_model.py:99) and, by default, randomised BatchNorm running stats and
    affine parameters so the folding is non-trivial.  Generated tensor by
    tensor from one torch.Generator, independent of any nn.Module construction order."""
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f, gain=1.0):
        b = gain / math.sqrt(in_f)
        return ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * b,
                (torch.rand(out_f, generator=g) * 2 - 1) * b)

    def bn(n):
        if randomize_bn:
            return {"weight": torch.rand(n, generator=g) + 0.5,
                    "bias": torch.randn(n, generator=g) * 0.2,
                    "running_mean": torch.randn(n, generator=g) * 0.3,
                    "running_var": torch.rand(n, generator=g) + 0.5,
                    "num_batches_tracked": torch.tensor(0)}
        return {"weight": torch.ones(n), "bias": torch.zeros(n), "running_mean": torch.zeros(n),
                "running_var": torch.ones(n), "num_batches_tracked": torch.tensor(0)}

    sd: StateDict = {}
    for idx, (o, i) in zip((0, 3, 6), ((H_dim, Dx), (H_dim, H_dim), (F_dim, H_dim))):
        sd[f"encoder_x.{idx}.weight"], sd[f"encoder_x.{idx}.bias"] = lin(o, i)
    for idx in (1, 4):
        for k, v in bn(H_dim).items():
            sd[f"encoder_x.{idx}.{k}"] = v
    for k, v in bn(F_dim).items():
        sd[f"norm.{k}"] = v
    ins = (2 * C if guidance else C, F_dim, F_dim)
    for l, i in zip((1, 2, 3), ins):
        sd[f"lin{l}.lin.weight"], sd[f"lin{l}.lin.bias"] = lin(F_dim, i)
        sd[f"lin{l}.embed.weight"] = torch.rand(T + 1, F_dim, generator=g)
        for k, v in bn(F_dim).items():
            sd[f"unetnorm{l}.{k}"] = v
    sd["lin4.weight"], sd["lin4.bias"] = lin(C, F_dim, eps_gain)
    return sd


def synth_inputs(seed: int, B: int, Dx: int, C: int):
    """x ~ U(0,1) (ToTensor range), y_0_hat = softmax(2 * randn) -- SURVEY.md §8d config 1."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, Dx, generator=g)
    yhat = torch.softmax(2 * torch.randn(B, C, generator=g), dim=1)
    return x, yhat


def synth_noise(seed: int, shape) -> torch.Tensor:
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def to_dtype(sd: StateDict, dtype) -> StateDict:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
