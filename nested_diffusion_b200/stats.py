"""Ensemble statistics over the K*D posterior samples, vectorised PyTorch that
runs on whatever device the gathered samples live on.

All functions take ``samples`` as ONE tensor ``[S, N, C]`` (S = K*D chains per image) instead of the
reference's Python list of S ``[N, C]`` tensors; ``as_sample_tensor`` converts either form.
Semantics follow classification_train_separately.py:51-68 (majority vote, ties -> smallest label),
:392-398 (convert_to_prob), :425-447 (ensemble confidence), :413-423 (ECE = torchmetrics 0.11.4
MulticlassCalibrationError(n_bins=10, norm='l1')), :102-140 (PIW) and :143-174 (variances).
"""
from __future__ import annotations

from typing import Sequence, Union

import torch

Samples = Union[torch.Tensor, Sequence[torch.Tensor]]


def as_sample_tensor(samples: Samples) -> torch.Tensor:
    if torch.is_tensor(samples):
        if samples.dim() == 4:  # [K, D, N, C]
            return samples.reshape(-1, *samples.shape[2:])
        return samples
    return torch.stack(list(samples))


def majority_voting_for_mc_samples(predictions: Samples) -> torch.Tensor:
    """[S, N, C] -> [N] int64.  Mode of the per-chain argmax; a tie goes to the smallest label, as the
    reference's sorted ``torch.unique`` + first ``argmax`` does."""
    s = as_sample_tensor(predictions)
    votes = s.argmax(dim=-1)  # [S, N]
    counts = torch.zeros(s.shape[1], s.shape[2], dtype=torch.int64, device=s.device)
    counts.scatter_add_(1, votes.T.contiguous(), torch.ones_like(votes.T, dtype=torch.int64))
    return counts.argmax(dim=1)  # first maximal index == smallest label on ties


def convert_to_prob(logits: torch.Tensor, temperature: float) -> torch.Tensor:
    return torch.softmax(((logits - 1.0) ** 2) * (-1.0) / temperature, dim=-1)


def compute_ensemble_confidence(outputs: Samples, temperature: float) -> torch.Tensor:
    """Mean over the chain axis of ``convert_to_prob`` -> [N, C].  (Unlike the reference, the input
    is not overwritten in place.)"""
    return convert_to_prob(as_sample_tensor(outputs), temperature).mean(dim=0)


def compute_ece(probs: torch.Tensor, target: torch.Tensor, n_bins: int = 10) -> torch.Tensor:
    """Top-label expected calibration error, l1 norm, ``n_bins`` uniform bins on [0, 1]."""
    conf, pred = probs.max(dim=-1)
    acc = (pred == target.to(pred.device)).to(conf.dtype)
    edges = torch.linspace(0, 1, n_bins + 1, dtype=conf.dtype, device=conf.device)
    idx = (torch.bucketize(conf, edges, right=True) - 1).clamp(0, n_bins - 1)
    zeros = torch.zeros(n_bins, dtype=conf.dtype, device=conf.device)
    count = zeros.scatter_add(0, idx, torch.ones_like(conf))
    conf_sum = zeros.scatter_add(0, idx, conf)
    acc_sum = zeros.scatter_add(0, idx, acc)
    safe = count.clamp(min=1)
    gap = (acc_sum / safe - conf_sum / safe).abs()
    return torch.sum(gap * count / count.sum())


def _group_means(values: torch.Tensor, cls: torch.Tensor, mask: torch.Tensor, C: int, empty: float) -> torch.Tensor:
    """Per-class mean of ``values[n]`` over the instances with ``mask[n]`` and class ``cls[n]`` -> [C] on the
    samples' device (one scatter-add, no per-class host loop); classes without such an instance get ``empty``."""
    w = mask.to(values.dtype)
    idx = cls.to(torch.int64)
    total = torch.zeros(C, dtype=values.dtype, device=values.device).index_add_(0, idx, torch.where(mask, values, 0.0))
    count = torch.zeros(C, dtype=values.dtype, device=values.device).index_add_(0, idx, w)
    return torch.where(count > 0, total / count.clamp(min=1), torch.full_like(total, empty))


def compute_mean_piws_for_class(prediction_tensors: Samples, mv: torch.Tensor, label: torch.Tensor):
    """Mean 2.5-97.5 % interval width of the predicted class, split by class and by correctness -> two [C] tensors.
    A class with no (in)correct prediction gets NaN, as the reference's mean over an empty selection does
    (classification_train_separately.py:102-140).  Unlike the reference nothing is copied to the host: the quantiles
    and the per-class means run on the device that holds the gathered samples."""
    s = as_sample_tensor(prediction_tensors).detach().float()
    mv, label = mv.detach().to(s.device), label.detach().to(s.device)
    q = torch.quantile(s, torch.tensor([0.025, 0.975], dtype=s.dtype, device=s.device), dim=0)   # [2, N, C]
    piw = (q[1] - q[0]).gather(1, mv.view(-1, 1).to(torch.int64)).squeeze(1)                      # predicted class
    C = s.shape[2]
    hit = mv == label
    return (_group_means(piw, mv, hit, C, float("nan")), _group_means(piw, mv, ~hit, C, float("nan")))


def calculate_variances(model_logits: Samples, predicted_classes: torch.Tensor, ground_truth: torch.Tensor):
    """Across-chain (unbiased) variance of the predicted class' output, averaged over the correct / incorrect
    instances of each class -> two [C] tensors, 0 where a class has none (classification_train_separately.py:143-174).
    Runs on the samples' device."""
    s = as_sample_tensor(model_logits).detach().float()
    pred, truth = predicted_classes.detach().to(s.device), ground_truth.detach().to(s.device)
    C = s.shape[2]
    own = s.gather(2, pred.view(1, -1, 1).expand(s.shape[0], -1, 1).to(torch.int64)).squeeze(2)   # [S, N]
    var = own.var(dim=0)                                                                          # NaN when S == 1
    hit = pred == truth
    return _group_means(var, pred, hit, C, 0.0), _group_means(var, pred, ~hit, C, 0.0)


def compute_accuracy(predictions: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """classification_train_separately.py:801-807 (a 0-dim tensor on the predictions' device)."""
    predictions = predictions.detach()
    labels = labels.detach().to(predictions.device)
    return torch.sum(predictions == labels).float() / predictions.numel()
