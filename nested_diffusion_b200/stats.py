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


def _masked_mean(values: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    return values[mask].mean()  # NaN for an empty selection, as in the reference


def compute_mean_piws_for_class(prediction_tensors: Samples, mv: torch.Tensor, label: torch.Tensor):
    """Mean 2.5-97.5 % interval width of the predicted class, split by class and by correctness."""
    s = as_sample_tensor(prediction_tensors).detach().float().cpu()
    mv, label = mv.detach().cpu(), label.detach().cpu()
    lo = torch.quantile(s, q=0.025, dim=0)
    hi = torch.quantile(s, q=0.975, dim=0)
    piw = (hi - lo)[torch.arange(s.shape[1]), mv]
    C = s.shape[2]
    correct, incorrect = torch.zeros(C), torch.zeros(C)
    for c in range(C):
        sel = mv == c
        correct[c] = _masked_mean(piw, sel & (mv == label))
        incorrect[c] = _masked_mean(piw, sel & (mv != label))
    return correct, incorrect


def calculate_variances(model_logits: Samples, predicted_classes: torch.Tensor, ground_truth: torch.Tensor):
    """Across-chain variance of the predicted class' output, averaged over correct / incorrect instances."""
    s = as_sample_tensor(model_logits).detach().float().cpu()
    pred, truth = predicted_classes.detach().cpu(), ground_truth.detach().cpu()
    C = s.shape[2]
    correct, incorrect = torch.zeros(C), torch.zeros(C)
    for c in range(C):
        ok = (pred == c) & (truth == c)
        ko = (pred == c) & (truth != c)
        if ok.any():
            correct[c] = s[:, ok, c].var(dim=0).mean()
        if ko.any():
            incorrect[c] = s[:, ko, c].var(dim=0).mean()
    return correct, incorrect


def compute_accuracy(predictions: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """classification_train_separately.py:801-807."""
    predictions, labels = predictions.detach().cpu(), labels.detach().cpu()
    return torch.sum(predictions == labels).float() / predictions.numel()
