"""Drop-in for the reference's ``latent_model`` module, restricted to the sampler's hot path.

``ConditionalLinear`` and ``ConditionalModel`` keep the reference's constructor signatures, forward
semantics and -- the checkpoint contract -- ``state_dict()`` key layout (latent_model.py:93-184;
SURVEY.md §8 a10), so ``load_state_dict(torch.load(path)['noise_estimator'])``
(classification_train_separately.py:689-690) works unchanged.  Only the encoder families the
shipped configs select are built ('linear', plus 'simple' and the toy Linear); the convolutional
encoders (FashionCNN / LeNet / ResNetEncoder, latent_model.py:216-368) are out of scope.

``forward`` is the plain PyTorch evaluation (used for training and as the input provider of the
step-invariant features); the accelerated reverse process lives in ``diffusion_utils.p_sample_loop``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

_IMAGE_DATASETS = ("FashionMNIST", "MNIST", "CIFAR10", "CIFAR100", "IMAGENE100", "RotatedMNIST")


def _is_image_dataset(name: str) -> bool:
    return name in _IMAGE_DATASETS or name.startswith("ChestXRay") or name.startswith("ISICSkinCancer")


class ConditionalLinear(nn.Module):
    """Linear layer gated by a learned per-timestep vector: ``embed(t) * lin(x)`` (latent_model.py:93-105)."""

    def __init__(self, num_in: int, num_out: int, n_steps: int):
        super().__init__()
        self.num_out = num_out
        self.lin = nn.Linear(num_in, num_out)
        self.embed = nn.Embedding(n_steps, num_out)
        nn.init.uniform_(self.embed.weight)  # U(0,1), latent_model.py:99

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        return self.embed(t).view(-1, self.num_out) * self.lin(x)


def _mlp_encoder(data_dim: int, widths, feature_dim: int, act) -> nn.Sequential:
    layers, prev = [], data_dim
    for w in widths:
        layers += [nn.Linear(prev, w), nn.BatchNorm1d(w), act()]
        prev = w
    layers.append(nn.Linear(prev, feature_dim))
    return nn.Sequential(*layers)  # indices 0,1,2,3,4,5,6 as in the reference


class ConditionalModel(nn.Module):
    """eps_theta(x, y_t, t, y_0_hat) -- latent_model.py:108-184."""

    def __init__(self, config, guidance: bool = False):
        super().__init__()
        n_steps = config.diffusion.timesteps + 1
        y_dim = config.data.num_classes
        feature_dim = config.model.feature_dim
        arch = config.model.arch
        self.guidance = guidance
        if config.data.dataset == "toy":
            self.encoder_x = nn.Linear(config.model.data_dim, feature_dim)
        elif _is_image_dataset(config.data.dataset) and arch == "linear":
            h = config.model.hidden_dim
            self.encoder_x = _mlp_encoder(config.model.data_dim, (h, h), feature_dim, nn.Softplus)
        elif _is_image_dataset(config.data.dataset) and arch == "simple":
            self.encoder_x = _mlp_encoder(config.model.data_dim, (300, 100), feature_dim, nn.ReLU)
        else:
            raise NotImplementedError(
                f"encoder arch {arch!r} for dataset {config.data.dataset!r} is outside the accelerated hot path "
                "(shipped configs use arch 'linear'); see DESIGN.md 'out of scope'")
        self.norm = nn.BatchNorm1d(feature_dim)
        self.lin1 = ConditionalLinear(2 * y_dim if guidance else y_dim, feature_dim, n_steps)
        self.unetnorm1 = nn.BatchNorm1d(feature_dim)
        self.lin2 = ConditionalLinear(feature_dim, feature_dim, n_steps)
        self.unetnorm2 = nn.BatchNorm1d(feature_dim)
        self.lin3 = ConditionalLinear(feature_dim, feature_dim, n_steps)
        self.unetnorm3 = nn.BatchNorm1d(feature_dim)
        self.lin4 = nn.Linear(feature_dim, y_dim)

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """Step-invariant image features ``norm(encoder_x(x))`` (latent_model.py:170-171)."""
        return self.norm(self.encoder_x(x))

    def forward(self, x, y, t, yhat=None):
        xf = self.encode(x)
        h = torch.cat([y, yhat], dim=-1) if self.guidance else y
        h = F.softplus(self.unetnorm1(self.lin1(h, t)))
        h = xf * h
        h = F.softplus(self.unetnorm2(self.lin2(h, t)))
        h = F.softplus(self.unetnorm3(self.lin3(h, t)))
        return self.lin4(h)
