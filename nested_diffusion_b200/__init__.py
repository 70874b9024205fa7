"""B200-native nested-ensemble reverse-diffusion sampler (drop-in for LaDiNE's hot path).

Public surface mirrors the reference modules the runner imports:
  nested_diffusion_b200.diffusion_utils  <->  diffusion/diffusion_utils.py
  nested_diffusion_b200.latent_model     <->  diffusion/latent_model.py (ConditionalLinear/ConditionalModel)
plus the batched entry points the 100-call loop collapses into (ensemble.NestedEnsemble,
ensemble.sample_ensemble) and the ensemble statistics (stats).
"""
from . import diffusion_utils, latent_model, schedule, stats  # noqa: F401
from .engine import (PackedEncoder, PackedMember, PackedModel, fill_noise, packed_member_of,  # noqa: F401
                     sample_chains)
from .ensemble import (NestedEnsemble, gather_image_shards, sample_ensemble, shard_bounds,  # noqa: F401
                       weighted_bounds)
from .latent_model import ConditionalLinear, ConditionalModel  # noqa: F401

__version__ = "0.1.0"
