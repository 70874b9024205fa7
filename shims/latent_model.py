"""Module-swap shim (INTEGRATION.md level 1): `from latent_model import ConditionalModel`
(classification_train_separately.py:22) binds the drop-in module tree (same constructor, same state_dict keys)."""
from nested_diffusion_b200.latent_model import ConditionalLinear, ConditionalModel  # noqa: F401
