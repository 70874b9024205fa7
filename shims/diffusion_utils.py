"""Module-swap shim (INTEGRATION.md level 1): put this directory ahead of the reference's `diffusion/` on sys.path
and `from diffusion_utils import *` (classification_train_separately.py:20) binds the fused B200 sampler."""
from nested_diffusion_b200.diffusion_utils import *  # noqa: F401,F403
from nested_diffusion_b200.diffusion_utils import __all__  # noqa: F401
