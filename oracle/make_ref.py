"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference modules of the hot path, staged for the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY (same rule as the rest of ``oracle/``: only tests, smoke() and bench.py's CPU legs
may use it, never the product path).

The reference is pure Python (no native code), so "compiling" it means nothing more than placing the
two modules the hot path lives in -- ``diffusion/diffusion_utils.py`` (p_sample_loop / p_sample / p_sample_t_1to0 /
schedules) and ``diffusion/latent_model.py`` (ConditionalModel / ConditionalLinear) -- where the GPU box can import
them: ``/root/reference`` does not exist there, ``oracle/_ref/`` travels with the gpurun snapshot (it is listed in
.gitignore, so the reference's sources never enter this repository's history, and not in .gpurunignore).
``__graft_entry__.build()`` runs this in the build container; ``bench.py --impl reference`` and ``cpu_baseline`` then
time the real reference (``kind: "reference"``) instead of the oracle port.  Files are copied byte for byte.
"""
from __future__ import annotations

import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/diffusion"
REF_DST = os.path.join(HERE, "_ref")
MODULES = ("diffusion_utils.py", "latent_model.py")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DST, m)) for m in MODULES)


def make(verbose: bool = False) -> bool:
    """Stage the reference modules; returns False (and leaves any earlier copy alone) when /root/reference is absent."""
    if not all(os.path.exists(os.path.join(REF_SRC, m)) for m in MODULES):
        return available()
    os.makedirs(REF_DST, exist_ok=True)
    lines = []
    for m in MODULES:
        shutil.copyfile(os.path.join(REF_SRC, m), os.path.join(REF_DST, m))
        with open(os.path.join(REF_DST, m), "rb") as f:
            lines.append(f"{hashlib.sha256(f.read()).hexdigest()}  {m}")
    with open(os.path.join(REF_DST, "SHA256SUMS"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if verbose:
        print("staged oracle/_ref:", ", ".join(MODULES))
    return True


def import_reference():
    """(diffusion_utils, latent_model) of the staged reference, or None when oracle/_ref is absent."""
    if not available():
        return None
    import importlib
    import sys

    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)   # latent_model does ``from diffusion_utils import ...``
    du = importlib.import_module("diffusion_utils")
    lm = importlib.import_module("latent_model")
    if os.path.dirname(os.path.abspath(du.__file__)) != REF_DST:
        raise RuntimeError(f"'diffusion_utils' resolved to {du.__file__}, not the staged reference")
    return du, lm


if __name__ == "__main__":
    print("ok" if make(verbose=True) else "reference not present; nothing staged")
