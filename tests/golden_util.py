"""Load tests/golden/*.npz fixtures (written by tests/golden/make_golden.py from the reference)."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np
import torch

from oracle import ladine_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().numpy()).tobytes())
    return h.hexdigest()[:16]


def names(kind=None):
    out = []
    for f in sorted(os.listdir(GOLDEN)):
        if f.endswith(".npz"):
            meta = json.loads(str(np.load(os.path.join(GOLDEN, f))["meta"]))
            if kind is None or meta["kind"] == kind:
                out.append(f[:-4])
    return out


class Fixture:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.meta = json.loads(str(z["meta"]))
        self.arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}

    def __getitem__(self, k):
        return self.arrays[k]

    def schedule(self):
        m = self.meta
        betas = orc.make_beta_schedule(m.get("sched", "linear"), m["T"], 1e-4, 0.02)
        return orc.schedule_tensors(betas, m.get("sched", "linear"))


class ChainFixture(Fixture):
    """kind == 'chain': one member, one p_sample_loop call."""

    def materialize(self, check=True):
        m = self.meta
        if m["stored_inputs"]:
            sd = {k[3:]: v for k, v in self.arrays.items() if k.startswith("sd/")}
            if m.get("sd_overlay"):
                # the big tensors (W2, W3, gamma tables) are the seeded ones; the trained small ones are stored
                base = orc.synth_state_dict(m["sd_seed"], m["F"], m["H"], m["Dx"], m["C"], m["T"])
                base.update(sd)
                sd = base
                if check:
                    assert _digest(*[sd[k].float() for k in sorted(sd)]) == m["sd_digest"], "member drifted"
            x, yhat, noise = self["x"], self["yhat"], self["noise"]
        else:
            sd = orc.synth_state_dict(m["sd_seed"], m["F"], m["H"], m["Dx"], m["C"], m["T"],
                                      guidance=m["guidance"], eps_gain=m["eps_gain"])
            x, yhat = orc.synth_inputs(m["in_seed"], m["B"], m["Dx"], m["C"])
            torch.manual_seed(m["noise_seed"])
            noise = torch.stack([torch.randn(m["B"], m["C"]) for _ in range(m["T"])])
            if check:
                assert _digest(x, yhat) == m["in_digest"], "synthetic inputs drifted from the golden run"
                assert _digest(*[sd[k].float() for k in sorted(sd)]) == m["sd_digest"], "member drifted"
        alphas, omabs = self.schedule()
        return sd, x, yhat, noise, alphas, omabs


class EnsembleFixture(Fixture):
    """kind == 'ensemble': K members x D draws, noise in the reference's call order."""

    def materialize(self):
        m = self.meta
        sds = [orc.synth_state_dict(m["sd_seed0"] + k, m["F"], m["H"], m["Dx"], m["C"], m["T"])
               for k in range(m["K"])]
        g = torch.Generator().manual_seed(m["in_seed"])
        x = torch.rand(m["N"], m["Dx"], generator=g)
        y0hats = [torch.softmax(2 * torch.randn(m["N"], m["C"], generator=g), dim=1) for _ in range(m["K"])]
        assert _digest(x, *y0hats) == m["in_digest"]
        torch.manual_seed(m["noise_seed"])
        noise = torch.stack([torch.randn(m["N"], m["C"]) for _ in range(m["K"] * m["D"] * m["T"])])
        noise = noise.reshape(m["K"], m["D"], m["T"], m["N"], m["C"])
        alphas, omabs = self.schedule()
        return sds, x, y0hats, noise, alphas, omabs


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max-abs error relative to max(1, ||b||_inf) -- the SURVEY.md §8d tolerance convention."""
    return float((a.double() - b.double()).abs().max() / max(1.0, float(b.abs().max())))
