"""Drop-in evaluation loop for the nested ensemble: what ``Diffusion.test_atk`` and
``Diffusion.test_calibrate`` (classification_train_separately.py:631-840, :449-629) do around the sampler,
with the 100-call loop replaced by one batched ``NestedEnsemble.sample`` per test batch.

Kept from the reference: members paired one-to-one with the guidance outputs
(``target_pred[ii]`` with ``noise_estimators[ii]``, :773-775), ``y_T_mean = target_pred`` (:762),
``mc_trials`` draws per member (20, :770), the order of the K*D sample list (member-major, then trial),
majority vote with ties to the smallest label, ensemble confidence = mean of ``convert_to_prob``, and the
metrics and log lines printed at :810-838 / :620-627.

Changed on purpose: members are packed once and stay on the GPU (the reference moves each 2.59 GiB member
CPU->GPU->CPU per batch, :773/:780) and samples stay on the device until the statistics.  The guidance
provider (ViT blocks + mapping MLPs, :330-350) is any callable ``images -> list of K [N, C] logits``: it
stays PyTorch (north_star) and lives outside this package.
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch

from . import stats
from .ensemble import NestedEnsemble, sample_ensemble

# classification_train_separately.py:318-325
DATASET_TEMPERATURE = {"ChestXRay": 0.1737, "ISICSkinCancer": 0.3162}


def temperature_for(dataset: str) -> float:
    for prefix, t in DATASET_TEMPERATURE.items():
        if dataset.startswith(prefix):
            return t
    raise NotImplementedError(f"no scaling temperature defined for dataset {dataset!r}")


def _content_key(path: str, precision: str) -> str:
    """blake2b of the checkpoint file's CONTENT (not its name or mtime) + what the packed form depends on."""
    import hashlib

    from . import _capi

    hh = hashlib.blake2b(digest_size=16)
    with open(path, "rb") as f:
        while True:
            chunk = f.read(1 << 24)
            if not chunk:
                break
            hh.update(chunk)
    return f"{hh.hexdigest()}-{precision}-abi{_capi.load().ladine_version()}"


def _read_image(path: str) -> torch.Tensor:
    n = os.path.getsize(path)
    buf = torch.empty(n, dtype=torch.uint8)
    with open(path, "rb") as f:
        got = f.readinto(memoryview(buf.numpy()))
    if got != n:
        raise OSError(f"short read of {path}")
    return buf


def _write_image(path: str, image: torch.Tensor) -> None:
    tmp = f"{path}.tmp{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(memoryview(image.numpy()))
    os.replace(tmp, path)   # atomic: a concurrent reader sees the old file or the complete new one


def load_noise_estimators(config, ckpt_paths: Sequence[str], device, guidance: Optional[bool] = None,
                          cache_dir: Optional[str] = None, precision: str = "auto") -> List:
    """Build one ConditionalModel per checkpoint and load ``state['noise_estimator']`` into it
    (classification_train_separately.py:684-697), directly on ``device`` and in eval mode.

    ``cache_dir``: keep the PACKED form of every checkpoint on disk, keyed by a hash of the checkpoint
    file's content, the packing precision and the library's ABI version.  A hit restores the packed trunk and the
    packed encoder straight into device memory (``ladine_member_import`` / ``ladine_encoder_import``) and skips building
    the 2.59 GiB module, ``load_state_dict`` and the packing kernels; a miss packs as usual and writes the two images.
    With a cache the members are returned as ``engine.PackedModel`` objects (packed form only -- the FP32 module is
    released), which ``NestedEnsemble`` / ``NestedDiffusionTester`` accept like modules.  A stale, truncated or corrupt
    image is ignored and rebuilt from the checkpoint.  Encoders the kernel does not cover are not cached (the module
    is returned)."""
    from . import engine
    from ._capi import LadineError
    from .latent_model import ConditionalModel

    if guidance is None:
        guidance = bool(config.diffusion.include_guidance)
    if cache_dir is not None:
        os.makedirs(cache_dir, exist_ok=True)
    members = []
    for path in ckpt_paths:
        files = None
        if cache_dir is not None:
            key = _content_key(path, precision)
            files = [os.path.join(cache_dir, f"{key}.{kind}.ladine") for kind in ("member", "encoder")]
            if all(os.path.exists(f) for f in files):
                try:
                    pm = engine.PackedMember.from_image(_read_image(files[0]), device)
                    pe = engine.PackedEncoder.from_image(_read_image(files[1]), device)
                    if pm.guidance != bool(guidance):
                        raise ValueError("cached member was packed with a different guidance setting")
                    members.append(engine.PackedModel(pm, pe))
                    continue
                except (LadineError, OSError, ValueError) as e:
                    logging.warning(f"packed cache entry for {path} ignored ({e}); re-packing from the checkpoint")
        state = torch.load(path, map_location="cpu")
        sd = state["noise_estimator"] if isinstance(state, dict) and "noise_estimator" in state else state
        m = ConditionalModel(config, guidance=guidance)
        m.load_state_dict(sd)
        m = m.to(device).eval()
        if files is not None:
            pe = engine.packed_encoder_of(m)
            if pe is not None:
                pm = engine.packed_member_of(m, precision)
                _write_image(files[0], pm.export_image())
                _write_image(files[1], pe.export_image())
                members.append(engine.PackedModel(pm, pe))
                continue
        members.append(m)
    return members


@dataclass
class SampleCache:
    """Samples of one pass over a loader, reusable across temperatures: the posterior samples do not depend
    on the softmax temperature (``convert_to_prob`` is applied afterwards), so a Nelder-Mead search over it
    (main.py:356-358) needs to sample only once.  Opt-in -- the reference re-samples every evaluation."""
    y0: List[torch.Tensor] = field(default_factory=list)       # per batch [S, N, C]
    target: List[torch.Tensor] = field(default_factory=list)   # per batch [N]


def ensemble_metrics(cache: SampleCache, temperature: float) -> Dict[str, torch.Tensor]:
    """Accuracy / ECE / PIW / variances over the cached samples, as computed at
    classification_train_separately.py:786-815 (majority vote and ensemble confidence per batch, then the
    dataset-level metrics on the concatenation)."""
    mv = torch.cat([stats.majority_voting_for_mc_samples(s) for s in cache.y0])
    prob = torch.cat([stats.compute_ensemble_confidence(s, temperature) for s in cache.y0])
    target = torch.cat(cache.target)
    pred_mc = torch.cat(cache.y0, dim=1)  # [S, N_total, C]  (:791-794)
    piw_ok, piw_ko = stats.compute_mean_piws_for_class(pred_mc, mv, target)
    var_ok, var_ko = stats.calculate_variances(pred_mc, mv, target)
    out = {"accuracy": stats.compute_accuracy(mv, target), "ece": stats.compute_ece(prob, target),
           "piw_correct": piw_ok, "piw_incorrect": piw_ko, "var_correct": var_ok, "var_incorrect": var_ko,
           "majority_vote": mv, "ensemble_prob": prob}
    # every statistic above ran on the device holding the gathered samples; only the (small) results of the report
    # cross to the host, once (the reference moves all K*D sample tensors, classification_train_separately.py:783-784)
    return {k: v.cpu() for k, v in out.items()}


class NestedDiffusionTester:
    def __init__(self, members: Sequence, guidance_fn: Callable[[torch.Tensor], Sequence[torch.Tensor]],
                 num_timesteps: int, alphas: torch.Tensor, one_minus_alphas_bar_sqrt: torch.Tensor,
                 temperature: float, mc_trials: int = 20, selected_block_indices: Optional[Sequence[int]] = None,
                 precision: str = "auto", flatten_images: bool = True, seed: Optional[int] = None):
        sel = list(range(len(members))) if selected_block_indices is None else list(selected_block_indices)
        self.selected = [i for i in sel if i < len(members)]
        self.ensemble = NestedEnsemble([members[i] for i in self.selected], precision=precision,
                                       member_ids=self.selected)
        self.guidance_fn = guidance_fn
        self.num_timesteps = int(num_timesteps)
        self.alphas, self.omabs = alphas, one_minus_alphas_bar_sqrt
        self.temperature = float(temperature)
        self.mc_trials = int(mc_trials)
        self.flatten_images = flatten_images
        self.seed = seed
        self.device = self.ensemble.device
        self._batches = 0

    # -- one test batch: the body of the loop at classification_train_separately.py:715-794 --------------
    def sample_batch(self, images: torch.Tensor) -> torch.Tensor:
        """-> ``[K*D, N, C]`` posterior samples in the reference's list order (member-major, then trial)."""
        images = images.to(self.device)
        with torch.no_grad():
            logits = self.guidance_fn(images)
            target_pred = torch.stack([torch.softmax(logits[i], dim=1) for i in self.selected])  # :753-758
            x = torch.flatten(images, 1) if self.flatten_images else images                        # :747
            seed = None if self.seed is None else self.seed + self._batches
            self._batches += 1
            y0, _ = sample_ensemble(self.ensemble, x, target_pred, self.mc_trials, self.num_timesteps, self.alphas,
                                    self.omabs, seed=seed)
        return y0.permute(1, 0, 2).contiguous()  # [N, K*D, C] -> [K*D, N, C]

    def collect(self, loader: Iterable) -> SampleCache:
        cache = SampleCache()
        for images, target in loader:
            cache.y0.append(self.sample_batch(images))
            cache.target.append(torch.as_tensor(target).to(self.device))
        return cache

    def metrics(self, cache: SampleCache, temperature: Optional[float] = None) -> Dict[str, torch.Tensor]:
        return ensemble_metrics(cache, self.temperature if temperature is None else float(temperature))

    def test_atk(self, loader: Iterable, cache: Optional[SampleCache] = None) -> torch.Tensor:
        """Majority-vote accuracy of the nested ensemble; prints/logs the reference's report (:820-838)."""
        cache = self.collect(loader) if cache is None else cache
        m = self.metrics(cache)
        report = (f"Majority voting accuracy for MC: {m['accuracy'] :.4f} \n" +
                  f"ECE: {m['ece'] :.4f} \n" +
                  f"Average correct PIW per class: {m['piw_correct']} \n" +
                  f"Average incorrect PIW per class: {m['piw_incorrect']} \n" +
                  f"Average correct variances per class: {m['var_correct']} \n" +
                  f"Average incorrect variances per class: {m['var_incorrect']}")
        print(report)
        logging.info(report + " \n")
        self.last_metrics = m
        return m["accuracy"]

    def test_calibrate(self, loader: Iterable, temp: Optional[float] = None,
                       cache: Optional[SampleCache] = None) -> torch.Tensor:
        """ECE at scaling temperature ``temp`` (objective of the Nelder-Mead search, main.py:356-358).
        Pass a ``SampleCache`` from ``collect`` to evaluate many temperatures on one set of samples."""
        if temp is not None:
            self.temperature = float(temp if not hasattr(temp, "__len__") else temp[0])
        cache = self.collect(loader) if cache is None else cache
        target = torch.cat(cache.target)
        prob = torch.cat([stats.compute_ensemble_confidence(s, self.temperature) for s in cache.y0])
        ece = stats.compute_ece(prob, target).cpu()
        print(f"Ours ECE: {ece} \n")
        logging.info(f"Ours ECE: {ece} \n")
        return ece
