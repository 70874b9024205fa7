"""Nested-ensemble driver: every member x posterior draw x image in one batched call, sharded over
the GPUs of one node by image tile, with a single all-gather of the per-draw class probabilities.

Replaces the loop at classification_train_separately.py:764-784 (K members x 20 sequential
``p_sample_loop`` calls on the same images, members shuttled CPU<->GPU, samples moved to the CPU one
by one).  Chains are independent, so sharding needs no data-path collective; the only exchange is
the gather that feeds the ensemble statistics.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import engine
from .schedule import coef_table


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of ``range(n_items)``: the first ``n_items % world`` ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def padded_shard_size(n_items: int, world: int) -> int:
    return -(-n_items // world)


def weighted_bounds(n_items: int, weights: Sequence[float]) -> List[Tuple[int, int]]:
    """Contiguous partition of ``range(n_items)`` into ``len(weights)`` shards whose sizes follow ``weights`` (the
    measured relative speed of each rank: the GPUs of one box differ by several per cent under the 1 kW power cap, and a
    step of the sharded sampler ends when the SLOWEST rank is done).  Every rank must pass the same weights.  Because
    the Philox streams are keyed on global (member, draw, image) ids, the partition does not change a single sample."""
    w = [max(float(x), 0.0) for x in weights]
    if not w or sum(w) <= 0:
        raise ValueError("weights must contain a positive entry")
    total, acc, edges = sum(w), 0.0, [0]
    for x in w[:-1]:
        acc += x
        edges.append(min(n_items, max(edges[-1], int(round(n_items * acc / total)))))
    edges.append(n_items)
    return [(edges[i], edges[i + 1]) for i in range(len(w))]


def image_tiles(n_images: int, draws: int, max_rows: int) -> List[Tuple[int, int]]:
    """Split ``range(n_images)`` into the fewest contiguous tiles of at most ``max_rows // draws`` images (at least one
    image per tile), all of (nearly) equal size so that no small remainder call is left; an empty batch is one empty
    tile (the call still has to produce its empty outputs)."""
    if n_images <= 0:
        return [(0, 0)]
    cap = max(1, min(n_images, int(max_rows) // max(1, int(draws))))
    n_tiles = -(-n_images // cap)
    tile = -(-n_images // n_tiles)
    return [(lo, min(n_images, lo + tile)) for lo in range(0, n_images, tile)]


def gather_image_shards(local: torch.Tensor, n_items: int, group=None,
                        bounds: Optional[Sequence[Tuple[int, int]]] = None) -> torch.Tensor:
    """All-gather per-rank tensors ``[n_local, ...]`` (image-major) into ``[n_items, ...]`` on every rank.

    Equal-count collective: shards are zero-padded to the largest shard (``ceil(n_items / world)`` rows for the default
    balanced partition; ``bounds`` gives the per-rank ``(lo, hi)`` of a weighted one), gathered with one
    ``all_gather_into_tensor`` (NCCL over NVLink on the B200 box, gloo in the CPU tests) and trimmed."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != n_items:
            raise ValueError("no process group: the local shard must hold every image")
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if bounds is None:
        bounds = [shard_bounds(n_items, r, world) for r in range(world)]
    if len(bounds) != world:
        raise ValueError("need one (lo, hi) per rank")
    lo, hi = bounds[rank]
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} should hold {hi - lo} images, got {local.shape[0]}")
    per = max(1, max(b - a for a, b in bounds))
    send = local.new_zeros((per,) + tuple(local.shape[1:]))
    send[: hi - lo] = local
    recv = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    parts = [recv[r * per: r * per + (bounds[r][1] - bounds[r][0])] for r in range(world)]
    return torch.cat(parts, dim=0)


@dataclass
class EnsembleResult:
    y0: torch.Tensor                 # [K, D, N_local, C] final y_0 of every chain of this rank's images
    probs: Optional[torch.Tensor]    # [K, D, N_local, C] softmax(-(y_0-1)^2 / temperature), if asked
    image_range: Tuple[int, int]     # global [lo, hi) of the images held by this rank


class NestedEnsemble:
    """K packed members kept resident on one GPU (no per-batch CPU<->GPU shuttle)."""

    # Chains (images x draws) per member per ladine_sample call.  Bounds the activation workspace (2 x rows x F 16-bit per
    # member) AND keeps the GEMM launches short enough for the L2: the CTAs of a launch walk a static tile schedule at
    # their own pace, and over a very long launch they drift apart until the 16 column tiles that share an activation row
    # tile (and the row tiles that share a weight tile) no longer meet in L2 -- measured DRAM reads per row of a layer-2
    # launch: 19 KB at 20 480 rows per member, 37 KB at 163 840, 52 KB at 256 000 (8 KB algorithmic), and under the 1 kW
    # cap that traffic costs clock: 27.5 -> 29.8 -> 32.7 ns per row.  Measured optimum (profiles/README.md): 32 768 --
    # the 5.12 M-chain sweep point runs at 18.4 k samples/s with it against 15.8-16.1 k at the round-1 cap of 262 144;
    # config 3 (20 480 rows per member) stays one call, and smaller caps lose again (more, shorter launches).
    MAX_ROWS_PER_CALL = 32768

    def __init__(self, models: Sequence, precision: str = "auto", member_ids: Optional[Sequence[int]] = None,
                 max_rows_per_call: Optional[int] = None):
        self.max_rows_per_call = int(max_rows_per_call if max_rows_per_call is not None else self.MAX_ROWS_PER_CALL)
        if len(models) < 1:
            raise ValueError("need at least one member")
        self.models = list(models)
        self.precision = precision
        self.members: List[engine.PackedMember] = [engine.packed_member_of(m, precision) for m in models]
        self.member_ids = list(member_ids) if member_ids is not None else list(range(len(models)))
        self.device = self.members[0].device

    @property
    def K(self) -> int:
        return len(self.members)

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """[K, N, F] step-invariant features: one ladine_encode call for all members (PyTorch modules for encoder
        archs the kernel does not cover)."""
        return engine.encode_members(self.models, x)

    def sample(self, x: torch.Tensor, y0hats, draws: int, n_steps: int, alphas, one_minus_alphas_bar_sqrt, *,
               y_T_means=None, noise: Optional[torch.Tensor] = None, seed: Optional[int] = None,
               temperature: Optional[float] = None, xf: Optional[torch.Tensor] = None, image_offset: int = 0,
               images_total: int = 0, draw_offset: int = 0, draws_total: int = 0) -> EnsembleResult:
        """All K x ``draws`` chains for the images in ``x`` ([N, ...]); ``y0hats``: [K, N, C] or list of K [N, C]."""
        # members fine-tuned / reloaded in place since the last call are re-packed (a cheap fingerprint hit otherwise);
        # train-mode BatchNorm (batch statistics) is not what the folded tables compute
        for m in self.models:
            if getattr(m, "training", False):
                raise NotImplementedError("NestedEnsemble.sample needs members in eval() mode (train-mode BatchNorm uses "
                                          "batch statistics; the kernels fold the running statistics)")
        if self.models:
            self.members = [engine.packed_member_of(m, self.precision) for m in self.models]
        y0hats = torch.stack(list(y0hats)) if not torch.is_tensor(y0hats) else y0hats
        mus = y0hats if y_T_means is None else (
            torch.stack(list(y_T_means)) if not torch.is_tensor(y_T_means) else y_T_means)
        if xf is None:
            xf = self.encode(x)
        if noise is None and seed is None:
            seed = engine.fresh_seed()
        coef = coef_table(alphas, one_minus_alphas_bar_sqrt, n_steps)
        n = xf.shape[1]
        total = images_total if images_total else n
        # bound the rows per call (workspace and L2 locality, see MAX_ROWS_PER_CALL): process image tiles in turn;
        # Philox ids are global, so tiling does not change a single sample
        ys, ps = [], []
        for lo, hi in image_tiles(n, draws, self.max_rows_per_call):
            out = engine.sample_chains(
                self.members, xf[:, lo:hi], y0hats[:, lo:hi], mus[:, lo:hi], coef, draws,
                noise=None if noise is None else noise[:, :, :, lo:hi], seed=seed or 0, member_ids=self.member_ids,
                image_offset=image_offset + lo, images_total=max(total, image_offset + n), draw_offset=draw_offset,
                draws_total=draws_total, temperature=temperature)
            ys.append(out["y"])
            ps.append(out.get("probs"))
        y = ys[0] if len(ys) == 1 else torch.cat(ys, dim=2)
        p = None if ps[0] is None else (ps[0] if len(ps) == 1 else torch.cat(ps, dim=2))
        return EnsembleResult(y, p, (image_offset, image_offset + n))


def sample_ensemble(models_or_ensemble, x, y0hats, draws, n_steps, alphas, one_minus_alphas_bar_sqrt, *,
                    seed: Optional[int] = None, temperature: Optional[float] = None, group=None,
                    precision: str = "auto", shard_weights: Optional[Sequence[float]] = None,
                    local_events: Optional[list] = None):
    """Sharded nested-ensemble sampling + the single all-gather.

    Every rank passes the SAME full ``x`` [N, ...] and ``y0hats`` [K, N, C]; each samples its image
    tile (``shard_bounds``) for all members and draws with Philox streams keyed on global
    (member, draw, image) ids -- so the gathered result is identical for any world size -- and
    returns on every rank ``(y0 [N, K*D, C], probs [N, K*D, C] or None)`` in image-major order.
    ``x`` / ``y0hats`` may live on the HOST (pinned memory makes the copy asynchronous): only this rank's
    image tile is copied to the device.  ``shard_weights`` (one entry per rank, identical on all ranks) sizes the
    image tiles by measured rank speed instead of equally (``weighted_bounds``); the result is the same either way.
    ``local_events``: a list that receives one ``(start, end)`` pair of CUDA events around this rank's own sampling
    (before the gather), for load-balance measurements."""
    import torch.distributed as dist

    ens = models_or_ensemble if isinstance(models_or_ensemble, NestedEnsemble) else NestedEnsemble(
        models_or_ensemble, precision)
    y0hats = torch.stack(list(y0hats)) if not torch.is_tensor(y0hats) else y0hats
    N = x.shape[0]
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    if seed is None:
        seed = engine.fresh_seed()
        if distributed:  # all ranks must agree on the key
            box = [seed]
            dist.broadcast_object_list(box, src=0, group=group)
            seed = box[0]
    if shard_weights is not None and world > 1:
        if len(shard_weights) != world:
            raise ValueError("shard_weights needs one entry per rank")
        bounds = weighted_bounds(N, shard_weights)
    else:
        bounds = [shard_bounds(N, r, world) for r in range(world)]
    lo, hi = bounds[rank]
    K, D = ens.K, int(draws)
    if local_events is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    C = y0hats.shape[-1]
    if hi > lo:
        x_local, yh_local = x[lo:hi], y0hats[:, lo:hi]
        if x_local.device != ens.device:
            x_local = x_local.to(ens.device, non_blocking=True)
        if yh_local.device != ens.device:
            yh_local = yh_local.to(ens.device, non_blocking=True)
        res = ens.sample(x_local, yh_local, D, n_steps, alphas, one_minus_alphas_bar_sqrt, seed=seed,
                         temperature=temperature, image_offset=lo, images_total=N)
        y_local = res.y0.permute(2, 0, 1, 3).reshape(hi - lo, K * D, C)
        p_local = res.probs.permute(2, 0, 1, 3).reshape(hi - lo, K * D, C) if res.probs is not None else None
    else:
        y_local = torch.empty((0, K * D, C), dtype=torch.float32, device=ens.device)
        p_local = torch.empty_like(y_local) if temperature is not None else None
    if local_events is not None:
        ev[1].record()
        local_events.append(ev)
    if p_local is not None:
        both = gather_image_shards(torch.cat([y_local, p_local], dim=-1).contiguous(), N, group, bounds)
        return both[..., :C].contiguous(), both[..., C:].contiguous()
    return gather_image_shards(y_local.contiguous(), N, group, bounds), None
