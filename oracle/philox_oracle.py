"""NumPy oracle of the counter-based noise generator of the CUDA sampler (TEST INFRASTRUCTURE ONLY).

The reference draws its N(0,1) noise with ``torch.randn_like`` (diffusion_utils.py:67, :139), which a
different device cannot reproduce; the product instead keys Philox4x32-10 (Salmon et al., "Parallel
random numbers: as easy as 1, 2, 3", SC'11 -- the generator behind Random123/cuRAND) on the logical
chain identity so samples do not depend on the GPU partition.  This file restates that published
algorithm on uint32 arrays and the Box-Muller mapping of csrc/ladine_common.cuh::philox_normal.
Pinned by the Random123 known-answer vectors in tests/test_host_logic.py.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def u01(x: np.ndarray) -> np.ndarray:
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def normal(seed: int, chain: np.ndarray, slot: np.ndarray, c: np.ndarray) -> np.ndarray:
    """N(0,1) number `c` of noise slot `slot` of chain `chain` (broadcastable integer arrays), float64 math."""
    chain = np.asarray(chain, dtype=np.uint64)
    shape = np.broadcast(chain, slot, c).shape
    ctr = np.empty(shape + (4,), dtype=np.uint32)
    ctr[..., 0] = np.broadcast_to(chain & MASK, shape)
    ctr[..., 1] = np.broadcast_to(chain >> np.uint64(32), shape)
    ctr[..., 2] = np.broadcast_to(np.asarray(slot, dtype=np.uint32), shape)
    ctr[..., 3] = np.broadcast_to(np.asarray(c, dtype=np.uint32) >> np.uint32(2), shape)
    key = np.empty(shape + (2,), dtype=np.uint32)
    key[..., 0] = seed & 0xFFFFFFFF
    key[..., 1] = (seed >> 32) & 0xFFFFFFFF
    r = philox4x32_10(ctr, key)
    cc = np.broadcast_to(np.asarray(c), shape)
    pair = (cc >> 1) & 1
    u1 = np.take_along_axis(u01(r), (2 * pair)[..., None], axis=-1)[..., 0].astype(np.float64)
    u2 = np.take_along_axis(u01(r), (2 * pair + 1)[..., None], axis=-1)[..., 0].astype(np.float64)
    rad = np.sqrt(-2.0 * np.log(u1))
    return np.where(cc & 1, rad * np.sin(2 * np.pi * u2), rad * np.cos(2 * np.pi * u2))


def noise_tensor(seed, K, D, S, N, C, member_ids=None, image_offset=0, images_total=None, draw_offset=0,
                 draws_total=None) -> np.ndarray:
    """[K, D, S, N, C] float64: what ladine_fill_noise writes (chain id = ((member*Dtot)+draw)*Ntot+image)."""
    images_total = images_total or N
    draws_total = draws_total or D
    mid = np.arange(K) if member_ids is None else np.asarray(member_ids)
    k, d, s, n, c = np.meshgrid(mid, np.arange(D) + draw_offset, np.arange(S), np.arange(N) + image_offset,
                                np.arange(C), indexing="ij")
    chain = (k.astype(np.uint64) * np.uint64(draws_total) + d.astype(np.uint64)) * np.uint64(images_total) + n.astype(np.uint64)
    return normal(seed, chain, s, c)
