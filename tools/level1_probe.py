"""Module-swap (INTEGRATION.md level 1) latency: the reference runner's own loop -- for every member, 20 sequential
p_sample_loop calls on the 70-image test batch (classification_train_separately.py:764-784) -- through the drop-in,
next to the batched NestedEnsemble call (level 2).   python tools/level1_probe.py [draws_per_member]"""
import argparse
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402  (member construction at the shipped shape)
import nested_diffusion_b200 as nd  # noqa: E402
from nested_diffusion_b200 import diffusion_utils as du  # noqa: E402
from nested_diffusion_b200.schedule import make_beta_schedule, schedule_tensors  # noqa: E402

draws = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda", 0)
models = bench.build_members(dev)
K, N, T = len(models), bench.N_IMAGES_C2, bench.T_STEPS   # the ChestXRay test batch: 70 images
alphas, omabs = schedule_tensors(make_beta_schedule("linear", T, 1e-4, 0.02))
alphas, omabs = alphas.to(dev), omabs.to(dev)
g = torch.Generator().manual_seed(1)
x = torch.rand(N, bench.DX, generator=g).to(dev)
yh = [torch.softmax(2 * torch.randn(N, bench.N_CLASSES, generator=g), -1).to(dev) for _ in range(K)]

def level1():
    outs = []
    with torch.no_grad():
        for ii in range(K):
            for _ in range(draws):
                outs.append(du.p_sample_loop(models[ii], x, yh[ii], yh[ii], T, alphas, omabs, only_last_sample=True))
    return torch.stack(outs)

def level2(ens):
    with torch.no_grad():
        return ens.sample(x, yh, draws, T, alphas, omabs, temperature=bench.TEMPERATURE_C2).y0

def level1_ahead():
    du.set_draws_ahead(draws)
    try:
        return level1()
    finally:
        du.set_draws_ahead(0)

for name, fn in (("level 1: K x draws sequential p_sample_loop calls", level1),
                 ("level 1 + draws-ahead (LADINE_DRAWS_AHEAD): the same calls served from one batched launch per member", level1_ahead),
                 ("level 2: one NestedEnsemble.sample call", None)):
    if fn is None:
        ens = nd.NestedEnsemble(models)
        fn = lambda: level2(ens)
    fn(); torch.cuda.synchronize()          # warm-up (packs the members)
    t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n = K * draws * N
    print(f"{name}: {dt:.3f} s per {N}-image batch ({n} chains, {n / dt:.0f} samples/s), finite={bool(torch.isfinite(out).all())}")
