"""Small-call timing / profiling target: the reference's own call shape (one member, B images, one draw, T = 1000,
F = 4096) through sample_chains, persistent (one cooperative launch) vs the three-launches-per-step tile kernels.

    python tools/profile_small.py [rows] [precision]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import nested_diffusion_b200 as nd  # noqa: E402
from nested_diffusion_b200 import engine  # noqa: E402
from nested_diffusion_b200.schedule import coef_table, make_beta_schedule, schedule_tensors  # noqa: E402
from tests.test_gpu_parity import _rand_trunk_sd  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
F, C, T = 4096, 2, 1000
dev = torch.device("cuda")
import os  # noqa: E402
if os.environ.get("PERSIST_DEBUG"):
    engine.set_option(0, "persist_debug", 1)
for K in ((1,) if os.environ.get("PERSIST_DEBUG") else (1, 2, 4)):
    pms = [nd.PackedMember(_rand_trunk_sd(2000 + k, F, C, T, dev), n_steps=T, precision=prec) for k in range(K)]
    g = torch.Generator(device="cuda").manual_seed(1)
    xf = torch.randn(K, rows, F, device=dev, generator=g)
    yh = torch.softmax(torch.randn(K, rows, C, device=dev, generator=g), -1)
    alphas, omabs = schedule_tensors(make_beta_schedule("linear", T, 1e-4, 0.02))
    coef = coef_table(alphas, omabs, T)
    for pers in (True, False):
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = engine.sample_chains(pms, xf, yh, yh, coef, 1, seed=3, persistent=pers)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print(f"K={K} rows={rows} {prec} persistent={pers}: {1e3 * dt:.2f} ms per chain call = {1e6 * dt / T:.2f} us/reverse step, "
              f"{K * rows / dt:.0f} samples/s, launches {engine.last_launches(0)}, finite {bool(torch.isfinite(out['y']).all())}")
    del pms
    torch.cuda.empty_cache()
