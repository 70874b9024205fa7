"""Light profiling target for `ncu --set full`: the config-3 call shape (K=5 members x N=1024 images x D=20 draws,
F=4096) for a few reverse steps, and the encoder prologue of one shipped-shape member on the same 1024 images.

    python tools/profile_step.py [precision] [n_steps]

Launch order (what -s / -c count against): 3 x ladine_encode (3 enc_gemm_kernel each: layer 1 with K=150528 first),
then two identical sample calls of 1 + 3 * n_steps kernels (tail/head init, then gemm2, gemm3, tail/head per step).
Not a benchmark: numbers printed here are for orientation only."""
import argparse
import sys
import time

import torch

sys.path.insert(0, ".")
import nested_diffusion_b200 as nd  # noqa: E402
from nested_diffusion_b200 import engine  # noqa: E402
from nested_diffusion_b200.schedule import coef_table, make_beta_schedule, schedule_tensors  # noqa: E402
from tests.test_gpu_parity import _rand_trunk_sd  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
K, N, D, F, C, T, Dx = 5, 1024, 20, 4096, 2, 1000, 150528
dev = torch.device("cuda")
ns = argparse.Namespace
cfg = ns(diffusion=ns(timesteps=T), data=ns(num_classes=C, dataset="ISICSkinCancer"),
         model=ns(data_dim=Dx, arch="linear", feature_dim=F, hidden_dim=F))
torch.manual_seed(0)
with torch.device(dev):
    model = nd.ConditionalModel(cfg, guidance=True).eval()
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(N, Dx, device=dev, generator=g)
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    xf1 = engine.encode_members([model], x, mode="kernel")
    torch.cuda.synchronize()
    print(f"encode 1 member x {N} images: {1e3 * (time.perf_counter() - t0):.2f} ms")

pms = [nd.PackedMember(_rand_trunk_sd(2000 + k, F, C, T, dev), n_steps=T, precision=prec) for k in range(K)]
xf = torch.randn(K, N, F, device=dev, generator=g)
yh = torch.softmax(torch.randn(K, N, C, device=dev, generator=g), -1)
alphas, omabs = schedule_tensors(make_beta_schedule("linear", T, 1e-4, 0.02))
coef = coef_table(alphas, omabs, T)
import os  # noqa: E402
for kv in os.environ.get("LADINE_OPTIONS", "").split(","):   # e.g. LADINE_OPTIONS=tail_vec=4,order=2
    if "=" in kv:
        engine.set_option(0, kv.split("=")[0], int(kv.split("=")[1]))
engine.set_profiling(0, True)
for i in range(2):
    torch.cuda.synchronize()
    engine.get_profile(0)
    t0 = time.perf_counter()
    out = engine.sample_chains(pms, xf, yh, yh, coef, D, seed=3, t_first=T - 1, t_last=T - n_steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = engine.get_profile(0)
    print(f"{prec}: {n_steps} reverse steps of {K * N * D} chains: {1e6 * dt / n_steps:.1f} us/step, "
          f"launches {engine.last_launches(0)}, finite {bool(torch.isfinite(out['y']).all())}; per launch (us): "
          + ", ".join(f"{k} {1e3 * v[0] / max(1, v[1]):.1f}" for k, v in prof.items() if isinstance(v, tuple)))
