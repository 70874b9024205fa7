"""Every kernel family of libladine at tiny shapes in one short process (seconds on a B200):

    python tools/kernel_family_sweep.py

Covers: member packing (all precisions), packed-image export / import, the encoder prologue (split-K + finish), the
tile path in its three geometries (single CTAs, CTA pairs with a half tile, slim tiles) for FP16 / BF16, the FP32X split
kernels, the fused tail + head option, the FP32 SMEM-resident kernel, the whole-chain persistent kernel, Philox and
injected noise, trajectory and probability outputs.  Meant as the target of a launch list / quick bring-up check on a new
box (compute-sanitizer is closed on the B200 pool, so memory safety rests on the bounded waits + the parity tests);
not a parity test (tests/test_gpu_parity.py is): it only checks that the results are finite."""
import argparse
import sys

import torch

sys.path.insert(0, ".")
import nested_diffusion_b200 as nd  # noqa: E402
from nested_diffusion_b200 import engine  # noqa: E402
from nested_diffusion_b200.schedule import coef_table, make_beta_schedule, schedule_tensors  # noqa: E402
from oracle import ladine_oracle as orc  # noqa: E402  (synthetic members only: this is a test tool)

ns = argparse.Namespace
dev = torch.device("cuda")
T, C = 4, 2
alphas, omabs = schedule_tensors(make_beta_schedule("linear", T, 1e-4, 0.02))
coef = coef_table(alphas, omabs, T)
g = torch.Generator().manual_seed(0)


def member(F, H, Dx, prec, seed):
    sd = orc.synth_state_dict(seed, F, H, Dx, C, T)
    cfg = ns(diffusion=ns(timesteps=T), data=ns(num_classes=C, dataset="ChestXRay"),
             model=ns(data_dim=Dx, arch="linear", feature_dim=F, hidden_dim=H))
    m = nd.ConditionalModel(cfg, guidance=True)
    m.load_state_dict(sd)
    return m.eval().to(dev)


def check(tag, *tensors):
    torch.cuda.synchronize()
    assert all(bool(torch.isfinite(t).all()) for t in tensors), tag
    print("ok:", tag, flush=True)


# encoder prologue + packed images
m256 = [member(256, 128, 300, "fp16", 10 + k) for k in range(2)]
x = torch.rand(150, 300, generator=g).to(dev)
xf = engine.encode_members(m256, x, mode="kernel")
check("encoder (2 members x 150 images, K=300 -> split-K + finish)", xf)
pm, pe = engine.packed_member_of(m256[0], "fp16"), engine.packed_encoder_of(m256[0])
packed = nd.PackedModel(nd.PackedMember.from_image(pm.export_image(), dev), nd.PackedEncoder.from_image(pe.export_image(), dev))
check("packed images round trip", engine.encode_members([packed], x))

yh = torch.softmax(torch.randn(2, 150, C, generator=g), -1).to(dev)
for prec in ("fp16", "bf16", "fp32x"):
    pms = [engine.packed_member_of(m, prec) for m in m256]
    for geom in ((1, 2, 3) if prec != "fp32x" else (0,)):
        engine.set_option(0, "ctas", geom)
        out = engine.sample_chains(pms, xf, yh, yh, coef, 2, seed=5, trajectory=True, temperature=0.17)
        check(f"tile path {prec} geometry {geom}: 2 members x 300 rows (pairs: 1 full + 1 half tile)", *out.values())
    engine.set_option(0, "ctas", 0)
noise = torch.randn(2, 2, T, 150, C, generator=g).to(dev)
out = engine.sample_chains([engine.packed_member_of(m, "fp16") for m in m256], xf, yh, yh, coef, 2, noise=noise)
check("tile path, injected noise", out["y"])
engine.set_option(0, "fuse", 1)
out = engine.sample_chains([engine.packed_member_of(m, "fp16") for m in m256], xf, yh, yh, coef, 2, seed=6)
engine.set_option(0, "fuse", 0)
check("tile path, fused tail + head", out["y"])

# FP32 SMEM-resident kernel
m64 = member(64, 32, 40, "fp32", 30)
x64 = torch.rand(33, 40, generator=g).to(dev)
out = engine.sample_chains([engine.packed_member_of(m64, "fp32")], engine.encode_members([m64], x64), yh[:1, :33], yh[:1, :33],
                           coef, 3, seed=7, trajectory=True, temperature=0.3)
check("resident FP32 kernel (F=64, 99 rows)", *out.values())

# whole-chain persistent kernel (one cooperative launch)
m512 = member(512, 64, 80, "fp16", 40)
x512 = torch.rand(40, 80, generator=g).to(dev)
out = engine.sample_chains([engine.packed_member_of(m512, "fp16")], engine.encode_members([m512], x512), yh[:1, :40],
                           yh[:1, :40], coef, 1, seed=8, persistent=True)
check(f"persistent chain kernel (F=512, 40 chains, launches {engine.last_launches(0)})", out["y"])
print("kernel family sweep done")
