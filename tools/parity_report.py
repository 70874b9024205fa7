"""Measured parity of the CUDA path against the reference's recorded outputs (tests/golden), one line per
fixture x precision:  python tools/parity_report.py > gpurun_out/parity.txt
Not a test (tests/test_gpu_parity.py asserts the bars); this prints the actual errors so the bars can be judged."""
import sys

import torch

sys.path.insert(0, ".")
import nested_diffusion_b200 as nd  # noqa: E402
from nested_diffusion_b200 import diffusion_utils as du  # noqa: E402
from tests.golden_util import ChainFixture, EnsembleFixture, Fixture, names, rel_err  # noqa: E402
from tests.test_gpu_parity import make_model, precisions_for  # noqa: E402

print(f"# device: {torch.cuda.get_device_name(0)}; error = max|y - y_ref| / max(1, max|y_ref|) over the stored trajectory rows")
print("fixture,precision,F,T,rows,ref_absmax,rel_err_traj,max_abs_y0,labels_equal")
for name in names("chain"):
    fx = ChainFixture(name)
    m = fx.meta
    sd, x, yhat, noise, alphas, omabs = fx.materialize()
    model = make_model(m, sd)
    for prec in precisions_for(m["F"]):
        with torch.no_grad():
            seq = du.p_sample_loop(model, x.cuda(), yhat.cuda(), yhat.cuda(), m["T"], alphas.cuda(), omabs.cuda(),
                                   only_last_sample=False, noise=noise.cuda(), precision=prec)
        traj = torch.stack(seq).cpu()
        y0 = traj[-1]
        keep = m.get("keep")
        e_traj = rel_err(traj[keep], fx["traj"]) if keep is not None and "traj" in fx.arrays else rel_err(y0, fx["y0"])
        print(f"{name},{prec},{m['F']},{m['T']},{m['B']},{float(fx['y0'].abs().max()):.3g},{e_traj:.3e},"
              f"{float((y0 - fx['y0']).abs().max()):.3e},{bool(torch.equal(y0.argmax(1), fx['y0'].argmax(1)))}")
    del model
    torch.cuda.empty_cache()
for name in names("ensemble"):
    fx = EnsembleFixture(name)
    m = fx.meta
    sds, x, y0hats, noise, alphas, omabs = fx.materialize()
    models = [make_model(m, sd) for sd in sds]
    for prec in precisions_for(m["F"]):
        ens = nd.NestedEnsemble(models, precision=prec)
        with torch.no_grad():
            res = ens.sample(x.cuda(), [y.cuda() for y in y0hats], m["D"], m["T"], alphas.cuda(), omabs.cuda(),
                             noise=noise.cuda(), temperature=0.1737)
        y0 = res.y0.cpu()
        print(f"{name},{prec},{m['F']},{m['T']},{m['K']}x{m['D']}x{m['N']},{float(fx['y0'].abs().max()):.3g},"
              f"{rel_err(y0, fx['y0']):.3e},{float((y0 - fx['y0']).abs().max()):.3e},"
              f"{bool(torch.equal(y0.argmax(-1), fx['y0'].argmax(-1)))}")


# ---- one layer of the FP32X path and the encoder prologue against torch FP64 (what "FP32-grade" means in numbers) ----
import ctypes as C  # noqa: E402

from nested_diffusion_b200 import _capi, engine  # noqa: E402
from oracle import ladine_oracle as orc  # noqa: E402

print("# FP32X layer (ladine_debug_layer) vs torch FP64 on the unrounded FP32 operands: error / max|output|")
print("kernel,F,rows,layer2_rel_err,lin4_partial_rel_err")
for F, rows in ((512, 300), (4096, 200)):
    T, Cc, t = 4, 2, 2
    sd = orc.synth_state_dict(7, F, 16, 16, Cc, T)
    pm = engine.PackedMember({k: v.cuda() for k, v in sd.items()}, n_steps=T, precision="fp32x")
    p = orc.fold_member(sd, T, torch.float64)
    g = torch.Generator().manual_seed(1)
    rows_pad = (rows + 255) // 256 * 256
    h32 = torch.zeros(rows_pad, pm.Fp)
    h32[:rows, :F] = torch.rand(rows, F, generator=g) * 2
    hi = h32.half()
    lo = (h32 - hi.float()).half()
    h_in = torch.cat([hi, lo], dim=1).contiguous().cuda()
    exact_in = (hi.double() + lo.double())[:rows, :F]
    lib, h = _capi.load(), _capi.handle(0)
    stream = torch.cuda.current_stream().cuda_stream
    h_out = torch.zeros(rows_pad, 2 * pm.Fp, dtype=torch.float16, device="cuda")
    _capi.check(h, lib.ladine_debug_layer(h, pm.ptr, 2, t, h_in.data_ptr(), rows, h_out.data_ptr(), None, stream))
    part = torch.zeros(rows_pad, pm.Fp // 256, 2, pm.Cp, device="cuda")
    _capi.check(h, lib.ladine_debug_layer(h, pm.ptr, 3, t, h_in.data_ptr(), rows, None, part.data_ptr(), stream))
    torch.cuda.synchronize()
    want = torch.nn.functional.softplus(p["A2"][t] * (exact_in @ p["W2"].T) + p["C2"][t])
    ho = h_out.cpu().double()
    got = ho[:rows, :F] + ho[:rows, pm.Fp:pm.Fp + F]
    h3 = torch.nn.functional.softplus(p["A3"][t] * (exact_in @ p["W3"].T) + p["C3"][t])
    want_eps = h3 @ p["W4"].T
    got_eps = part[:rows, :, :, :Cc].double().sum(dim=(1, 2)).cpu()
    print(f"trunk_split_kernel,{F},{rows},{float((got - want).abs().max() / want.abs().max()):.2e},"
          f"{float((got_eps - want_eps).abs().max() / max(1.0, float(want_eps.abs().max()))):.2e}")

print("# encoder prologue (ladine_encode) vs torch FP64; PyTorch FP32 beside it: error / max|xf|")
print("kernel,Dx,H,F,N,kernel_rel_err,torch_fp32_rel_err")
for Dx, H, F, N in ((1000, 200, 300, 70), (150528, 256, 256, 70), (4096, 4096, 4096, 128)):
    meta = dict(T=4, C=2, Dx=Dx, F=F, H=H, guidance=True)
    sd = orc.synth_state_dict(31, F, H, Dx, 2, 4)
    model = make_model(meta, sd)
    g = torch.Generator().manual_seed(8)
    x = torch.rand(N, Dx, generator=g).cuda()
    with torch.no_grad():
        got = engine.encode_features(model, x, mode="kernel").cpu().double()
        ref32 = engine.encode_features(model, x, mode="torch").cpu().double()
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        want = orc.encoder_features(sd64, x.cpu().double())
    scale = float(want.abs().max())
    print(f"enc_gemm_kernel,{Dx},{H},{F},{N},{float((got - want).abs().max()) / scale:.2e},"
          f"{float((ref32 - want).abs().max()) / scale:.2e}")
    del model
    torch.cuda.empty_cache()
