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
