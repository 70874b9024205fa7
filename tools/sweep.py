"""Config sweep (BASELINE.json configs[2..4]): posterior samples/s over
T x draws x images at the shipped trunk width.  Members are random-init packed trunks; features are synthetic.
    python tools/sweep.py [quick]
Long chains are timed on `steps_timed` reverse steps and extrapolated to T (the per-step cost does not depend on t)."""
import json, sys, time
import torch
sys.path.insert(0, ".")
import nested_diffusion_b200 as nd
from nested_diffusion_b200 import engine
from nested_diffusion_b200.schedule import make_beta_schedule, schedule_tensors

F, C, K = 4096, 2, 5
dev = torch.device("cuda")

def member(seed, T):
    gg = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.rand(*s, device=dev, generator=gg)
    sd = {}
    for l, i in ((1, 2 * C), (2, F), (3, F)):
        b = 1 / i ** 0.5
        sd[f"lin{l}.lin.weight"] = (r(F, i) * 2 - 1) * b
        sd[f"lin{l}.lin.bias"] = (r(F) * 2 - 1) * b
        sd[f"lin{l}.embed.weight"] = r(T + 1, F)
        sd[f"unetnorm{l}.weight"] = r(F) + 0.5
        sd[f"unetnorm{l}.bias"] = torch.randn(F, device=dev, generator=gg) * 0.2
        sd[f"unetnorm{l}.running_mean"] = torch.randn(F, device=dev, generator=gg) * 0.3
        sd[f"unetnorm{l}.running_var"] = r(F) + 0.5
    sd["lin4.weight"] = (r(C, F) * 2 - 1) / F ** 0.5
    sd["lin4.bias"] = (r(C) * 2 - 1) / F ** 0.5
    return sd

def run(T, D, N, k=K, budget_rowsteps=7e8):
    members = [nd.PackedMember(member(s, T), n_steps=T, precision="fp16") for s in range(k)]
    ens = nd.NestedEnsemble.__new__(nd.NestedEnsemble)
    ens.members, ens.member_ids, ens.device, ens.models, ens.max_rows_per_call = members, list(range(k)), dev, [], nd.NestedEnsemble.MAX_ROWS_PER_CALL
    g = torch.Generator(device="cuda").manual_seed(0)
    xf = torch.randn(k, N, F, device=dev, generator=g)
    yh = torch.softmax(torch.randn(k, N, C, device=dev, generator=g), -1)
    rows = k * N * D
    steps = int(max(8, min(T, budget_rowsteps // rows)))
    alphas, omabs = schedule_tensors(make_beta_schedule("linear", steps, 1e-4, 0.02))
    best = None
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        res = ens.sample(None, yh, D, steps, alphas, omabs, seed=it, xf=xf, temperature=0.3162)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    per_step_us = best * 1e3 / steps
    sps = rows / (per_step_us * 1e-6 * T)
    tf = rows * (4.0 * F * F + 6 * F * C) / (per_step_us * 1e-6) / 1e12
    out = dict(T=T, D=D, N=N, K=k, chains=rows, steps_timed=steps, us_per_step=round(per_step_us, 1),
               samples_per_s=round(sps, 1), tflops=round(tf, 1), finite=bool(torch.isfinite(res.y0).all()))
    print(json.dumps(out), flush=True)
    del members, ens, xf
    torch.cuda.empty_cache()
    return out

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":   # one bounded configuration per process: python tools/sweep.py one T D N K
        T, D, N, k = (int(v) for v in sys.argv[2:6])
        run(T, D, N, k=k)
        sys.exit(0)
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    print("# config 1: single member, batch 64"); run(1000, 1, 64, k=1)
    print("# config 2: K=5, N=70, D=20 (and D=100/member)"); run(1000, 20, 70); run(1000, 100, 70)
    print("# config 3: ISIC-shaped, N=1024, D=20"); run(1000, 20, 1024)
    print("# config 5 sweep")
    for T in (100, 1000):
        for D in (10, 100, 1000):
            for N in ((64, 1024) if quick else (64, 1024, 16384)):
                if N * D * K > 100e6:
                    continue
                run(T, D, N)
