"""Quick timing probe of the batched sampler (not the bench): python tools/perf_probe.py K N D F T prec"""
import sys, time
import torch
sys.path.insert(0, ".")
import nested_diffusion_b200 as nd
from nested_diffusion_b200 import engine
from nested_diffusion_b200.schedule import coef_table, make_beta_schedule, schedule_tensors

K, N, D, F, T = (int(v) for v in sys.argv[1:6])
prec = sys.argv[6] if len(sys.argv) > 6 else "auto"
lanes = int(sys.argv[7]) if len(sys.argv) > 7 else 2
engine.set_option(0, "lanes", lanes)
ctas = int(sys.argv[8]) if len(sys.argv) > 8 else 0
engine.set_option(0, "ctas", ctas)
pdl = 0  # programmatic dependent launch was removed in round 2 (argv[9] kept for positional compatibility)
fuse = int(sys.argv[10]) if len(sys.argv) > 10 else 0
engine.set_option(0, "fuse", fuse)
tail_vec = int(sys.argv[11]) if len(sys.argv) > 11 else 0
engine.set_option(0, "tail_vec", tail_vec)
order = int(sys.argv[12]) if len(sys.argv) > 12 else 0
engine.set_option(0, "order", order)
import os
for kv in os.environ.get("LADINE_OPTIONS", "").split(","):   # e.g. LADINE_OPTIONS=pace=2
    if "=" in kv:
        engine.set_option(0, kv.split("=")[0], int(kv.split("=")[1]))
C = 2
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
def member(seed):
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
members = [nd.PackedMember(member(k), n_steps=T, precision=prec) for k in range(K)]
xf = torch.randn(K, N, F, device=dev, generator=g)
yh = torch.softmax(torch.randn(K, N, C, device=dev, generator=g), -1)
alphas, omabs = schedule_tensors(make_beta_schedule("linear", T, 1e-4, 0.02))
coef = coef_table(alphas, omabs, T)
for it in range(4):
    if it == 3:
        engine.set_profiling(0, True)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0.record()
    y = engine.sample_chains(members, xf, yh, yh, coef, D, seed=it)["y"]
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flops = K * N * D * T * (4.0 * F * F + 6 * F * C)
    print(f"K={K} N={N} D={D} F={F} T={T} {members[0].precision}: {ms:.2f} ms  ({ms/T*1e3:.1f} us/step)  "
          f"{flops/ms/1e9:.1f} TFLOP/s  samples/s(T=1000 equiv)={K*N*D/(ms/1e3)*T/1000:.0f}  host enqueue {1e3*(t1-t0):.1f} ms  "
          f"finite={bool(torch.isfinite(y).all())} launches={engine.last_launches(0)} lanes={lanes} ctas={ctas} pdl={pdl} fuse={fuse} tail_vec={tail_vec} order={order}")
print("profile:", engine.get_profile(0))
