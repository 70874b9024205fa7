"""Config-5 sweep (BASELINE.json configs[4]): posterior samples/s over T x draws x images at the shipped trunk width,
on 1..8 GPUs of one box (launch with torchrun for N > 1) through the product's own sharded entry point
`sample_ensemble`, with the reference CPU sampler beside the small end (N = 64, D <= 10).

    python tools/sweep_dist.py [quick] > gpurun_out/sweep.jsonl
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_dist.py

Members are random-init packed trunks; the image features xf are synthetic (the sweep measures the sampler; the encoder
prologue is reported by bench.py).  Long chains are timed on `steps_timed` reverse steps and extrapolated to T (the
per-step cost does not depend on t).  Time = max over ranks of the CUDA-event time around the whole sharded call,
gather included."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import nested_diffusion_b200 as nd  # noqa: E402
from nested_diffusion_b200.schedule import make_beta_schedule, schedule_tensors  # noqa: E402
from tests.test_gpu_parity import _rand_trunk_sd  # noqa: E402

F, C, K = 4096, 2, 5
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


class FeatureEnsemble(nd.NestedEnsemble):
    """Packed trunks + precomputed features: `encode` looks the rows up by global image index (first column of x)."""

    def __init__(self, members, xf):
        self.members, self.member_ids, self.device = members, list(range(len(members))), dev
        self.models, self.xf = [], xf
        self.max_rows_per_call = int(os.environ.get("LADINE_MAX_ROWS", nd.NestedEnsemble.MAX_ROWS_PER_CALL))   # A/B of the row cap

    def encode(self, xx):
        lo = int(xx[0, 0].item())
        return self.xf[:, lo:lo + xx.shape[0]]


_CPU = []


def cpu_reference(T, N, D):
    """The reference's own sampler on the host cores (oracle/_ref when staged, else the oracle port): one shipped-shape
    member, 64 images, one draw, as written (encoder re-evaluated every step).  Only used for N = 64, D <= 10 (the CPU
    rate per chain does not depend on D or N); T = 100 runs all 100 steps, T = 1000 is timed on 50 and extrapolated."""
    import bench
    if not _CPU:
        _CPU.append(bench.CpuReference())
    cpu = _CPU[0]
    steps = T if T <= 100 else 50
    v, dt = cpu.run(steps)            # samples/s extrapolated to bench.T_STEPS = 1000 reverse steps
    return v * bench.T_STEPS / T, cpu.kind, cpu.cores, steps


def run(T, D, N, budget_rowsteps=float(os.environ.get("SWEEP_BUDGET", "7e8")), reps=int(os.environ.get("SWEEP_REPS", "3"))):
    members = [nd.PackedMember(_rand_trunk_sd(s, F, C, min(T, 1000), dev), n_steps=min(T, 1000), precision="fp16")
               for s in range(K)]
    g = torch.Generator(device="cuda").manual_seed(0)
    xf = torch.randn(K, N, F, device=dev, generator=g)
    yh = torch.softmax(torch.randn(K, N, C, device=dev, generator=g), -1)
    ens = FeatureEnsemble(members, xf)
    x_idx = torch.arange(N, device=dev, dtype=torch.float32).view(N, 1)
    rows = K * N * D
    steps = int(max(8, min(T, budget_rowsteps * world // rows)))
    alphas, omabs = schedule_tensors(make_beta_schedule("linear", steps, 1e-4, 0.02))
    best = None
    for it in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        y0, probs = nd.sample_ensemble(ens, x_idx, yh, D, steps, alphas, omabs, seed=it, temperature=0.3162)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        best = ms if best is None or ms < best else best
    per_step_us = best * 1e3 / steps
    sps = rows / (per_step_us * 1e-6 * T)
    out = dict(n_gpus=world, T=T, D=D, N=N, K=K, chains=rows, steps_timed=steps, us_per_step=round(per_step_us, 1),
               samples_per_s=round(sps, 1), tflops_per_gpu=round(rows * (4.0 * F * F + 6 * F * C) / (per_step_us * 1e-6) / 1e12 / world, 1),
               finite=bool(torch.isfinite(y0).all()), images_per_rank=-(-N // world))
    if rank == 0 and N == 64 and D <= 10 and os.environ.get("SWEEP_CPU", "1") == "1":
        v, kind, cores, st = cpu_reference(T, N, D)
        out.update(cpu_samples_per_s=round(v, 3), cpu_kind=kind, cpu_cores=cores, cpu_steps_timed=st,
                   gpu_over_cpu=round(sps / v, 1))
    if rank == 0:
        print(json.dumps(out), flush=True)
    del members, ens, xf
    torch.cuda.empty_cache()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "points":      # python tools/sweep_dist.py points T,D,N [T,D,N ...]
        for spec in sys.argv[2:]:
            run(*[int(v) for v in spec.split(",")])
        sys.exit(0)
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    for T in (100, 1000):
        for D in (10, 100, 1000):
            for N in ((64, 1024) if quick else (64, 1024, 16384)):
                if N * D * K > 100e6:
                    continue
                run(T, D, N)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
