"""Time the step-invariant encoder prologue (PyTorch) at the shipped shape: python tools/encoder_probe.py"""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda")
models = bench.build_members(dev)[:2]
x = torch.rand(70, bench.DX, device=dev)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
m = models[0]
with torch.no_grad():
    print("encode (fp32, allow_tf32=%s): %.2f ms" % (torch.backends.cuda.matmul.allow_tf32, t(lambda: m.encode(x))))
    W = m.encoder_x[0].weight
    print("  first Linear only: %.2f ms" % t(lambda: torch.nn.functional.linear(x, W)))
    print("  x @ W.T (mm): %.2f ms" % t(lambda: x @ W.t()))
    Wt = W.t().contiguous()
    print("  x @ Wt (K-major W): %.2f ms" % t(lambda: x @ Wt))
    torch.backends.cuda.matmul.allow_tf32 = True
    print("encode (allow_tf32=True): %.2f ms" % t(lambda: m.encode(x)))
    torch.backends.cuda.matmul.allow_tf32 = False
    xb, Wb = x.bfloat16(), W.bfloat16()
    print("  bf16 linear: %.2f ms" % t(lambda: torch.nn.functional.linear(xb, Wb)))
    x2 = torch.rand(140, bench.DX, device=dev)
    print("  first Linear, 140 rows: %.2f ms" % t(lambda: torch.nn.functional.linear(x2, W)))
