#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -E "passed|failed|rror|rel err vs FP64" | tail -8
echo "== encoder"; timeout 200 python - <<'PY'
import sys, torch; sys.path.insert(0, ".")
import bench
from nested_diffusion_b200 import engine
dev=torch.device("cuda"); m=bench.build_members(dev)[0]; x=torch.rand(70,bench.DX,device=dev)
def t(fn,n=5):
    fn(); torch.cuda.synchronize(); e0,e1=torch.cuda.Event(True),torch.cuda.Event(True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print("fp32 %.2f ms  split %.2f ms" % (t(lambda: engine.encode_features(m,x,"fp32")), t(lambda: engine.encode_features(m,x))))
PY
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['clocks'])"; tail -3 gpurun_out/bench.err
