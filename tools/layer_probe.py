"""Time one GEMM layer launch (ladine_debug_layer) per tile geometry: python tools/layer_probe.py F rows [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from nested_diffusion_b200 import _capi, engine
from oracle import ladine_oracle as orc

F, rows = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
T = 4
sd = orc.synth_state_dict(7, F, 16, 16, 2, T)
pm = engine.PackedMember({k: v.cuda() for k, v in sd.items()}, n_steps=T, precision="fp16")
rows_pad = (rows + 255) // 256 * 256
h_in = torch.rand(rows_pad, pm.Fp, device="cuda").half()
h_out = torch.zeros(rows_pad, pm.Fp, dtype=torch.float16, device="cuda")
part = torch.zeros(rows_pad, pm.Fp // 256, 2, pm.Cp, device="cuda")
lib, h = _capi.load(), _capi.handle(0)
st = torch.cuda.current_stream().cuda_stream
for ctas in (1, 2, 3):
    engine.set_option(0, "ctas", ctas)
    for layer in (2, 3):
        for _ in range(5):
            _capi.check(h, lib.ladine_debug_layer(h, pm.ptr, layer, 1, h_in.data_ptr(), rows, h_out.data_ptr(), part.data_ptr(), st))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps):
            _capi.check(h, lib.ladine_debug_layer(h, pm.ptr, layer, 1, h_in.data_ptr(), rows, h_out.data_ptr(), part.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        print(f"F={F} rows={rows} ctas={ctas} layer={layer}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per launch (incl. schedule upload)")
engine.set_option(0, "ctas", 0)
