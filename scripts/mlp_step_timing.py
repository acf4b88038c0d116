"""Phase timestamps (SM clock cycles, CTA 0) of one stage of the step-fused tcgen05 MLP launch
(tode_mlp_tanh256_step_forward): needs a library built with -DTODE_MLP_TIMING=<stage>
(TORCHODE_B200_LIB=build_variants/mlp_timing_s<stage>.so).  Also times the untimed-build launch with CUDA events
when run against the normal library (no stamps: prints only the event time)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
import torchode_b200 as to
from torchode_b200 import _cabi, _launch
from torchode_b200.fields import TanhMLP256

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
field = TanhMLP256(torch.randn(3, 256, 256) * 0.06, torch.randn(3, 256) * 0.1).to("cuda")
tab = to.Dopri5().to_cabi()
y = torch.randn(B, 256, device="cuda")
ks = [torch.randn(B, 256, device="cuda") for _ in range(7)]
y1 = torch.empty_like(y)
dt = torch.full((B,), 0.05, device="cuda")
st = _launch._minimal_state(y, dt)
lib, stream = _cabi.lib(), _launch.stream_ptr(y.device)


def launch():
    _cabi.check(lib.tode_mlp_tanh256_step_forward(C.byref(tab), C.byref(st), _launch.kptrs(ks), y1.data_ptr(),
                                                  field.weights.data_ptr(), field.biases.data_ptr(), 3, stream), "step")


for _ in range(5):
    launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    launch()
e1.record()
torch.cuda.synchronize()
print(f"B={B}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per step launch (6 evaluations)")
if "timing" in os.environ.get("TORCHODE_B200_LIB", ""):
    stamps = ks[6][0].view(torch.int64)[:17].tolist()
    n = stamps.index(-1) if -1 in stamps else 16
    if "_op" in os.environ["TORCHODE_B200_LIB"]:  # -DTODE_MLP_TIMING_OPERAND=<thread>: extra stamps in the operand phase
        names = ["kernel start", "stage begins", "newest operand read", "y read (TMEM)", "rows formed + stored",
                 "proxy fence", "barrier"] + ["(layer phases; 'weights in' only for thread 0)"] * 10
        prev = 0
        for name, s_ in zip(names, stamps[:n]):
            print(f"{name:28s} {s_:8d} cycles  (+{s_ - prev})")
            prev = s_
        sys.exit(0)
    names = ["kernel start", "stage begins", "operand rows formed"] + [x for l in range(3) for x in (
        f"L{l} weights in", f"L{l} MMA done", f"L{l} epilogue done")] + ["next stage begins"] + ["..."] * 8
    prev = 0
    for name, s_ in zip(names, stamps[:n]):
        print(f"{name:22s} {s_:8d} cycles  (+{s_ - prev})")
        prev = s_
