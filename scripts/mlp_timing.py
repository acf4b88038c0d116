"""Phase timestamps (SM clock cycles, CTA 0) of the tcgen05 MLP field: needs a library built with
-DTODE_MLP_TIMING=0 (stamps from stage 0 on, i.e. the plain evaluation; TORCHODE_B200_LIB=build_variants/mlp_timing.so)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
f = bench._mlp_field("cuda")
y = torch.randn(8192, 256, device="cuda")
for _ in range(3):
    out = f(None, y)
torch.cuda.synchronize()
st = out[0].view(torch.int64)[:16].tolist()
names = ["start", "A tile loaded"] + [x for l in range(3) for x in (f"L{l} weights in", f"L{l} MMA done", f"L{l} epilogue done")]
prev = 0
for n, s in zip(names, st):
    print(f"{n:20s} {s:8d} cycles  (+{s - prev})")
    prev = s
