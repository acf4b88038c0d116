cd /root/repo
timeout 300 python -m pytest tests/test_gpu_mlp_field.py -x -q 2>&1 | tail -25
timeout 120 python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_gpu_mlp_field import make_field
f = make_field(3)
y = torch.randn(8192, 256, device="cuda")
with torch.no_grad():
    got = f(None, y); want = f.forward_reference(None, y)
torch.cuda.synchronize()
print("max abs diff", (got - want).abs().max().item(), "scale", want.abs().max().item(), "median diff", (got - want).abs().median().item())
print("row0 got", got[0, :6].tolist()); print("row0 want", want[0, :6].tolist())
import torch.nn.functional as Fn
wb, bb = f.weights, f.biases.to(torch.bfloat16)
def torch_bf16():
    h = y.to(torch.bfloat16)
    for l in range(3):
        h = Fn.linear(h, wb[l], bb[l])
        if l < 2: h = torch.tanh(h)
    return h.float()
for fn, name in ((lambda: f(None, y), "tcgen05 kernel"), (torch_bf16, "torch bf16 (cuBLAS) chain"), (lambda: f.forward_reference(None, y), "torch fp32 reference")):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 20 * 1e3, "us per eval (B=8192), eager launches")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 20 * 1e3, "us per eval (B=8192), CUDA-graph replay of 20")
PY
