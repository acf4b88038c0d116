"""Built-in analytic vector fields.

Each field is an ordinary ``nn.Module`` computing ``f(t, y)`` with plain PyTorch ops, so
it works as the ``f`` of an ``ODETerm`` anywhere (including the reference torchode).
When such a field is the term of a solve on a B200, ``AutoDiffAdjoint.solve`` recognises
it and runs the fully fused whole-solve kernel (``tode_solve_fused``), whose in-register
evaluation performs exactly the arithmetic of ``forward`` below: one IEEE rounding per
PyTorch op, in the same order, no FMA contraction.
"""
import torch
import torch.nn as nn

from . import _cabi


class BuiltinField(nn.Module):
    """Marker base class: ``field_id`` / ``params`` select the fused kernel's field."""

    field_id: int = -1
    n_features = None  # required feature count, None = any (<= 4 on the fused path)

    def params(self):
        raise NotImplementedError


class LinearDecay(BuiltinField):
    """``y' = rate * y`` (README example: ``rate = -0.5``)."""

    field_id = _cabi.FIELD_LINEAR

    def __init__(self, rate: float):
        super().__init__()
        self.rate = float(rate)

    def params(self):
        return [self.rate]

    def forward(self, t, y):
        return self.rate * y


class VanDerPol(BuiltinField):
    """``x' = v``, ``v' = mu (1 - x^2) v - x``."""

    field_id = _cabi.FIELD_VAN_DER_POL
    n_features = 2

    def __init__(self, mu: float):
        super().__init__()
        self.mu = float(mu)

    def params(self):
        return [self.mu]

    def forward(self, t, y):
        x, v = y[..., 0], y[..., 1]  # `...`: also valid for an unbatched sample (torch.func.vmap)
        dv = self.mu * (1 - x * x) * v - x
        return torch.stack((v, dv), dim=-1)


class LotkaVolterra(BuiltinField):
    """``x' = alpha x - beta x z``, ``z' = delta x z - gamma z``."""

    field_id = _cabi.FIELD_LOTKA_VOLTERRA
    n_features = 2

    def __init__(self, alpha: float = 1.5, beta: float = 1.0, delta: float = 1.0, gamma: float = 3.0):
        super().__init__()
        self.alpha, self.beta, self.delta, self.gamma = map(float, (alpha, beta, delta, gamma))

    def params(self):
        return [self.alpha, self.beta, self.delta, self.gamma]

    def forward(self, t, y):
        x, z = y[..., 0], y[..., 1]
        xz = x * z
        dx = self.alpha * x - self.beta * xz
        dz = self.delta * xz - self.gamma * z
        return torch.stack((dx, dz), dim=-1)


class _KernelField(torch.autograd.Function):
    """A field evaluated by a forward-only CUDA kernel, differentiable with respect to its input: the
    backward pass is the vector-Jacobian product of the same computation written in PyTorch ops
    (``forward_reference``), recomputed from the saved input."""

    @staticmethod
    def forward(ctx, field, t, y):
        ctx.field, ctx.t = field, t
        ctx.save_for_backward(y)
        return field._forward_kernel(t, y)

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        with torch.enable_grad():
            yr = y.detach().requires_grad_()
            out = ctx.field.forward_reference(ctx.t, yr)
            (gy,) = torch.autograd.grad(out, yr, g)
        return None, None, gy


class TanhMLP256(nn.Module):
    """Neural-ODE vector field ``y -> W_L tanh(... tanh(W_1 y + b_1) ...) + b_L`` of width 256
    (BASELINE.json configs[3]) evaluated by ONE hand-written tcgen05 kernel per call: bf16 tensor
    core GEMMs with fp32 accumulation in TMEM, bias + tanh epilogue out of TMEM, the activation
    tile stays in shared memory between layers (``tode_mlp_tanh256_forward``).

    It is an ordinary ``f`` for ``ODETerm`` (autonomous: ``t`` is ignored) and runs through the
    stage-wise route.  ``forward_reference`` is the same computation in plain PyTorch fp32 ops
    on the bf16-rounded operands (for tests and for running the field on the reference).
    """

    WIDTH = 256

    def __init__(self, weights: torch.Tensor, biases: torch.Tensor):
        super().__init__()
        assert weights.ndim == 3 and weights.shape[1:] == (self.WIDTH, self.WIDTH)
        assert biases.shape == (weights.shape[0], self.WIDTH)
        self.register_buffer("weights", weights.detach().to(torch.bfloat16).contiguous())
        self.register_buffer("biases", biases.detach().to(torch.float32).contiguous())

    @staticmethod
    def from_sequential(seq: nn.Sequential) -> "TanhMLP256":
        linears = [m for m in seq if isinstance(m, nn.Linear)]
        others = [m for m in seq if not isinstance(m, (nn.Linear, nn.Tanh))]
        assert linears and not others, "expected Linear layers separated by Tanh"
        return TanhMLP256(torch.stack([m.weight for m in linears]), torch.stack([m.bias for m in linears]))

    @property
    def n_layers(self) -> int:
        return self.weights.shape[0]

    def forward(self, t, y):
        # weights are buffers (inference field): gradients flow to the state only
        if torch.is_grad_enabled() and y.requires_grad:
            return _KernelField.apply(self, t, y)
        return self._forward_kernel(t, y)

    def _forward_kernel(self, t, y):
        from . import _launch

        _launch.require_cuda(y, self.weights)
        assert y.dtype == torch.float32 and y.ndim == 2 and y.shape[1] == self.WIDTH
        y = _launch.dense16(y)
        out = torch.empty_like(y)
        with torch.cuda.device(y.device):
            _cabi.check(_cabi.lib().tode_mlp_tanh256_forward(
                y.data_ptr(), self.weights.data_ptr(), self.biases.data_ptr(), out.data_ptr(), y.shape[0],
                self.n_layers, _launch.stream_ptr(y.device)), "tode_mlp_tanh256_forward")
        return out

    def forward_reference(self, t, y):
        h = y.to(torch.bfloat16).to(torch.float32)
        for layer in range(self.n_layers):
            h = h @ self.weights[layer].to(torch.float32).T + self.biases[layer]
            if layer + 1 < self.n_layers:
                h = torch.tanh(h).to(torch.bfloat16).to(torch.float32)
        return h


class Heat1D(nn.Module):
    """Method-of-lines field of the 1-D heat equation with Dirichlet ends (BASELINE.json configs[4]):
    ``out[:, i] = kappa * ((y[:, i+1] - 2 y[:, i]) + y[:, i-1])``, zero at both ends -- one HBM pass in
    a hand-written kernel (``tode_heat1d_forward``) instead of the five passes of the PyTorch
    expression.  An ordinary autonomous ``f`` for ``ODETerm`` (stage-wise route);
    ``forward_reference`` is the same arithmetic, same rounding order, in PyTorch ops."""

    def __init__(self, kappa: float):
        super().__init__()
        self.kappa = float(kappa)

    def forward(self, t, y):
        if torch.is_grad_enabled() and y.requires_grad:
            return _KernelField.apply(self, t, y)
        return self._forward_kernel(t, y)

    def _forward_kernel(self, t, y):
        from . import _launch

        _launch.require_cuda(y)
        assert y.ndim == 2
        y = _launch.dense16(y)
        out = torch.empty_like(y)
        with torch.cuda.device(y.device):
            _cabi.check(_cabi.lib().tode_heat1d_forward(
                y.data_ptr(), out.data_ptr(), y.shape[0], y.shape[1], self.kappa, _launch.dtype_id(y.dtype),
                _launch.stream_ptr(y.device)), "tode_heat1d_forward")
        return out

    def forward_reference(self, t, y):
        out = torch.zeros_like(y)
        out[:, 1:-1] = self.kappa * ((y[:, 2:] - 2 * y[:, 1:-1]) + y[:, :-2])
        return out
