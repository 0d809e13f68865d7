"""Drop-in for ``4-bit quantization/quant_ultra.py`` (QU): the DoReFa-style fixed-grid quantizers and the
``Conv2d_Q`` / ``Linear_Q`` factories, on sm_100a kernels.

Same names, constructor arguments, ``w_bit`` / ``quantize_fn`` attributes and ``state_dict`` (plain nn.Conv2d /
nn.Linear keys) as upstream; the returned classes keep ``nn.Conv2d`` / ``nn.Linear`` as their DIRECT base
(``torch_export.py:27,62,100,110`` dispatches on ``__base__``).

No-grad forward: weight codes come from the fused tanh/max/round kernels (cached per weight version) and feed the
integer-weight convolution / GEMM kernels.  With autograd the DoReFa straight-through graph is kept: the
differentiable prefix (tanh, /max, clamp) is ordinary autograd and only the rounding step is our kernel with an
identity backward, exactly the reference's ``uniform_quantize`` (QU:8-27).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


def uniform_quantize(k):
    """QU:8-27: returns the ``apply`` of an autograd.Function: identity (k = 32), sign (k = 1), else
    round(x * n) / n with n = 2^k - 1; backward is the straight-through identity."""

    class qfn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input):
            if k == 32:
                return input
            ops._lib.require_cuda(input)
            return ops.uniform_quantize(input, k)

        @staticmethod
        def backward(ctx, grad_output):
            return grad_output.clone()

    return qfn.apply


class weight_quantize_fn(nn.Module):
    """QU:30-56."""

    def __init__(self, w_bit):
        super().__init__()
        assert w_bit <= 8 or w_bit == 32
        self.w_bit = w_bit
        self.uniform_q = uniform_quantize(k=w_bit - 1)

    def forward(self, x):
        if self.w_bit == 32:
            return x
        if self.w_bit == 1:
            E = torch.mean(torch.abs(x)).detach()
            return (self.uniform_q(x / E) + 1) / 2 * E        # upstream quirk: uniform_quantize(k=0) -> NaN
        if not (torch.is_grad_enabled() and x.requires_grad):
            codes = ops.ultra_weight_codes(x, self.w_bit)     # fused tanh / max / round kernels
            return self.values_from_codes(codes)
        weight = torch.tanh(x)
        weight = weight / torch.max(torch.abs(weight))
        return self.uniform_q(weight)

    def values_from_codes(self, codes: torch.Tensor) -> torch.Tensor:
        """codes / (2^(b-1)-1): bit-for-bit round(v*n)/n of QU:18-19 (sign() at w_bit == 2)."""
        # tensor / tensor is an IEEE division on CUDA (tensor / python-scalar is rewritten to a reciprocal multiply)
        n = torch.full((), float(2 ** (self.w_bit - 1) - 1), dtype=torch.float32, device=codes.device)
        return codes.to(torch.float32) / n


class activation_quantize_fn(nn.Module):
    """QU:59-73."""

    def __init__(self, a_bit):
        super().__init__()
        assert a_bit <= 8 or a_bit == 32
        self.a_bit = a_bit
        self.uniform_q = uniform_quantize(k=a_bit)

    def forward(self, x):
        if self.a_bit == 32:
            return x
        if not (torch.is_grad_enabled() and x.requires_grad):
            ops._lib.require_cuda(x)
            return ops.ultra_act(x, self.a_bit, want_codes=False, want_values=True)[1]
        return self.uniform_q(torch.clamp(x, 0, 1))


class _WeightCodeCache:
    """int8 weight codes of a Conv2d_Q / Linear_Q, recomputed when the weight Parameter changes."""

    def __init__(self):
        self.key = None
        self.codes = None

    def get(self, weight: torch.Tensor, w_bit: int) -> torch.Tensor:
        key = (weight.data_ptr(), weight._version, tuple(weight.shape), str(weight.device))
        if key != self.key:
            self.codes = ops.ultra_weight_codes(weight.detach(), w_bit)
            self.key = key
        return self.codes


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def conv2d_Q_fn(w_bit):
    """QU:76-91."""

    class Conv2d_Q(nn.Conv2d):
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
            super(Conv2d_Q, self).__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
            self.w_bit = w_bit
            self.quantize_fn = weight_quantize_fn(w_bit=w_bit)
            self.__dict__["_wcache"] = _WeightCodeCache()

        def weight_codes(self) -> torch.Tensor:
            c = self.__dict__.get("_wcache")
            if c is None:
                c = self.__dict__["_wcache"] = _WeightCodeCache()
            return c.get(self.weight, self.w_bit)

        def train(self, mode: bool = True):
            self.__dict__["_wcache"] = _WeightCodeCache()
            return super(Conv2d_Q, self).train(mode)

        def forward(self, input, order=None):
            ops._lib.require_cuda(input, self.weight)
            needs_grad = torch.is_grad_enabled() and (input.requires_grad or self.weight.requires_grad)
            if (not needs_grad and 2 <= self.w_bit <= 8 and self.groups == 1 and input.dim() == 4
                    and input.dtype == torch.float32 and not isinstance(self.padding, str)
                    and self.in_channels * self.kernel_size[0] * self.kernel_size[1] <= 12288):
                # the input is NOT quantised here (QU:85-89): fp32 activations x integer weight codes
                return ops.conv2d_f32_wcodes(input, self.weight_codes(), float(2 ** (self.w_bit - 1) - 1), self.bias,
                                             _pair(self.stride), _pair(self.padding), _pair(self.dilation))
            weight_q = self.quantize_fn(self.weight)
            return F.conv2d(input, weight_q, self.bias, self.stride, self.padding, self.dilation, self.groups)

    return Conv2d_Q


def _folded_bn(mod):
    """(w, b) of QU:106-111 / QU:190-195: w = gamma / (sqrt(var) + eps), b = beta - mean / (sqrt(var) + eps) * gamma
    (eps OUTSIDE the sqrt), computed by the K5 fold kernel."""
    return ops.bn_fold(mod.weight.detach(), mod.bias.detach(), mod.running_mean, mod.running_var, mod.eps, mode=1)


def batchNorm2d_Q_fn(w_bit):
    """QU:94-132.  The fold is clamped to [-1, 1], mapped to [0, 1], quantised on the 2^w_bit-1 grid and mapped back; the
    layer then applies y = x * w_q + b_q (upstream calls F.batch_norm with zero mean, unit variance and eps = 0; torch >= 2
    put a Python-level `eps <= 0` guard in front of the unchanged ATen op, so the golden outputs this class is tested
    against were produced by the reference class with that guard bypassed - oracle/make_golden.py::golden_bnq).
    Unused by mymodel.py (it uses nn.BatchNorm2d, MM:74)."""

    class BatchNorm2d_Q(nn.BatchNorm2d):
        def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
            super(BatchNorm2d_Q, self).__init__(num_features, eps, momentum, affine, track_running_stats)
            self.w_bit = w_bit
            self.quantize_fn = uniform_quantize(k=w_bit)

        def folded_quantized(self):
            w, b = _folded_bn(self)
            w_q = 2 * self.quantize_fn(torch.clamp(w, -1, 1) / 2 + 0.5) - 1
            b_q = 2 * self.quantize_fn(torch.clamp(b, -1, 1) / 2 + 0.5) - 1
            return w_q, b_q

        def forward(self, input):
            ops._lib.require_cuda(input, self.weight)
            w_q, b_q = self.folded_quantized()
            shape = (1, -1) + (1,) * (input.dim() - 2)
            return input * w_q.view(shape) + b_q.view(shape)

    return BatchNorm2d_Q


def batchNorm1d_Q_fn(w_bit):
    """QU:135-207: the folded scale is quantised directly (no range mapping), the folded bias is NOT quantised
    (QU:190-206).  Same eps = 0 caveat as BatchNorm2d_Q."""

    class BatchNorm1d_Q(nn.BatchNorm1d):
        def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
            super(BatchNorm1d_Q, self).__init__(num_features, eps, momentum, affine, track_running_stats)
            self.w_bit = w_bit
            self.quantize_fn = uniform_quantize(k=w_bit)

        def forward(self, input):
            self._check_input_dim(input)
            ops._lib.require_cuda(input, self.weight)
            w, b = _folded_bn(self)
            _ = self.quantize_fn(w)           # upstream computes w_q and then applies the UN-quantised w (QU:196, 203)
            shape = (1, -1) + (1,) * (input.dim() - 2)
            return input * w.view(shape) + b.view(shape)

    return BatchNorm1d_Q


def linear_Q_fn(w_bit):
    """QU:210-222."""

    class Linear_Q(nn.Linear):
        def __init__(self, in_features, out_features, bias=True):
            super(Linear_Q, self).__init__(in_features, out_features, bias)
            self.w_bit = w_bit
            self.quantize_fn = weight_quantize_fn(w_bit=w_bit)
            self.__dict__["_wcache"] = _WeightCodeCache()

        def weight_codes(self) -> torch.Tensor:
            c = self.__dict__.get("_wcache")
            if c is None:
                c = self.__dict__["_wcache"] = _WeightCodeCache()
            return c.get(self.weight, self.w_bit)

        def train(self, mode: bool = True):
            self.__dict__["_wcache"] = _WeightCodeCache()
            return super(Linear_Q, self).train(mode)

        def forward(self, input):
            ops._lib.require_cuda(input, self.weight)
            needs_grad = torch.is_grad_enabled() and (input.requires_grad or self.weight.requires_grad)
            if not needs_grad and 2 <= self.w_bit <= 8 and input.dtype == torch.float32:
                # a linear layer is a 1x1 convolution over a [rows, K, 1, 1] map: same integer-weight kernel
                x2 = input.reshape(-1, self.in_features, 1, 1)
                w4 = self.weight_codes().reshape(self.out_features, self.in_features, 1, 1)
                if self.in_features <= 12288:
                    y = ops.conv2d_f32_wcodes(x2, w4, float(2 ** (self.w_bit - 1) - 1), self.bias, (1, 1), (0, 0), (1, 1))
                    return y.reshape(*input.shape[:-1], self.out_features)
            weight_q = self.quantize_fn(self.weight)
            return F.linear(input, weight_q, self.bias)

    return Linear_Q
