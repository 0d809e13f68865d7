"""Oracle (TEST INFRASTRUCTURE) for GETA's learned-step symmetric quantizers and the quantized
Linear/Conv2d layers built on them.

Restates, with stock PyTorch-CPU fp32 ops:
  * SymQuantizerLinear.forward / backward        QL:136-161 / QL:163-205
  * SymQuantizerNonLinear.forward / backward     QL:40-69   / QL:71-125
  * DGEQuantizer.backward (grad_x factor)        QL:248-290
  * initialize_quant_layer                       QL:413-440
  * QuantizeMixin.weight_bit / activation_bit    QL:383-410
  * QuantizeLinear.forward / QuantizeConv2d.forward   QL:495-499 / QL:575-587

Conventions: ``d``, ``q_m``, ``t`` are fp32 tensors of shape (1,) (the reference's nn.Parameters,
QL:315-325); ``q_s`` is always 0 in the reference (QL:335, QL:362) and is kept only in comments.

Integer view: the reference output is ``sign(x) * d * round(p / d)`` with
``p = |x|`` (linear) or ``exp(t*log|x|)`` (non-linear), saturated to ``round(r / d)``,
``r = |q_m|`` or ``exp(t*log(|q_m|+1e-6))``, wherever ``|x| >= q_m`` (signed compare).  Hence
``value == code * |d|`` with ``code = sign(x) * |round(p/d)|`` - that integer is what the CUDA
kernels must reproduce bit-exactly.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

_EPS_RANGE = 1e-6  # QL:62 / QL:82: "+ 1e-6" inside the log of the range


def _as_param(v) -> torch.Tensor:
    if isinstance(v, torch.Tensor):
        return v.detach().to(torch.float32).reshape(-1)[:1].cpu()
    return torch.tensor([float(v)], dtype=torch.float32)


def _domain(x_abs: torch.Tensor, q_m: torch.Tensor, t: Optional[torch.Tensor]):
    """(p, r): magnitude of the input and of the range in the quantizer's domain.
    linear QL:154-155; non-linear QL:62-63."""
    if t is None:
        return x_abs, torch.abs(q_m)
    r = torch.exp(t * torch.log(torch.abs(q_m) + _EPS_RANGE))
    p = torch.exp(t * torch.log(x_abs))
    return p, r


def sym_forward(x: torch.Tensor, d, q_m, t=None) -> torch.Tensor:
    """Fake-quantized values, op-for-op the reference forward (QL:146-161 linear, QL:50-69 non-linear)."""
    x = x.detach().to(torch.float32).cpu()
    d, q_m = _as_param(d), _as_param(q_m)
    t = None if t is None else _as_param(t)
    x_abs = torch.abs(x)
    p, r = _domain(x_abs, q_m, t)
    out = d * torch.round(p.div(d))                    # QL:157 / QL:65
    out[x_abs <= 0.0] = 0                              # QL:158 / QL:66 (q_s == 0)
    out[x_abs >= q_m] = d * torch.round(r.div(d))      # QL:159 / QL:67 (signed q_m in the compare)
    return torch.sign(x) * out                         # QL:160 / QL:68


def sym_codes(x: torch.Tensor, d, q_m, t=None) -> torch.Tensor:
    """Signed integer codes (int64) such that ``sym_forward(x) == codes * |d|`` (NaN inputs excluded).
    Derived from the same fp32 quotient/round as the reference, not from the values."""
    x = x.detach().to(torch.float32).cpu()
    d, q_m = _as_param(d), _as_param(q_m)
    t = None if t is None else _as_param(t)
    x_abs = torch.abs(x)
    p, r = _domain(x_abs, q_m, t)
    k = torch.abs(torch.round(p.div(d)))
    k_sat = torch.abs(torch.round(r.div(d)))
    k = torch.where(x_abs <= 0.0, torch.zeros_like(k), k)
    k = torch.where(x_abs >= q_m, k_sat.expand_as(k), k)
    k = torch.sign(x) * k
    return torch.nan_to_num(k, nan=0.0, posinf=0.0, neginf=0.0).to(torch.int64)


def saturation_code(d, q_m, t=None) -> int:
    """``round(r/d)``: the largest code magnitude the quantizer can emit (QL:159 / QL:67)."""
    d, q_m = _as_param(d), _as_param(q_m)
    t = None if t is None else _as_param(t)
    _, r = _domain(torch.zeros(1), q_m, t)
    v = torch.abs(torch.round(r.div(d))).item()
    return int(v) if math.isfinite(v) else -1


def sym_backward(x: torch.Tensor, g: torch.Tensor, d, q_m, t=None,
                 clip: Tuple[float, float] = (-2.0, 2.0)) -> Dict[str, torch.Tensor]:
    """Gradients of the reference autograd.Functions.

    linear QL:163-205 -> grad_x, grad_d, grad_qm ; non-linear QL:71-125 -> + grad_t.
    ``clip`` is the module attribute weight_clip_val / act_clip_val (QL:311-312)."""
    x = x.detach().to(torch.float32).cpu()
    g = g.detach().to(torch.float32).cpu()
    d, q_m = _as_param(d), _as_param(q_m)
    t = None if t is None else _as_param(t)
    x_abs = torch.abs(x)
    sgn = torch.sign(x)

    grad_x = g.clone()                                  # QL:169-171 / QL:77-79: STE, zero outside clip
    grad_x[x.ge(clip[1])] = 0
    grad_x[x.le(clip[0])] = 0

    p, r = _domain(x_abs, q_m, t)
    resid = torch.round(p.div(d)) - p.div(d)            # QL:177 / QL:89
    resid[x_abs >= q_m] = torch.round(r.div(d)) - r.div(d)   # QL:178-180 / QL:90-92
    resid[x_abs <= 0.0] = 0                             # QL:181 / QL:93
    grad_d = torch.sum(g * (sgn * resid)).reshape(1)    # QL:182-183 / QL:94-95

    if t is None:
        dqm = sgn.clone()                               # QL:185-187
        dqm[x_abs <= q_m] = 0
        grad_qm = torch.sum(g * dqm).reshape(1)
        return {"grad_x": grad_x, "grad_d": grad_d, "grad_qm": grad_qm}

    r_low = torch.exp((t - 1) * torch.log(torch.abs(q_m) + _EPS_RANGE))    # QL:84-86
    dqm = sgn * (t * r_low).expand_as(x)                # QL:97-99
    dqm[x_abs <= q_m] = 0
    grad_qm = torch.sum(g * dqm).reshape(1)

    dt = p * torch.log(x_abs)                           # QL:101-105
    dt[x_abs >= q_m] = r * torch.log(torch.abs(q_m) + _EPS_RANGE)
    dt[x_abs <= 0.0] = 0
    grad_t = torch.sum(g * (sgn * dt)).reshape(1)
    return {"grad_x": grad_x, "grad_d": grad_d, "grad_qm": grad_qm, "grad_t": grad_t}


def dge_grad_x(x: torch.Tensor, g: torch.Tensor, d, num_bits: float,
               clip: Tuple[float, float] = (-2.0, 2.0)) -> torch.Tensor:
    """DGEQuantizer.backward's input gradient (QL:253-265): STE-clip, times
    (1/k)|x - d/2|^(1/k-1) with k = 5*4/num_bits (QL:236), clamped to +-3."""
    x = x.detach().to(torch.float32).cpu()
    g = g.detach().to(torch.float32).cpu()
    d = _as_param(d)
    k = torch.tensor(5.0 * (4.0 / num_bits))
    grad_x = g.clone()
    grad_x[x.ge(clip[1])] = 0
    grad_x[x.le(clip[0])] = 0
    scale = (1 / k) * torch.pow(torch.abs(x - d / 2), 1 / k - 1)
    return torch.clamp(grad_x * scale, -3.0, 3.0)


def init_quant_params(weight: torch.Tensor, num_bits: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(d, q_m) at conversion time: q_m = max|W|, d = q_m / (2^(b-1) - 1)  (QL:423-427).
    The activation quantizer gets the SAME two values (QL:436-437)."""
    q_m = torch.max(torch.abs(weight.detach().to(torch.float32).cpu()))
    d = (q_m - torch.tensor(0.0)) / (2 ** (float(num_bits) - 1) - 1)
    return d.reshape(1), q_m.reshape(1)


def bit_width(d: float, q_m: float, t: float = 1.0) -> float:
    """log2(|q_m|^t / |d| + 1) + 1  (QL:394, QM:118-119); the module properties round() it."""
    return math.log2(math.exp(t * math.log(abs(q_m))) / abs(d) + 1) + 1


# --------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------

def quantize_linear_forward(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                            wq: Dict[str, torch.Tensor], aq: Optional[Dict[str, torch.Tensor]] = None,
                            want_int: bool = True) -> Dict[str, torch.Tensor]:
    """QuantizeLinear.forward (QL:495-499).  ``wq``/``aq`` = {"d":..,"q_m":..[,"t":..]};
    ``aq is None`` <=> WEIGHT_ONLY mode (QL:497).

    Returns y (fp32, = F.linear on the fake-quant values, exactly what the reference computes) and,
    if ``want_int``, the integer codes and the exact int64 accumulators K_a @ K_w^T."""
    x = x.detach().to(torch.float32).cpu()
    w_q = sym_forward(weight, wq["d"], wq["q_m"], wq.get("t"))
    x_q = x if aq is None else sym_forward(x, aq["d"], aq["q_m"], aq.get("t"))
    out = {"y": F.linear(x_q, w_q, None if bias is None else bias.detach().float().cpu()),
           "w_q": w_q, "x_q": x_q}
    if want_int:
        out["w_codes"] = sym_codes(weight, wq["d"], wq["q_m"], wq.get("t"))
        if aq is not None:
            out["a_codes"] = sym_codes(x, aq["d"], aq["q_m"], aq.get("t"))
            a2 = out["a_codes"].reshape(-1, x.shape[-1])
            out["acc"] = (a2 @ out["w_codes"].t()).reshape(*x.shape[:-1], weight.shape[0])
    return out


def quantize_conv2d_forward(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                            wq: Dict[str, torch.Tensor], aq: Optional[Dict[str, torch.Tensor]] = None,
                            stride=1, padding=1, dilation=1, groups=1,
                            want_int: bool = True) -> Dict[str, torch.Tensor]:
    """QuantizeConv2d.forward (QL:575-587).  Integer accumulators via unfold (im2col with
    K ordered (c, kh, kw) = weight.reshape(O, -1); zero padding contributes code 0)."""
    x = x.detach().to(torch.float32).cpu()
    w_q = sym_forward(weight, wq["d"], wq["q_m"], wq.get("t"))
    x_q = x if aq is None else sym_forward(x, aq["d"], aq["q_m"], aq.get("t"))
    b = None if bias is None else bias.detach().float().cpu()
    out = {"y": F.conv2d(x_q, w_q, b, stride, padding, dilation, groups), "w_q": w_q, "x_q": x_q}
    if want_int:
        out["w_codes"] = sym_codes(weight, wq["d"], wq["q_m"], wq.get("t"))
        if aq is not None:
            out["a_codes"] = sym_codes(x, aq["d"], aq["q_m"], aq.get("t"))
            if groups == 1:
                out["acc"] = int_conv2d(out["a_codes"], out["w_codes"], stride, padding, dilation)
    return out


def int_conv2d(a_codes: torch.Tensor, w_codes: torch.Tensor, stride=1, padding=0, dilation=1) -> torch.Tensor:
    """Exact integer convolution (groups == 1) on int64 codes via im2col: [B, O, OH, OW] int64."""
    B, C, H, W = a_codes.shape
    O, _, kh, kw = w_codes.shape
    cols = F.unfold(a_codes.to(torch.float64), (kh, kw), dilation=dilation, padding=padding, stride=stride)
    cols = cols.to(torch.int64)                                       # [B, C*kh*kw, L]
    acc = torch.einsum("ok,bkl->bol", w_codes.reshape(O, -1), cols)    # int64 exact
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    ph, pw = (padding, padding) if isinstance(padding, int) else padding
    dh, dw = (dilation, dilation) if isinstance(dilation, int) else dilation
    OH = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    OW = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    return acc.reshape(B, O, OH, OW)


# --------------------------------------------------------------------------------------------
# autograd view of the quantizers (QAT parity): the oracle's forward / backward restatements above wired
# into torch.autograd so that a whole model can be differentiated exactly the way the reference's
# autograd.Functions do it (QL:33-205), without the reference's classes being present.
# --------------------------------------------------------------------------------------------

class SymQuantFn(torch.autograd.Function):
    """forward = sym_forward, backward = sym_backward (STE inside clip, fused reductions for d / q_m / t)."""

    @staticmethod
    def forward(ctx, x, d, q_m, t, clip_lo, clip_hi):
        ctx.save_for_backward(x, d, q_m) if t is None else ctx.save_for_backward(x, d, q_m, t)
        ctx.nl, ctx.clip = t is not None, (clip_lo, clip_hi)
        return sym_forward(x, d, q_m, t)

    @staticmethod
    def backward(ctx, g):
        if ctx.nl:
            x, d, q_m, t = ctx.saved_tensors
        else:
            (x, d, q_m), t = ctx.saved_tensors, None
        r = sym_backward(x, g, d, q_m, t, ctx.clip)
        return r["grad_x"], r["grad_d"], r["grad_qm"], r.get("grad_t"), None, None


def quantize_linear_autograd(x, weight, bias, wq, aq, clip=(-2.0, 2.0)):
    """QuantizeLinear.forward (QL:495-499) under autograd: F.linear(Q_a(x), Q_w(W), b)."""
    w_q = SymQuantFn.apply(weight, wq["d"], wq["q_m"], wq.get("t"), clip[0], clip[1])
    x_q = x if aq is None else SymQuantFn.apply(x, aq["d"], aq["q_m"], aq.get("t"), clip[0], clip[1])
    return F.linear(x_q, w_q, bias)


def quantize_conv2d_autograd(x, weight, bias, wq, aq, stride=1, padding=1, dilation=1, groups=1, clip=(-2.0, 2.0)):
    """QuantizeConv2d.forward (QL:575-587) under autograd."""
    w_q = SymQuantFn.apply(weight, wq["d"], wq["q_m"], wq.get("t"), clip[0], clip[1])
    x_q = x if aq is None else SymQuantFn.apply(x, aq["d"], aq["q_m"], aq.get("t"), clip[0], clip[1])
    return F.conv2d(x_q, w_q, bias, stride, padding, dilation, groups)
