"""Drop-in replacements for the reference's GETA quantized layers, running on hand-written sm_100a kernels.

Mirrors ``QViT_with_GETA/only_train_once/quantization/quant_layers.py`` (reference file:line cited per symbol):
same class names, constructor arguments, parameter names, ``state_dict`` layout, enums and error
behaviour, so ``model_to_quantize_model`` / GETA / checkpoints keep working unchanged.  What changes is what runs:

* inference (no autograd): activations -> int8 codes (``qvit_quantize_sym``), weights -> int8 codes cached per
  parameter version, exact integer contraction on the tcgen05 ``kind::i8`` pipe with TMEM int32 accumulators and
  a fused dequant + bias epilogue (``qvit_gemm_i8``).  The reference runs ~11 elementwise ATen kernels per
  quantizer and an fp32 GEMM on fake-quant values (QL:495-499).
* training (autograd): ``QuantLinearFunction`` - exact int8 tensor-core forward, the two gradient GEMMs on tcgen05
  ``kind::f16`` with the fp32 gradient split exactly into three bf16 planes (the reference never quantizes
  ``grad_output``), ONE fused kernel per quantizer for the STE mask and the step-size / range /
  exponent reductions (``qvit_sym_backward``).  QuantizeConv2d trains through the same Functions on the im2col matrix.
* configurations the int8 pipe cannot carry (> 8-bit codes - where GETA starts, train.py:247-250 - and weight-only mode)
  run ``WideQuantLinearFunction`` / ``ops.matmul_f32_tc``: fp32-equivalent tensor-core GEMMs (both operands as three
  exact bf16 planes).  A sync-free monitor (``_DispatchMonitor``) moves a layer between the two paths as GETA walks its
  bit width.  Only grouped convolutions fall back to the reference op chain on library kernels.  There is no CPU path.
"""
from __future__ import annotations

import logging
import math
from enum import Enum
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

logger = logging.getLogger(__name__)


class NanInGradientError(Exception):
    """QL:10-13."""

    def __init__(self, message):
        self.message = message
        super().__init__(self.message)


class QuantizationType(Enum):          # QL:20-24
    SYMMETRIC_LINEAR = "symmetric+linear"
    SYMMETRIC_NONLINEAR = "symmetric+nonlinear"
    DGE = "dge"


class QuantizationMode(Enum):          # QL:27-29
    WEIGHT_ONLY = "weight_only"
    WEIGHT_AND_ACTIVATION = "weight_and_activation"


# ---------------------------------------------------------------------------------------------------
# NaN-in-gradient reporting without a per-layer host sync: kernels OR a bit into a per-device flag word; it
# is polled once per step by check_nan_flags() (or eagerly when QVIT_EAGER_NAN_CHECK is set by a test).
# ---------------------------------------------------------------------------------------------------
_flag_words = {}
EAGER_NAN_CHECK = False
TENSOR_CORE_BACKWARD = True     # QAT gradient GEMMs on tcgen05 (exact bf16 split); False = library fp32 GEMMs
MN_MAJOR_WEIGHT_GRADIENT = True   # grad_w = g^T x_q from the row planes / codes as MN-major operands (no transposed copies)
GRADIENT_PLANES = 3             # bf16 planes of the gradient operand: 3 = exact fp32 (default), 2 = 16 significant bits
                                # (relative error <= 2^-17 per element, ~1e-5 on the result, a third less tensor-core work)


def _flags_for(device: torch.device) -> torch.Tensor:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    w = _flag_words.get(key)
    if w is None:
        w = ops.new_flags(device)
        _flag_words[key] = w
    return w


def check_nan_flags(raise_error: bool = True) -> int:
    """Poll (one host sync) and clear the device flag words.  Raises NanInGradientError if a fused backward saw a
    NaN in the reduced step-size gradient - the deferred equivalent of QL:189-204 / QL:107-123."""
    bits = 0
    for w in _flag_words.values():
        bits |= int(w.item())                     # host sync: every asynchronous read-back below has landed as well
        w.zero_()
    for mon in _monitors.values():
        bits |= mon.drain()
    if raise_error and bits & ops._lib.QVIT_FLAG_NAN_GRAD:
        raise NanInGradientError("Error: NaN appears in gradient! (reported by the fused quantizer backward)")
    if raise_error and bits & ops._lib.QVIT_FLAG_OVERFLOW:
        raise RuntimeError("a quantizer produced codes beyond +-127 on the int8 path (d_quant shrank below q_m/127): "
                           "call module.invalidate_quant_cache() / re-create the modules so the wide path is selected")
    return bits


class _DispatchMonitor:
    """Sync-free, one-step-late host view of (a) the saturation codes round(r/d) of every registered layer's weight and
    activation quantizer and (b) the device flag word.

    GETA moves d_quant / q_m / t_quant every step through raw ``.data`` writes (geta.py:571-772) and walks the bit width
    from the conversion value (32 in train.py:247-250) down to ~4-8 bits (geta.py:895-900).  Whether a layer's codes fit the
    int8 tensor-core pipe therefore changes DURING training, and asking the device each step would synchronise every
    layer.  Instead, once per training step (detected as a layer running again) one tiny kernel (``qvit_quant_sat_levels``) evaluates all saturation
    codes through a pointer table, the result and the flag word are copied to pinned memory asynchronously, and the copy is
    consumed by whichever forward() first finds its event complete.  A layer takes the int8 path only while both codes
    are <= MARGIN (120 < 127): one step of lag cannot push a code past 127, so the int8 kernels never have to clamp; above
    the margin the layer runs the wide path, which is exact for any bit width.  NaN-in-gradient bits found in the flag
    word raise NanInGradientError from that forward (the deferred equivalent of QL:189-204), so an unchanged training loop
    sees the error without calling check_nan_flags()."""
    MARGIN = 120.0

    def __init__(self, device: torch.device):
        import weakref
        self._weakref = weakref
        self.device = device
        self.mods = []                 # weak references, slot = index
        self.sat = []                  # [2 * n] python floats (weight, activation), lagged
        self.seen = set()              # slots that ran since the last refresh was launched
        self.n_live = 0                # registered modules still alive
        self.pending = None            # CUDA event of the read-back in flight
        self.key = None
        self.ptab = self.dev_sat = self.host_sat = None
        self.host_flags = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.sticky = 0

    def register(self, mod) -> int:
        slot = len(self.mods)
        self.mods.append(self._weakref.ref(mod, self._on_dead))
        self.n_live += 1
        sw = QuantizeMixin._sat_level(*mod._wt_qparams())             # one host read per quantizer, once per module
        sa = QuantizeMixin._sat_level(*mod._act_qparams())
        self.sat += [sw, sa]
        self.key = None
        return slot

    def _on_dead(self, _ref):
        self.n_live -= 1

    def _ptr_key(self):
        ptrs = []
        for r in self.mods:
            m = r()
            ps = (None,) * 6 if m is None else (*m._wt_qparams(), *m._act_qparams())
            ptrs += [0 if p is None else p.data_ptr() for p in ps]
        return tuple(ptrs)

    def _launch(self):
        n = len(self.mods)
        key = self._ptr_key()
        if key != self.key:                                            # parameters re-created (pruning) / new modules
            self.ptab = torch.tensor(key, dtype=torch.int64).to(self.device)
            self.dev_sat = torch.empty(2 * n, dtype=torch.float32, device=self.device)
            self.host_sat = torch.empty(2 * n, dtype=torch.float32).pin_memory()
            self.key = key
        ops._lib.check(ops._lib.lib().qvit_quant_sat_levels(self.ptab.data_ptr(), n, self.dev_sat.data_ptr(), ops._lib.stream()),
                       "qvit_quant_sat_levels")
        self.host_sat.copy_(self.dev_sat, non_blocking=True)
        w = _flags_for(self.device)
        self.host_flags.copy_(w, non_blocking=True)
        w.zero_()                                                      # stream-ordered after the copy: later bits are kept
        self.pending = torch.cuda.Event()
        self.pending.record()

    def _consume(self) -> int:
        self.pending = None
        vals = self.host_sat.tolist()
        if len(vals) == len(self.sat):
            self.sat = [v if v >= 0 else s for v, s in zip(vals, self.sat)]
        bits = int(self.host_flags[0])
        self.host_flags[0] = 0
        self.sticky |= bits
        return bits

    def drain(self) -> int:
        """After a host sync: fold a landed read-back in and hand the collected flag bits over."""
        if self.pending is not None and self.pending.query():
            self._consume()
        bits, self.sticky = self.sticky, 0
        return bits

    def tick(self, slot: int = -1):
        if torch.cuda.is_current_stream_capturing():
            return
        if self.pending is not None and self.pending.query():
            bits = self._consume()
            if bits & ops._lib.QVIT_FLAG_OVERFLOW:
                logger.warning("a quantizer's saturation code moved past 127 within one step; the affected int8 codes were "
                               "clamped for that step (the layer is on the wide path from now on)")
            if bits & ops._lib.QVIT_FLAG_NAN_GRAD:
                self.sticky &= ~ops._lib.QVIT_FLAG_NAN_GRAD
                raise NanInGradientError("Error: NaN appears in gradient! (reported by the fused quantizer backward)")
        # a refresh is due once per training step: a step boundary is a layer calling again before the read-back was renewed
        # (robust against modules that are registered but not part of the running model)
        if slot in self.seen:
            self.seen.clear()
            if self.pending is None:
                self._launch()
        self.seen.add(slot)

    def int8_ok(self, mod) -> bool:
        slot = mod.__dict__.get("_mon_slot")
        if slot is None or slot[0] is not self or self.mods[slot[1]]() is not mod:
            slot = mod.__dict__["_mon_slot"] = (self, self.register(mod))
        self.tick(slot[1])
        i = slot[1]
        return self.sat[2 * i] <= self.MARGIN and 0 <= self.sat[2 * i + 1] <= self.MARGIN


_monitors = {}


def _monitor_for(device: torch.device) -> _DispatchMonitor:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    m = _monitors.get(key)
    if m is None:
        m = _monitors[key] = _DispatchMonitor(torch.device(*key))
    return m


def _clip_pair(clip_val) -> Tuple[float, float]:
    if isinstance(clip_val, torch.Tensor):
        lo, hi = clip_val.detach().flatten().tolist()[:2]      # reference API passes a tensor (QL:334); host read
        return float(lo), float(hi)
    return float(clip_val[0]), float(clip_val[1])


def _sym_forward(ctx, input, d_quant, q_m, t_quant, clip_val):
    dev = input.device
    ops._lib.require_cuda(input)
    d_quant, q_m = d_quant.to(dev), q_m.to(dev)                 # QL:148-151
    t_quant = None if t_quant is None else t_quant.to(dev)
    x = input.contiguous()
    ctx.clip = _clip_pair(clip_val)
    ctx.save_for_backward(x, d_quant, q_m) if t_quant is None else ctx.save_for_backward(x, d_quant, q_m, t_quant)
    return ops.fake_quantize_sym(x, d_quant, q_m, t_quant)


def _sym_backward(ctx, grad_output, nonlinear: bool):
    if nonlinear:
        x, d_quant, q_m, t_quant = ctx.saved_tensors
    else:
        (x, d_quant, q_m), t_quant = ctx.saved_tensors, None
    flags = _flags_for(x.device)
    grad_x, s = ops.sym_backward(x, grad_output.contiguous(), d_quant, q_m, t_quant, ctx.clip,
                                 want_grad_x=ctx.needs_input_grad[0], flags=flags)
    if EAGER_NAN_CHECK:
        check_nan_flags()
    return grad_x, s[0:1], s[1:2], s[2:3]


class SymQuantizerNonLinear(torch.autograd.Function):
    """QL:33-125.  forward: sign(x) * d * round(exp(t*log|x|) / d) with zero / saturation overrides;
    backward: STE inside clip_val + fused reductions for d, q_m, t."""

    @staticmethod
    def forward(ctx, input, d_quant, q_m, t_quant, clip_val, q_s):
        return _sym_forward(ctx, input, d_quant, q_m, t_quant, clip_val)

    @staticmethod
    def backward(ctx, grad_output):
        gx, gd, gq, gt = _sym_backward(ctx, grad_output, True)
        return gx, gd, gq, gt, None, None


class SymQuantizerLinear(torch.autograd.Function):
    """QL:128-205."""

    @staticmethod
    def forward(ctx, input, d_quant, q_m, clip_val, q_s):
        return _sym_forward(ctx, input, d_quant, q_m, None, clip_val)

    @staticmethod
    def backward(ctx, grad_output):
        gx, gd, gq, _ = _sym_backward(ctx, grad_output, False)
        return gx, gd, gq, None, None


class DGEQuantizer(torch.autograd.Function):
    """QL:207-290 ("experimental WIP, not used" upstream).  Forward = the linear quantizer kernel; backward reuses
    the fused kernel for the d / q_m reductions and applies the DGE factor (1/k)|x - d/2|^(1/k-1), clamp +-3."""

    @staticmethod
    def forward(ctx, input, d_quant, q_m, clip_val, q_s, num_bits):
        ctx.k = 5.0 * (4.0 / float(num_bits))                    # QL:236
        return _sym_forward(ctx, input, d_quant, q_m, None, clip_val)

    @staticmethod
    def backward(ctx, grad_output):
        x, d_quant, q_m = ctx.saved_tensors
        gx, gd, gq, _ = _sym_backward(ctx, grad_output, False)
        k = ctx.k
        scale = (1.0 / k) * torch.pow(torch.abs(x - d_quant / 2), 1.0 / k - 1.0)     # QL:259-261
        gx = torch.clamp(gx * scale, -3.0, 3.0)                                     # QL:262-265
        if EAGER_NAN_CHECK and torch.isnan(gx).any():
            raise NanInGradientError("NaN in gradient computation")
        return gx, gd, gq, None, None, None


class QuantLinearFunction(torch.autograd.Function):
    """Fused QAT step of QuantizeLinear (QL:495-499 forward; QL:163-205 / 71-125 backward for both quantizers).

    forward : activation and weight codes (K2/K1) -> exact int8 tensor-core GEMM -> fp32 dequant + bias (K3).  This equals
              F.linear(x_q, w_q, bias) of the reference up to fp32 rounding, without materialising x_q / w_q.
    backward: grad_x_q = g @ w_q and grad_w_q = g^T @ x_q: the gradient operand is never quantized upstream, so it is split
              exactly into three bf16 planes and multiplied with the integer codes on tcgen05 kind::f16 (fp32 accumulation,
              qvit_gemm_bf16_split); then ONE fused kernel per quantizer for the STE mask and the step-size /
              range / exponent gradient reductions (K6).  Saved for backward: x, W and the int8 codes (1 B/element)."""

    @staticmethod
    def forward(ctx, x, weight, bias, d_a, qm_a, t_a, d_w, qm_w, t_w, clip_a, clip_w, *extra):
        pre_gelu = bool(extra[0]) if extra else False
        residual = extra[1] if len(extra) > 1 else None        # fp32 [..., N]: added in the GEMM epilogue (Block.forward's x + ...)
        ctx.n_extra, ctx.has_res = len(extra), residual is not None
        K, N = weight.shape[1], weight.shape[0]
        x2 = x.reshape(-1, K).contiguous()
        flags = _flags_for(x.device)
        # pre_gelu: x is the PRE-activation of the nn.GELU in front of this layer (Mlp.forward, vit_model.py:173): the quantizer
        # kernel applies it, the backward kernel recomputes it - the fp32 activation and its gradient never exist in HBM
        ctx.pre_gelu = bool(pre_gelu)
        a_codes = ops.quantize_sym(x2, d_a, qm_a, t_a, flags=flags, gelu=True) if ctx.pre_gelu else \
            ops.quantize_sym(x2, d_a, qm_a, t_a, ld_codes=ops.pad16(K), flags=flags)
        w_codes = ops.quantize_sym(weight.detach(), d_w, qm_w, t_w, ld_codes=ops.pad16(K), flags=flags)
        y = ops.gemm_i8(a_codes, w_codes, K, N, out_kind=ops.QVIT_OUT_F32, scale_a=d_a, scale_w=d_w,
                        bias=None if bias is None else bias.detach(), flags=flags,
                        residual=None if residual is None else residual.detach().reshape(-1, N))
        ctx.clip_a, ctx.clip_w, ctx.has_bias, ctx.nl = clip_a, clip_w, bias is not None, t_a is not None
        ctx.x_shape = x.shape
        saved = [x2, weight, d_a, qm_a, d_w, qm_w, a_codes, w_codes] + ([t_a, t_w] if t_a is not None else [])
        ctx.save_for_backward(*saved)
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, g):
        if ctx.nl:
            x2, weight, d_a, qm_a, d_w, qm_w, a_codes, w_codes, t_a, t_w = ctx.saved_tensors
        else:
            (x2, weight, d_a, qm_a, d_w, qm_w, a_codes, w_codes), t_a, t_w = ctx.saved_tensors, None, None
        K, N = weight.shape[1], weight.shape[0]
        g2 = g.reshape(-1, N).contiguous()
        flags = _flags_for(g.device)
        M = g2.shape[0]
        # the activation quantizer's scalar gradients (QL:177-187 / 89-105) are sums over grad_output of quantize_act whether
        # or not the layer INPUT needs a gradient (a first layer fed raw data still trains d_quant_act / q_m_act)
        need_act = any(ctx.needs_input_grad[i] for i in (0, 3, 4, 5))
        if TENSOR_CORE_BACKWARD and K % 4 == 0:
            # grad_x_q = g @ w_q = |d_w| * (g1 + g2 + g3) @ codes_w   and   grad_w_q = g^T @ x_q = |d_a| * (g^T planes) @ codes_a:
            # exact 3-way bf16 split of g, integer codes as bf16, tcgen05 kind::f16 with fp32 accumulation
            # (both plane forms of g and the bias gradient come from one pass over g: ops.grad_prep)
            if MN_MAJOR_WEIGHT_GRADIENT:
                # the weight-gradient GEMM reads the ROW planes of g and the codes as MN-major operands: no transposed copy of
                # either is made (the transposed planes were 6 of the 16 bytes per element grad_prep moved)
                g_rows, _, grad_b = ops.grad_prep(g2, want_rows=True, want_colsum=ctx.has_bias, want_trans=False)
                grad_wq = ops.gemm_bf16_split_t(g_rows, ops.codes_to_bf16(a_codes, K), N, K, planes=GRADIENT_PLANES, scale=d_a)
            else:
                g_rows, g_trans, grad_b = ops.grad_prep(g2, want_rows=need_act, want_colsum=ctx.has_bias)
                grad_wq = ops.gemm_bf16_split(g_trans, ops.codes_to_bf16_t(a_codes, K), M, planes=GRADIENT_PLANES, scale=d_a)
            grad_xq = ops.gemm_bf16_split(g_rows, ops.codes_to_bf16_t(w_codes, K), N, planes=GRADIENT_PLANES, scale=d_w) \
                if need_act else None
        else:
            # fake-quant values from the saved codes: value = code * |d| (exactly what the reference forward produced)
            x_q = a_codes[:, :K].to(torch.float32) * d_a.detach().abs()
            w_q = w_codes[:, :K].to(torch.float32) * d_w.detach().abs()
            grad_xq = g2 @ w_q if need_act else None
            grad_wq = g2.t() @ x_q
            grad_b = g2.sum(0) if ctx.has_bias else None
        if grad_xq is None:
            grad_x, s_a = None, torch.zeros(3, dtype=torch.float32, device=g.device)
        else:
            grad_x, s_a = ops.sym_backward(x2, grad_xq, d_a, qm_a, t_a, ctx.clip_a, want_grad_x=ctx.needs_input_grad[0], flags=flags,
                                           gelu=ctx.pre_gelu)
        grad_w, s_w = ops.sym_backward(weight.detach(), grad_wq, d_w, qm_w, t_w, ctx.clip_w, flags=flags)
        if EAGER_NAN_CHECK:
            check_nan_flags()
        return (None if grad_x is None else grad_x.view(ctx.x_shape), grad_w, grad_b, s_a[0:1], s_a[1:2],
                s_a[2:3] if ctx.nl else None, s_w[0:1], s_w[1:2], s_w[2:3] if ctx.nl else None, None, None) + \
            ((None, g if ctx.has_res else None) + (None,) * (ctx.n_extra - 2) if ctx.n_extra >= 2 else (None,) * ctx.n_extra)


class _LinearF32Function(torch.autograd.Function):
    """F.linear on fp32 operands as fp32-equivalent tensor-core GEMMs (ops.matmul_f32_tc), forward and backward."""

    @staticmethod
    def forward(ctx, x, w, bias):
        x2 = x.reshape(-1, w.shape[1]).contiguous()
        ctx.save_for_backward(x2, w)
        ctx.x_shape, ctx.has_bias = x.shape, bias is not None
        return ops.matmul_f32_tc(x2, w.detach().contiguous(), None if bias is None else bias.detach()).reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, g):
        x2, w = ctx.saved_tensors
        g2 = g.reshape(-1, w.shape[0]).contiguous()
        gx = ops.matmul_f32_tc(g2, w.detach().contiguous(), b_transposed=True).view(ctx.x_shape) if ctx.needs_input_grad[0] else None
        gw = ops.matmul_f32_tc(g2, x2, a_transposed=True, b_transposed=True) if ctx.needs_input_grad[1] else None
        return gx, gw, (g2.sum(0) if ctx.has_bias else None)


class WideQuantLinearFunction(torch.autograd.Function):
    """QuantizeLinear (QL:495-499) for quantizers whose codes do NOT fit the int8 pipe - more than 8 bits (GETA starts at 16 or
    32 bits and walks down, train.py:247-250 / geta.py:895-900) or weight-only mode - still on the tensor cores.

    forward : fake-quant values from the fused quantizer kernels, then F.linear as an fp32-equivalent tcgen05 GEMM (both fp32
              operands split exactly into three bf16 planes, six plane products, fp32 accumulation - ops.matmul_f32_tc).
    backward: the two gradient GEMMs the same way, then ONE fused kernel per quantizer (STE mask + step-size / range /
              exponent reductions), exactly as on the int8 path.  a_q is None in weight-only mode (QL:497)."""

    @staticmethod
    def forward(ctx, x, weight, bias, d_a, qm_a, t_a, d_w, qm_w, t_w, clip_a, clip_w):
        K, N = weight.shape[1], weight.shape[0]
        x2 = x.reshape(-1, K).contiguous()
        wv = ops.fake_quantize_sym(weight.detach(), d_w, qm_w, t_w)
        xv = x2 if d_a is None else ops.fake_quantize_sym(x2, d_a, qm_a, t_a)
        y = ops.matmul_f32_tc(xv, wv, None if bias is None else bias.detach())
        ctx.clip_a, ctx.clip_w, ctx.has_bias = clip_a, clip_w, bias is not None
        ctx.x_shape, ctx.has_a, ctx.nl_a, ctx.nl_w = x.shape, d_a is not None, t_a is not None, t_w is not None
        saved = [x2, weight, xv, wv, d_w, qm_w] + ([t_w] if t_w is not None else []) + \
                ([d_a, qm_a] + ([t_a] if t_a is not None else []) if d_a is not None else [])
        ctx.save_for_backward(*saved)
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, g):
        sv = list(ctx.saved_tensors)
        x2, weight, xv, wv, d_w, qm_w = sv[:6]
        sv = sv[6:]
        t_w = sv.pop(0) if ctx.nl_w else None
        d_a = qm_a = t_a = None
        if ctx.has_a:
            d_a, qm_a = sv.pop(0), sv.pop(0)
            t_a = sv.pop(0) if ctx.nl_a else None
        N = weight.shape[0]
        g2 = g.reshape(-1, N).contiguous()
        flags = _flags_for(g.device)
        need_x = ctx.needs_input_grad[0] or ctx.has_a
        grad_xq = ops.matmul_f32_tc(g2, wv, b_transposed=True) if need_x else None          # g @ w_q        [M, K]
        grad_wq = ops.matmul_f32_tc(g2, xv, a_transposed=True, b_transposed=True)            # g^T @ x_q      [N, K]
        if not ctx.has_a:
            grad_x, s_a = grad_xq, None
        else:
            grad_x, s_a = ops.sym_backward(x2, grad_xq, d_a, qm_a, t_a, ctx.clip_a, want_grad_x=ctx.needs_input_grad[0], flags=flags)
        grad_w, s_w = ops.sym_backward(weight.detach(), grad_wq, d_w, qm_w, t_w, ctx.clip_w, flags=flags)
        grad_b = g2.sum(0) if ctx.has_bias else None
        if EAGER_NAN_CHECK:
            check_nan_flags()
        gx = grad_x.view(ctx.x_shape) if (grad_x is not None and ctx.needs_input_grad[0]) else None
        return (gx, grad_w, grad_b,
                None if s_a is None else s_a[0:1], None if s_a is None else s_a[1:2], s_a[2:3] if (s_a is not None and ctx.nl_a) else None,
                s_w[0:1], s_w[1:2], s_w[2:3] if ctx.nl_w else None, None, None)


def _get_quantizer(qtype: QuantizationType):
    """QL:292-300."""
    if qtype == QuantizationType.SYMMETRIC_LINEAR:
        return SymQuantizerLinear
    elif qtype == QuantizationType.SYMMETRIC_NONLINEAR:
        return SymQuantizerNonLinear
    elif qtype == QuantizationType.DGE:
        return DGEQuantizer
    else:
        raise NotImplementedError


class _QuantCache:
    """Integer weight codes + saturation levels, valid for one (parameter identity, version) tuple.

    OTO pruning replaces ``module.weight`` with a new sliced Parameter (operator.py:481-499) and optimizers may
    write through ``.data`` without bumping ``_version``; the key therefore includes data_ptr/shape, and
    ``train()`` / ``load_state_dict`` / ``invalidate_quant_cache()`` drop the cache."""
    __slots__ = ("key", "w_codes", "w_sat", "a_sat", "w_q", "w_tc")

    def __init__(self):
        self.key = None
        self.w_codes = None
        self.w_sat = None
        self.a_sat = None
        self.w_q = None
        self.w_tc = None


def _bit_width(d: float, qmax: float, t: float) -> int:
    return round(math.log2(math.exp(t * math.log(qmax)) / abs(d) + 1) + 1)      # QL:394


class QuantizeMixin:
    """QL:303-410: owns d_quant_*/q_m_*/t_quant_* (each an nn.Parameter of shape (1,))."""

    def init_quantization(self, d_quant_init: float = 1.0, t_quant_init: float = 1.0, q_m_init: float = 1.0,
                          quant_type: QuantizationType = QuantizationType.SYMMETRIC_LINEAR,
                          quant_mode: QuantizationMode = QuantizationMode.WEIGHT_ONLY,
                          weight_clip_val: Tuple[float, float] = (-2.0, 2.0),
                          act_clip_val: Tuple[float, float] = (-2.0, 2.0)):
        self.d_quant_wt = nn.Parameter(torch.tensor([d_quant_init]))
        self.q_m_wt = nn.Parameter(torch.tensor([q_m_init]))
        if quant_type == QuantizationType.SYMMETRIC_NONLINEAR:
            self.t_quant_wt = nn.Parameter(torch.tensor([t_quant_init]))
        if quant_mode == QuantizationMode.WEIGHT_AND_ACTIVATION:
            self.d_quant_act = nn.Parameter(torch.tensor([d_quant_init]))
            self.q_m_act = nn.Parameter(torch.tensor([q_m_init]))
            if quant_type == QuantizationType.SYMMETRIC_NONLINEAR:
                self.t_quant_act = nn.Parameter(torch.tensor([t_quant_init]))
        self.quant_type = quant_type
        self.quant_mode = quant_mode
        self.weight_clip_val = weight_clip_val
        self.act_clip_val = act_clip_val
        self.__dict__["_qcache"] = _QuantCache()      # plain attribute: never in state_dict, not a buffer

    # ---- reference API -------------------------------------------------------------------------
    def quantize_weight(self, weight: torch.Tensor) -> torch.Tensor:
        """QL:332-354: fake-quantized weight (autograd-aware)."""
        quantizer = _get_quantizer(self.quant_type)
        if self.quant_type == QuantizationType.SYMMETRIC_LINEAR:
            return quantizer.apply(weight, self.d_quant_wt, self.q_m_wt, self.weight_clip_val, 0.0)
        # the reference passes t_quant_wt for every non-linear type (a DGE module has none -> AttributeError, QL:346-354)
        return quantizer.apply(weight, self.d_quant_wt, self.q_m_wt, self.t_quant_wt, self.weight_clip_val, 0.0)

    def quantize_act(self, activation: torch.Tensor) -> torch.Tensor:
        """QL:356-381."""
        if self.quant_mode != QuantizationMode.WEIGHT_AND_ACTIVATION:
            return activation
        quantizer = _get_quantizer(self.quant_type)
        if self.quant_type == QuantizationType.SYMMETRIC_LINEAR:
            return quantizer.apply(activation, self.d_quant_act, self.q_m_act, self.act_clip_val, 0.0)
        return quantizer.apply(activation, self.d_quant_act, self.q_m_act, self.t_quant_act, self.act_clip_val, 0.0)

    @property
    def weight_bit(self) -> int:
        """QL:383-394."""
        d = self.d_quant_wt.item()
        qmax = abs(self.q_m_wt.item())
        if self.quant_type == QuantizationType.SYMMETRIC_LINEAR:
            t = 1.0
        elif self.quant_type == QuantizationType.SYMMETRIC_NONLINEAR:
            t = self.t_quant_wt.item()
        else:
            raise NotImplementedError
        return _bit_width(d, qmax, t)

    @property
    def activation_bit(self) -> int:
        """QL:396-410."""
        if self.quant_mode != QuantizationMode.WEIGHT_AND_ACTIVATION:
            return 32
        d = self.d_quant_act.item()
        qmax = abs(self.q_m_act.item())
        if self.quant_type == QuantizationType.SYMMETRIC_LINEAR:
            t = 1.0
        elif self.quant_type == QuantizationType.SYMMETRIC_NONLINEAR:
            t = self.t_quant_act.item()
        else:
            raise NotImplementedError
        return _bit_width(d, qmax, t)

    # ---- integer-path plumbing -----------------------------------------------------------------
    def invalidate_quant_cache(self) -> None:
        self.__dict__["_qcache"] = _QuantCache()

    def _wt_qparams(self):
        return self.d_quant_wt, self.q_m_wt, getattr(self, "t_quant_wt", None)

    def _act_qparams(self):
        return self.d_quant_act, self.q_m_act, getattr(self, "t_quant_act", None)

    def _cache_key(self):
        ps = [self.weight, *[p for p in self._wt_qparams() if p is not None]]
        if self.quant_mode == QuantizationMode.WEIGHT_AND_ACTIVATION:
            ps += [p for p in self._act_qparams() if p is not None]
        return tuple((p.data_ptr(), p._version, tuple(p.shape), str(p.device)) for p in ps)

    @staticmethod
    def _sat_level(d, q_m, t) -> float:
        """round(r/d): largest code magnitude (QL:159 / QL:67), evaluated in fp32 like the kernels."""
        with torch.no_grad():
            r = torch.abs(q_m.detach().float())
            if t is not None:
                r = torch.exp(t.detach().float() * torch.log(r + 1e-6))
            v = torch.abs(torch.round(r / torch.abs(d.detach().float()))).item()      # one host read per parameter change
        return v if math.isfinite(v) else float("inf")

    def _refresh_cache(self) -> _QuantCache:
        c = self.__dict__.get("_qcache")
        if c is None:
            c = self.__dict__["_qcache"] = _QuantCache()
        key = self._cache_key()
        if c.key == key:
            return c
        c.key, c.w_codes, c.w_q, c.w_tc = key, None, None, None
        d, q, t = self._wt_qparams()
        c.w_sat = self._sat_level(d, q, t)
        c.a_sat = None
        if self.quant_mode == QuantizationMode.WEIGHT_AND_ACTIVATION:
            c.a_sat = self._sat_level(*self._act_qparams())
        return c

    def _int8_ok(self, c: _QuantCache) -> bool:
        return (self.quant_mode == QuantizationMode.WEIGHT_AND_ACTIVATION and self.quant_type != QuantizationType.DGE
                and c.w_sat <= 127 and c.a_sat is not None and c.a_sat <= 127)

    def _int8_train_ok(self) -> bool:
        """Training dispatch: is the exact int8 forward applicable?  Decided from the dispatch monitor's one-step-late,
        sync-free view of the saturation codes (see _DispatchMonitor): a model converted at 16 / 32 bits (train.py:247-250)
        moves onto the int8 tensor-core path when GETA has walked its bit width down, and back to the wide path should a
        code approach 127 again."""
        if self.quant_mode != QuantizationMode.WEIGHT_AND_ACTIVATION or self.quant_type == QuantizationType.DGE:
            return False
        ok = _monitor_for(self.weight.device).int8_ok(self)
        return ok and not self.__dict__.get("_force_wide", False)      # (_force_wide: tests compare the two paths)

    def _wide_autograd(self, x2d_or_nd: torch.Tensor, weight2d: torch.Tensor, bias) -> torch.Tensor:
        """Training forward on the wide path (DGE keeps the reference op chain: its backward is its own formula)."""
        if self.quant_type == QuantizationType.DGE or x2d_or_nd.dtype != torch.float32:
            w = self.quantize_weight(weight2d)
            x = self.quantize_act(x2d_or_nd)
            return _LinearF32Function.apply(x, w, bias) if x.dtype == torch.float32 else F.linear(x, w, bias)
        d_w, q_w, t_w = self._wt_qparams()
        if self.quant_mode == QuantizationMode.WEIGHT_AND_ACTIVATION:
            d_a, q_a, t_a = self._act_qparams()
        else:
            d_a = q_a = t_a = None
        return WideQuantLinearFunction.apply(x2d_or_nd, weight2d, bias, d_a, q_a, t_a, d_w, q_w, t_w,
                                             _clip_pair(self.act_clip_val), _clip_pair(self.weight_clip_val))

    def _weight_codes(self, c: _QuantCache) -> torch.Tensor:
        """[out, pad16(K)] int8 codes, K = in_features or C*kh*kw ordered (c, kh, kw) = weight.reshape(O, -1)."""
        if c.w_codes is None:
            w2 = self.weight.detach().reshape(self.weight.shape[0], -1)
            d, q, t = self._wt_qparams()
            c.w_codes = ops.quantize_sym(w2, d, q, t, ld_codes=ops.pad16(w2.shape[1]))
        return c.w_codes

    def _weight_fake(self, c: _QuantCache) -> torch.Tensor:
        if c.w_q is None:
            d, q, t = self._wt_qparams()
            c.w_q = ops.fake_quantize_sym(self.weight.detach(), d, q, t)
        return c.w_q

    def _needs_autograd(self, input_: torch.Tensor) -> bool:
        if not torch.is_grad_enabled():
            return False
        return input_.requires_grad or any(p.requires_grad for p in self.parameters(recurse=False))

    def train(self, mode: bool = True):
        self.invalidate_quant_cache()
        return super().train(mode)

    def _load_from_state_dict(self, *args, **kwargs):
        self.invalidate_quant_cache()
        return super()._load_from_state_dict(*args, **kwargs)

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_mon_slot"] = None
        st["_qcache"] = None          # whole-model pickles (pruning_compression.py:34) never carry device caches
        return st


def initialize_quant_layer(layer, num_bits: int = 16, quant_type: QuantizationType = QuantizationType.SYMMETRIC_LINEAR,
                           quant_mode: QuantizationMode = QuantizationMode.WEIGHT_ONLY) -> None:
    """QL:413-440: q_m = max|W|, d = q_m / (2^(b-1)-1); the activation quantizer gets the SAME values."""
    if not isinstance(layer, (QuantizeConv2d, QuantizeLinear)):
        return
    num_bits = float(num_bits)
    with torch.no_grad():
        w = layer.weight.detach()
        if w.is_cuda:
            qm = ops.absmax(w)[0]                  # warp-shuffle reduction on the device, no host sync
        else:
            qm = torch.max(torch.abs(w))           # module construction on the host happens before .to(device)
        d = (qm - 0.0) / (2 ** (num_bits - 1) - 1)
        layer.d_quant_wt.fill_(d)
        layer.q_m_wt.fill_(qm)
        if quant_type == QuantizationType.SYMMETRIC_NONLINEAR:
            layer.t_quant_wt.fill_(1.0)
        if quant_mode == QuantizationMode.WEIGHT_AND_ACTIVATION:
            layer.d_quant_act.fill_(d)
            layer.q_m_act.fill_(qm)
            if quant_type == QuantizationType.SYMMETRIC_NONLINEAR:
                layer.t_quant_act.fill_(1.0)
    layer.invalidate_quant_cache()


class QuantizeLinear(QuantizeMixin, nn.Linear):
    """QL:443-499.  (MRO note: the mixin comes first so its train()/_load_from_state_dict hooks run; the
    reference lists nn.Linear first, which only matters for attribute lookup of names both define - none.)"""

    def __init__(self, in_features, out_features, bias=True, d_quant_init=1.0, t_quant_init=1.0, q_m_init=1.0,
                 quant_type=QuantizationType.SYMMETRIC_LINEAR, quant_mode=QuantizationMode.WEIGHT_ONLY):
        nn.Linear.__init__(self, in_features, out_features, bias)
        self.init_quantization(d_quant_init, t_quant_init, q_m_init, quant_type, quant_mode)

    @staticmethod
    def from_module(module=None, d_quant_init=1.0, t_quant_init=1.0, q_m_init=1.0,
                    quant_type=QuantizationType.SYMMETRIC_LINEAR, quant_mode=QuantizationMode.WEIGHT_ONLY,
                    quant_init_by_module=True, num_bits=8):
        """QL:460-493."""
        q = QuantizeLinear(in_features=module.in_features, out_features=module.out_features,
                           bias=module.bias is not None, d_quant_init=d_quant_init, t_quant_init=t_quant_init,
                           q_m_init=q_m_init, quant_type=quant_type, quant_mode=quant_mode)
        q.to(module.weight.device)
        q.weight.data.copy_(module.weight.data)
        if module.bias is not None:
            q.bias.data.copy_(module.bias.data)
        if quant_init_by_module:
            initialize_quant_layer(q, num_bits=num_bits, quant_type=quant_type, quant_mode=quant_mode)
        return q

    fuses_pre_act = True     # forward(x, pre_act="gelu"): the caller's nn.GELU in front of this layer is applied here

    def forward(self, input_: torch.Tensor, pre_act: Optional[str] = None, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        """QL:495-499.  Extensions used by the drop-in ViT: ``pre_act="gelu"`` - ``input_`` is the pre-activation of the exact-erf
        GELU in front of this layer (fused into the quantizer kernels on the int8 QAT path, simply applied first elsewhere);
        ``residual`` - returns ``residual + layer(input_)`` (in the GEMM epilogue on the int8 QAT path)."""
        ops._lib.require_cuda(input_, self.weight)
        if pre_act not in (None, "gelu"):
            raise ValueError(f"QuantizeLinear: unsupported pre_act {pre_act!r}")
        if self._needs_autograd(input_) and input_.dtype == torch.float32 and self._int8_train_ok():
            d_a, q_a, t_a = self._act_qparams()
            d_w, q_w, t_w = self._wt_qparams()
            fuse = pre_act == "gelu" and self.in_features % 16 == 0
            if pre_act and not fuse:
                input_ = F.gelu(input_)
            res_ok = residual is not None and residual.dtype == torch.float32 and residual.is_cuda and \
                tuple(residual.shape) == tuple(input_.shape[:-1]) + (self.out_features,)
            y = QuantLinearFunction.apply(input_, self.weight, self.bias, d_a, q_a, t_a, d_w, q_w, t_w,
                                          _clip_pair(self.act_clip_val), _clip_pair(self.weight_clip_val), fuse,
                                          residual if res_ok else None)
            return y if (residual is None or res_ok) else residual + y
        if pre_act:
            input_ = F.gelu(input_)
        y = self._forward_plain(input_)
        return y if residual is None else residual + y

    def _forward_plain(self, input_: torch.Tensor) -> torch.Tensor:
        if self._needs_autograd(input_):
            return self._wide_autograd(input_, self.weight, self.bias)
        c = self._refresh_cache()
        if self._int8_ok(c) and input_.dtype == torch.float32:
            K, N = self.in_features, self.out_features
            flags = _flags_for(input_.device)
            d_a, q_a, t_a = self._act_qparams()
            a_codes = ops.quantize_sym(input_, d_a, q_a, t_a, ld_codes=ops.pad16(K), flags=flags)
            y = ops.gemm_i8(a_codes, self._weight_codes(c), K, N, out_kind=ops.QVIT_OUT_F32, scale_a=d_a,
                            scale_w=self.d_quant_wt, bias=self.bias, flags=flags)
            return y.reshape(*input_.shape[:-1], N)
        # wide path: codes do not fit int8 (or weight-only mode) -> fused quantizer kernels + fp32-equivalent tensor-core GEMM
        x = input_.reshape(-1, self.in_features)
        if self.quant_mode == QuantizationMode.WEIGHT_AND_ACTIVATION:
            x = ops.fake_quantize_sym(x, *self._act_qparams())
        y = ops.matmul_f32_tc(x.float(), self._weight_fake(c), self.bias)
        return y.reshape(*input_.shape[:-1], self.out_features).to(input_.dtype)


class QuantizeConv2d(QuantizeMixin, nn.Conv2d):
    """QL:502-587 (defaults padding=1, bias=False as upstream)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=1, dilation=1, groups=1, bias=False,
                 d_quant_init=1.0, t_quant_init=1.0, q_m_init=1.0, quant_type=QuantizationType.SYMMETRIC_LINEAR,
                 quant_mode=QuantizationMode.WEIGHT_ONLY):
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias=bias)
        self.init_quantization(d_quant_init, t_quant_init, q_m_init, quant_type, quant_mode)

    @staticmethod
    def from_module(module=None, d_quant_init=1.0, t_quant_init=1.0, q_m_init=1.0,
                    quant_type=QuantizationType.SYMMETRIC_LINEAR, quant_mode=QuantizationMode.WEIGHT_ONLY,
                    quant_init_by_module=True, num_bits=8):
        """QL:534-573."""
        q = QuantizeConv2d(in_channels=module.in_channels, out_channels=module.out_channels,
                           kernel_size=module.kernel_size, stride=module.stride, padding=module.padding,
                           dilation=module.dilation, groups=module.groups, bias=module.bias is not None,
                           d_quant_init=d_quant_init, t_quant_init=t_quant_init, q_m_init=q_m_init,
                           quant_type=quant_type, quant_mode=quant_mode)
        q.to(module.weight.device)
        q.weight.data.copy_(module.weight.data)
        if module.bias is not None:
            q.bias.data.copy_(module.bias.data)
        if quant_init_by_module:
            initialize_quant_layer(q, num_bits=num_bits, quant_type=quant_type, quant_mode=quant_mode)
        return q

    def _patch_matrix(self, x: torch.Tensor):
        """im2col of an NCHW tensor -> ([B * OH * OW, C * kh * kw] fp32 with K ordered (c, kh, kw), OH * OW, OH), autograd-aware.
        Non-overlapping patches (the ViT patch embedding, vit_model.py:94-103) are a pure permutation: one strided copy for the
        whole batch; everything else goes through F.unfold (one im2col launch per image)."""
        B, Cc, Hh, Ww = x.shape
        kh, kw = self.kernel_size
        if self.stride == self.kernel_size and self.padding == (0, 0) and self.dilation == (1, 1) and Hh % kh == 0 and Ww % kw == 0:
            OH, OW = Hh // kh, Ww // kw
            cols2 = x.reshape(B, Cc, OH, kh, OW, kw).permute(0, 2, 4, 1, 3, 5).reshape(B * OH * OW, Cc * kh * kw)
            return cols2, OH * OW, OH
        cols = F.unfold(x, self.kernel_size, self.dilation, self.padding, self.stride)                 # [B, C*kh*kw, L]
        L = cols.shape[-1]
        OH = (Hh + 2 * self.padding[0] - self.dilation[0] * (kh - 1) - 1) // self.stride[0] + 1
        return cols.transpose(1, 2).reshape(B * L, -1), L, OH

    def forward(self, input_: torch.Tensor) -> torch.Tensor:
        """QL:575-587."""
        ops._lib.require_cuda(input_, self.weight)
        plain = (self.groups == 1 and input_.dim() == 4 and not isinstance(self.padding, str) and self.padding_mode == "zeros"
                 and input_.dtype == torch.float32)
        if self._needs_autograd(input_):
            if not plain:                      # grouped / exotic convolutions: the reference op chain on library kernels
                weight = self.quantize_weight(self.weight)
                if self.quant_mode == QuantizationMode.WEIGHT_AND_ACTIVATION:
                    input_ = self.quantize_act(input_)
                return F.conv2d(input_, weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
            # conv = im2col + QuantizeLinear on the patch matrix, through the same autograd Functions as QuantizeLinear: the
            # quantizers are elementwise, so quantizing the unfolded matrix gives the same values, and the step-size / range
            # sums over the duplicated elements equal the reference's sums over x against the folded gradient (padding zeros
            # contribute nothing).  F.unfold / its backward (col2im) are data movement, not arithmetic.
            B, O = input_.shape[0], self.out_channels
            cols2, L, OH = self._patch_matrix(input_)
            w2 = self.weight.reshape(O, -1)
            if self._int8_train_ok():
                d_a, q_a, t_a = self._act_qparams()
                d_w, q_w, t_w = self._wt_qparams()
                y2 = QuantLinearFunction.apply(cols2, w2, self.bias, d_a, q_a, t_a, d_w, q_w, t_w,
                                               _clip_pair(self.act_clip_val), _clip_pair(self.weight_clip_val))
            else:
                y2 = self._wide_autograd(cols2, w2, self.bias)
            return y2.view(B, L, O).permute(0, 2, 1).reshape(B, O, OH, L // OH)
        c = self._refresh_cache()
        if (self._int8_ok(c) and plain and self.padding[0] == self.padding[1]
                and ops.ultra_conv_tc_supported(self.in_channels, self.out_channels, *self.kernel_size)):
            # implicit GEMM on the tensor cores: NHWC int8 activation codes, packed weight codes resident in shared memory, the
            # im2col tile gathered on chip - nothing of the kh*kw-fold im2col matrix ever exists in HBM
            flags = _flags_for(input_.device)
            d_a, q_a, t_a = self._act_qparams()
            a = ops.quantize_sym(input_.permute(0, 2, 3, 1).contiguous(), d_a, q_a, t_a, flags=flags).view(
                input_.shape[0], input_.shape[2], input_.shape[3], self.in_channels)
            if c.w_tc is None:
                wc = self._weight_codes(c)[:, : self.in_channels * self.kernel_size[0] * self.kernel_size[1]]
                wc = wc.reshape(self.out_channels, self.in_channels, *self.kernel_size).permute(0, 2, 3, 1).contiguous()
                c.w_tc = ops.pack_conv_weights_tc(wc)
            return ops.conv2d_i8_tc(a, c.w_tc, self.out_channels, self.kernel_size, self.stride, self.padding[0], self.dilation, d_a,
                                    self.d_quant_wt, self.bias)
        if self._int8_ok(c) and plain:
            flags = _flags_for(input_.device)
            d_a, q_a, t_a = self._act_qparams()
            cols, OH, OW = ops.im2col_quantize_sym(input_, self.kernel_size, self.stride, self.padding, self.dilation,
                                                   d_a, q_a, t_a, flags=flags)
            K = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
            y = ops.gemm_i8(cols, self._weight_codes(c), K, self.out_channels, out_kind=ops.QVIT_OUT_F32, scale_a=d_a,
                            scale_w=self.d_quant_wt, bias=self.bias, flags=flags)
            # [B*OH*OW, O] is NHWC memory; hand back the NCHW-shaped view, i.e. a torch.channels_last tensor - the same
            # memory format cuDNN hands back for channels_last inputs (PatchEmbed's flatten(2).transpose(1,2),
            # vit_model.py:100, turns it into the contiguous token matrix without a copy; code that calls .view() on a conv
            # output needs .contiguous() first, exactly as with any channels_last tensor)
            return y.view(input_.shape[0], OH, OW, self.out_channels).permute(0, 3, 1, 2)
        x = input_
        if self.quant_mode == QuantizationMode.WEIGHT_AND_ACTIVATION:
            x = ops.fake_quantize_sym(input_, *self._act_qparams())
        if plain:                              # wide path: im2col + fp32-equivalent tensor-core GEMM
            B, O = x.shape[0], self.out_channels
            cols2, L, OH = self._patch_matrix(x)
            y2 = ops.matmul_f32_tc(cols2, self._weight_fake(c).reshape(O, -1), self.bias)
            return y2.view(B, L, O).permute(0, 2, 1).reshape(B, O, OH, L // OH)
        return F.conv2d(x, self._weight_fake(c), self.bias, self.stride, self.padding, self.dilation, self.groups)


LAYER_TO_QUANTLAYER = {"Linear": QuantizeLinear, "Conv2d": QuantizeConv2d}     # QL:590
