"""Tensor-level wrappers over the C ABI (include/qvit_b200.h).

PyTorch is plumbing here: it owns device memory (``torch.empty``) and the current stream; every function
below only validates arguments, allocates outputs and forwards raw device pointers to libqvit_b200.
All inputs must live on a CUDA device - there is no CPU path (the reference's CPU semantics live in
``oracle/`` and are used by tests only).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (QVIT_ACT_GELU, QVIT_ACT_NONE, QVIT_ACT_RELU, QVIT_GEMM_AUTO, QVIT_GEMM_SIMT, QVIT_GEMM_TCGEN05,
                   QVIT_OUT_BF16, QVIT_OUT_F16X2, QVIT_OUT_F32, QVIT_OUT_I8, QVIT_OUT_I32, QVIT_OUT_NONE)

__all__ = ["pad16", "quantize_sym", "fake_quantize_sym", "sym_backward", "absmax", "im2col_quantize_sym", "gemm_i8",
           "layernorm_quantize", "embed_assemble", "layernorm_fwd", "layernorm_bwd", "layernorm_supported", "attention_train_supported", "attention_train_fwd", "attention_train_bwd", "attention_f32", "attention_f32_supported", "split3_bf16", "grad_prep", "codes_to_bf16", "gemm_bf16_split_t", "codes_to_bf16_t", "gemm_bf16_split", "matmul_f32_tc", "ultra_weight_codes", "ultra_act", "uniform_quantize", "ultra_bn_act_pool_nchw", "conv2d_f32_wcodes", "ultra_conv_bn_act", "ultra_conv_tc", "conv2d_i8_tc", "pack_conv_weights_tc", "ultra_conv_tc_supported", "bn_fold",
           "bn_act_quantize_int", "pack_int4", "unpack_int4", "new_flags", "QVIT_OUT_I32", "QVIT_OUT_F32",
           "QVIT_OUT_BF16", "QVIT_OUT_I8", "QVIT_OUT_NONE", "QVIT_OUT_F16X2", "attention_f16x2", "split2_f16", "f16x2_exponent", "QVIT_ACT_NONE", "QVIT_ACT_GELU", "QVIT_ACT_RELU", "QVIT_GEMM_AUTO",
           "QVIT_GEMM_TCGEN05", "QVIT_GEMM_SIMT"]

_OUT_DTYPE = {QVIT_OUT_I32: torch.int32, QVIT_OUT_F32: torch.float32, QVIT_OUT_BF16: torch.bfloat16,
              QVIT_OUT_I8: torch.int8, QVIT_OUT_F16X2: torch.float16}


def pad16(k: int) -> int:
    """Row pitch (bytes) of an int8 code matrix: TMA needs global strides that are multiples of 16 bytes."""
    return (int(k) + 15) // 16 * 16


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    _lib.require_cuda(t)
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _scalar_param(t, dev, what: str) -> torch.Tensor:
    """(1,) fp32 device tensor for a quantizer parameter (nn.Parameter, tensor or Python number)."""
    if isinstance(t, torch.Tensor):
        if t.device != dev or t.dtype != torch.float32:
            t = t.detach().to(device=dev, dtype=torch.float32)     # the reference does .to(device) too (QL:148-151)
        if t.numel() != 1:
            raise ValueError(f"{what}: quantizer parameters are per-tensor scalars of shape (1,)")
        return t.detach().contiguous()
    return torch.tensor([float(t)], dtype=torch.float32, device=dev)


def new_flags(device) -> torch.Tensor:
    return torch.zeros(1, dtype=torch.int32, device=device)


# ------------------------------------------------------------------------------------------ GETA quantizers
def quantize_sym(x: torch.Tensor, d, q_m, t=None, ld_codes: Optional[int] = None, flags: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None, gelu: bool = False) -> torch.Tensor:
    """int8 codes of SymQuantizerLinear/NonLinear.forward (QL:136-161 / QL:40-69) for x viewed as [rows, cols]
    (cols = last dim).  Returns [rows, ld_codes] int8, padding columns zero.  gelu=True: the codes of gelu(x) (nn.GELU of
    Mlp.forward fused into the quantizer; fp32, cols a multiple of 16, ld_codes == cols)."""
    if gelu:
        x = _f32c(x, "quantize_sym")
        x2 = x.reshape(-1, x.shape[-1])
        rows, cols = x2.shape
        if (ld_codes is not None and int(ld_codes) != cols) or out is not None or cols % 4:
            raise ValueError("quantize_sym(gelu=True): contiguous codes only (ld_codes == cols, cols % 4 == 0)")
        dev = x2.device
        d_, q_ = _scalar_param(d, dev, "d_quant"), _scalar_param(q_m, dev, "q_m")
        t_ = None if t is None else _scalar_param(t, dev, "t_quant")
        out = torch.empty((rows, cols), dtype=torch.int8, device=dev)
        _lib.check(_lib.lib().qvit_gelu_quantize_sym(_lib.ptr(x2), x2.numel(), _lib.ptr(d_), _lib.ptr(q_), _lib.ptr(t_), _lib.ptr(out),
                                                     _lib.ptr(flags), _lib.stream()), "qvit_gelu_quantize_sym")
        return out
    if x.dtype == torch.bfloat16:
        _lib.require_cuda(x)
        x2 = x.reshape(-1, x.shape[-1]) if x.dim() > 1 else x.reshape(1, -1)
        x2 = x2 if x2.stride(-1) == 1 else x2.contiguous()
        fn = _lib.lib().qvit_quantize_sym_bf16
    else:
        x = _f32c(x, "quantize_sym")
        x2 = x.reshape(-1, x.shape[-1]) if x.dim() > 1 else x.reshape(1, -1)
        fn = _lib.lib().qvit_quantize_sym
    rows, cols = x2.shape
    ld = cols if ld_codes is None else int(ld_codes)
    if ld < cols:
        raise ValueError("ld_codes < cols")
    dev = x2.device
    d_, q_ = _scalar_param(d, dev, "d_quant"), _scalar_param(q_m, dev, "q_m")
    t_ = None if t is None else _scalar_param(t, dev, "t_quant")
    if out is None:
        out = torch.empty((rows, ld), dtype=torch.int8, device=dev)
    elif out.shape != (rows, ld) or out.dtype != torch.int8 or not out.is_contiguous():
        raise ValueError("quantize_sym: bad `out`")
    _lib.check(fn(_lib.ptr(x2), rows, cols, x2.stride(0) if rows > 1 else max(cols, x2.stride(0)), _lib.ptr(d_),
                  _lib.ptr(q_), _lib.ptr(t_), _lib.ptr(out), ld, _lib.ptr(flags), _lib.stream()), "qvit_quantize_sym")
    return out


def fake_quantize_sym(x: torch.Tensor, d, q_m, t=None) -> torch.Tensor:
    """fp32 fake-quant values, bit-for-bit the reference Function output (sign(x) * d * round(p/d))."""
    x = _f32c(x, "fake_quantize_sym")
    dev = x.device
    d_, q_ = _scalar_param(d, dev, "d_quant"), _scalar_param(q_m, dev, "q_m")
    t_ = None if t is None else _scalar_param(t, dev, "t_quant")
    out = torch.empty_like(x)
    _lib.check(_lib.lib().qvit_fake_quantize_sym(_lib.ptr(x), x.numel(), _lib.ptr(d_), _lib.ptr(q_), _lib.ptr(t_),
                                                 _lib.ptr(out), _lib.stream()), "qvit_fake_quantize_sym")
    return out


def sym_backward(x: torch.Tensor, g: torch.Tensor, d, q_m, t=None, clip: Tuple[float, float] = (-2.0, 2.0),
                 want_grad_x: bool = True, flags: Optional[torch.Tensor] = None, gelu: bool = False):
    """One fused pass: (grad_x | None, scalars[3] = grad_d, grad_qm, grad_t).  QL:163-205 / QL:71-125.
    gelu=True: x is the pre-activation of a GELU in front of the quantizer; grad_x is the gradient w.r.t. that pre-activation."""
    x = _f32c(x, "sym_backward x")
    g = _f32c(g, "sym_backward g")
    if g.shape != x.shape:
        raise ValueError("sym_backward: grad_output shape mismatch")
    dev = x.device
    d_, q_ = _scalar_param(d, dev, "d_quant"), _scalar_param(q_m, dev, "q_m")
    t_ = None if t is None else _scalar_param(t, dev, "t_quant")
    grad_x = torch.empty_like(x) if want_grad_x else None
    scalars = torch.zeros(3, dtype=torch.float32, device=dev)
    fn = _lib.lib().qvit_gelu_sym_backward if gelu else _lib.lib().qvit_sym_backward
    _lib.check(fn(_lib.ptr(x), _lib.ptr(g), x.numel(), _lib.ptr(d_), _lib.ptr(q_), _lib.ptr(t_), float(clip[0]), float(clip[1]),
                  _lib.ptr(grad_x), _lib.ptr(scalars), _lib.ptr(flags), _lib.stream()), "qvit_sym_backward")
    return grad_x, scalars


def absmax(x: torch.Tensor) -> torch.Tensor:
    """max|x| as a (1,) device tensor (initialize_quant_layer, QL:423) - no host sync."""
    x = _f32c(x, "absmax")
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().qvit_absmax(_lib.ptr(x), x.numel(), _lib.ptr(out), _lib.stream()), "qvit_absmax")
    return out


def im2col_quantize_sym(x: torch.Tensor, kernel, stride, padding, dilation, d, q_m, t=None,
                        flags: Optional[torch.Tensor] = None):
    """Activation quantizer fused with im2col: NCHW fp32 -> ([B*OH*OW, pad16(C*kh*kw)] int8, OH, OW)."""
    x = _f32c(x, "im2col_quantize_sym")
    B, Cc, H, W = x.shape
    kh, kw = kernel
    sh, sw = stride
    ph, pw = padding
    dh, dw = dilation
    OH = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    OW = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    if OH <= 0 or OW <= 0:
        raise ValueError("im2col_quantize_sym: empty output")
    K = Cc * kh * kw
    ld = pad16(K)
    dev = x.device
    d_, q_ = _scalar_param(d, dev, "d_quant"), _scalar_param(q_m, dev, "q_m")
    t_ = None if t is None else _scalar_param(t, dev, "t_quant")
    cols = torch.empty((B * OH * OW, ld), dtype=torch.int8, device=dev)
    _lib.check(_lib.lib().qvit_im2col_quantize_sym(_lib.ptr(x), B, Cc, H, W, kh, kw, sh, sw, ph, pw, dh, dw, _lib.ptr(d_),
                                                   _lib.ptr(q_), _lib.ptr(t_), _lib.ptr(cols), ld, _lib.ptr(flags),
                                                   _lib.stream()), "qvit_im2col_quantize_sym")
    return cols, OH, OW


def layernorm_quantize(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, d, q_m, t=None,
                       ld_codes: Optional[int] = None, want_ln: bool = False, flags: Optional[torch.Tensor] = None):
    """codes = Q(LayerNorm(x)) for x [rows, cols] (vit_model.py:206-207 feeding QL:356-381).  Returns (codes, ln|None)."""
    x = _f32c(x, "layernorm_quantize")
    x2 = x.reshape(-1, x.shape[-1])
    rows, cols = x2.shape
    ld = pad16(cols) if ld_codes is None else int(ld_codes)
    dev = x.device
    gamma, beta = _f32c(gamma, "gamma"), _f32c(beta, "beta")
    d_, q_ = _scalar_param(d, dev, "d_quant"), _scalar_param(q_m, dev, "q_m")
    t_ = None if t is None else _scalar_param(t, dev, "t_quant")
    codes = torch.empty((rows, ld), dtype=torch.int8, device=dev)
    ln = torch.empty_like(x2) if want_ln else None
    _lib.check(_lib.lib().qvit_layernorm_quantize(_lib.ptr(x2), rows, cols, _lib.ptr(gamma), _lib.ptr(beta), float(eps),
                                                  _lib.ptr(d_), _lib.ptr(q_), _lib.ptr(t_), _lib.ptr(codes), ld,
                                                  _lib.ptr(ln), _lib.ptr(flags), _lib.stream()), "qvit_layernorm_quantize")
    return codes, ln


def embed_assemble(tok: torch.Tensor, pos: torch.Tensor, cls: torch.Tensor, B: int) -> torch.Tensor:
    """cat(cls_token, x) + pos_embed (vit_model.py:295-305) in one pass: tok [B * P, D] fp32 -> h [B, P + 1, D] fp32."""
    tok, pos, cls = _f32c(tok, "embed_assemble tok"), _f32c(pos, "pos_embed"), _f32c(cls, "cls_token")
    D = tok.shape[-1]
    P = tok.shape[0] // B
    if pos.numel() != (P + 1) * D or cls.numel() != D or tok.shape[0] != B * P:
        raise ValueError("embed_assemble: shapes do not match (tok [B*P, D], pos [P+1, D], cls [D])")
    h = torch.empty((B, P + 1, D), dtype=torch.float32, device=tok.device)
    _lib.check(_lib.lib().qvit_embed_assemble(_lib.ptr(tok), _lib.ptr(pos), _lib.ptr(cls), B, P, D, _lib.ptr(h), _lib.stream()),
               "qvit_embed_assemble")
    return h


def layernorm_supported(cols: int) -> bool:
    return cols % 128 == 0 and 0 < cols <= 1024


def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float):
    """(y, mean, rstd) of LayerNorm over the last dim (training caller; the inference engine uses layernorm_quantize)."""
    x = _f32c(x, "layernorm_fwd")
    x2 = x.reshape(-1, x.shape[-1])
    rows, cols = x2.shape
    gamma, beta = _f32c(gamma.detach(), "gamma"), _f32c(beta.detach(), "beta")
    y = torch.empty_like(x2)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().qvit_layernorm_fwd(_lib.ptr(x2), rows, cols, _lib.ptr(gamma), _lib.ptr(beta), float(eps), _lib.ptr(y),
                                             _lib.ptr(mean), _lib.ptr(rstd), _lib.stream()), "qvit_layernorm_fwd")
    return y.view(x.shape), mean, rstd


def layernorm_bwd(x: torch.Tensor, gy: torch.Tensor, gamma: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor,
                  add: Optional[torch.Tensor] = None):
    """(gx, dgamma, dbeta) of LayerNorm; ``add``: a gradient reaching x over the residual path, summed into gx in the same pass."""
    x, gy = _f32c(x, "layernorm_bwd x"), _f32c(gy, "layernorm_bwd gy")
    if add is not None:
        add = _f32c(add, "layernorm_bwd add")
        if add.shape != x.shape:
            raise ValueError("layernorm_bwd: add must have the shape of x")
    x2, g2 = x.reshape(-1, x.shape[-1]), gy.reshape(-1, x.shape[-1])
    rows, cols = x2.shape
    gamma = _f32c(gamma.detach(), "gamma")
    gx = torch.empty_like(x2)
    dg = torch.empty(cols, dtype=torch.float32, device=x.device)
    db = torch.empty(cols, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().qvit_layernorm_bwd(_lib.ptr(x2), _lib.ptr(g2), rows, cols, _lib.ptr(gamma), _lib.ptr(mean), _lib.ptr(rstd),
                                             _lib.ptr(add), _lib.ptr(gx), _lib.ptr(dg), _lib.ptr(db), _lib.stream()), "qvit_layernorm_bwd")
    return gx.view(x.shape), dg, db


def attention_f32(qkv: torch.Tensor, num_heads: int, scale: Optional[float] = None) -> torch.Tensor:
    """softmax(q k^T * scale) v for qkv [B, T, 3*H*64] fp32 (ViTAttention.forward, vit_model.py:133-149) -> [B, T, H*64] fp32.
    Exact 3-way bf16 split on tcgen05 kind::f16: fp32-equivalent accuracy.  Requires head_dim == 64 and T <= 208."""
    qkv = _f32c(qkv, "attention_f32")
    B, T, C3 = qkv.shape
    hd = C3 // (3 * num_heads)
    if hd * 3 * num_heads != C3:
        raise ValueError("attention_f32: last dim must be 3 * num_heads * head_dim")
    out = torch.empty((B, T, num_heads * hd), dtype=torch.float32, device=qkv.device)
    sc = float(hd) ** -0.5 if scale is None else float(scale)
    _lib.check(_lib.lib().qvit_attention_f32(_lib.ptr(qkv), B, T, num_heads, hd, sc, _lib.ptr(out), _lib.stream()),
               "qvit_attention_f32")
    return out


def attention_quantize_sym(qkv: torch.Tensor, num_heads: int, d, q_m, t=None, *, scale: Optional[float] = None,
                           want_context: bool = False, flags: Optional[torch.Tensor] = None):
    """attention_f32 with the consumer layer's quantize_act fused into the epilogue (ViTAttention.forward, vit_model.py:141-151:
    the context only feeds `proj`).  Returns (codes [B*T, pad16(H*64)] int8, context [B, T, H*64] fp32 | None)."""
    qkv = _f32c(qkv, "attention_quantize_sym")
    B, T, C3 = qkv.shape
    hd = C3 // (3 * num_heads)
    if hd * 3 * num_heads != C3:
        raise ValueError("attention_quantize_sym: last dim must be 3 * num_heads * head_dim")
    dev = qkv.device
    C = num_heads * hd
    ld = pad16(C)
    codes = torch.empty((B * T, ld), dtype=torch.int8, device=dev)
    if ld > C:
        codes[:, C:].zero_()
    ctx = torch.empty((B, T, C), dtype=torch.float32, device=dev) if want_context else None
    d_, q_ = _scalar_param(d, dev, "d_quant"), _scalar_param(q_m, dev, "q_m")
    t_ = None if t is None else _scalar_param(t, dev, "t_quant")
    sc = float(hd) ** -0.5 if scale is None else float(scale)
    _lib.check(_lib.lib().qvit_attention_quantize_sym(_lib.ptr(qkv), B, T, num_heads, hd, sc, _lib.ptr(d_), _lib.ptr(q_), _lib.ptr(t_),
                                                      _lib.ptr(codes), ld, _lib.ptr(ctx), _lib.ptr(flags), _lib.stream()),
               "qvit_attention_quantize_sym")
    return codes, ctx


def f16x2_exponent(bound: float) -> int:
    """Power of two e such that |x| <= bound implies |x * 2^e| <= 2^15 (< 65504, the largest fp16): the scale the producer of
    a two-plane fp16 tensor applies."""
    import math
    if not (bound > 0.0) or not math.isfinite(bound):
        return 0
    return 15 - int(math.ceil(math.log2(bound)))


def split2_f16(x: torch.Tensor, col_exp: torch.Tensor, flags: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [rows, cols] -> two-plane fp16 [rows, 2*cols] (hi | lo) of x * 2^col_exp[c]."""
    x = _f32c(x, "split2_f16")
    x2 = x.reshape(-1, x.shape[-1])
    rows, cols = x2.shape
    _lib.require_cuda(col_exp)
    ce = col_exp.to(torch.int32).contiguous()
    out = torch.empty((rows, 2 * cols), dtype=torch.float16, device=x.device)
    _lib.check(_lib.lib().qvit_split2_f16(_lib.ptr(x2), rows, cols, x2.stride(0), _lib.ptr(ce), _lib.ptr(out), 2 * cols, cols,
                                          _lib.ptr(flags), _lib.stream()), "qvit_split2_f16")
    return out


def attention_f16x2(planes: torch.Tensor, B: int, T: int, num_heads: int, exps, d=None, q_m=None, t=None, *,
                    plane_off: Optional[int] = None, scale: Optional[float] = None, want_codes: bool = True,
                    want_context: bool = False, flags: Optional[torch.Tensor] = None, prof: Optional[torch.Tensor] = None):
    """softmax(q k^T * scale) v from the two-plane fp16 qkv [B*T, ld] (QVIT_OUT_F16X2 of the qkv GEMM); exps = (eq, ek, ev).
    Returns (codes [B*T, pad16(H*64)] int8 | None, context [B, T, H*64] fp32 | None)."""
    _lib.require_cuda(planes)
    if planes.dtype != torch.float16 or planes.dim() != 2 or planes.stride(1) != 1:
        raise TypeError("attention_f16x2: planes must be a 2-D fp16 tensor with unit inner stride")
    ld = planes.stride(0)
    hd = 64
    C = num_heads * hd
    plane_off = ld // 2 if plane_off is None else int(plane_off)
    dev = planes.device
    codes = ctx = None
    ldc = pad16(C)
    d_ = q_ = t_ = None
    if want_codes:
        codes = torch.empty((B * T, ldc), dtype=torch.int8, device=dev)
        if ldc > C:
            codes[:, C:].zero_()
        d_, q_ = _scalar_param(d, dev, "d_quant"), _scalar_param(q_m, dev, "q_m")
        t_ = None if t is None else _scalar_param(t, dev, "t_quant")
    if want_context:
        ctx = torch.empty((B, T, C), dtype=torch.float32, device=dev)
    sc = float(hd) ** -0.5 if scale is None else float(scale)
    _lib.check(_lib.lib().qvit_attention_f16x2(_lib.ptr(planes), ld, plane_off, B, T, num_heads, hd, sc, int(exps[0]), int(exps[1]),
                                               int(exps[2]), _lib.ptr(d_), _lib.ptr(q_), _lib.ptr(t_), _lib.ptr(codes), ldc,
                                               _lib.ptr(ctx), _lib.ptr(flags), _lib.ptr(prof), _lib.stream()), "qvit_attention_f16x2")
    return codes, ctx


def attention_train_supported(T: int, head_dim: int) -> bool:
    return head_dim == 64 and 1 <= T <= 208


def attention_train_fwd(qkv: torch.Tensor, num_heads: int, scale: Optional[float] = None):
    """Training forward of the attention core on the output of the qkv layer ([B, T, 3 * H * 64] fp32, parts q | k | v, head-major
    inside a part).  Returns (out [B, T, H * 64] fp32, lse [B, H, 256] fp32 - what attention_train_bwd needs)."""
    qkv = _f32c(qkv, "attention_train_fwd")
    B, T, D3 = qkv.shape
    hd = D3 // (3 * num_heads)
    sc = float(hd) ** -0.5 if scale is None else float(scale)
    planes = torch.empty((B * T, 2 * D3), dtype=torch.float16, device=qkv.device)
    out = torch.empty((B, T, D3 // 3), dtype=torch.float32, device=qkv.device)
    lse = torch.empty((B, num_heads, 256), dtype=torch.float32, device=qkv.device)
    _lib.check(_lib.lib().qvit_attention_train_fwd(_lib.ptr(qkv), B, T, num_heads, hd, sc, _lib.ptr(planes), _lib.ptr(out), _lib.ptr(lse),
                                                   _lib.stream()), "qvit_attention_train_fwd")
    return out, lse


def attention_train_bwd(qkv: torch.Tensor, out: torch.Tensor, lse: torch.Tensor, dout: torch.Tensor, num_heads: int,
                        scale: Optional[float] = None, prof: Optional[torch.Tensor] = None) -> torch.Tensor:
    """d loss / d qkv ([B, T, 3 * H * 64] fp32) from d loss / d out."""
    qkv, out, dout = _f32c(qkv, "qkv"), _f32c(out, "out"), _f32c(dout, "dout")
    B, T, D3 = qkv.shape
    hd = D3 // (3 * num_heads)
    sc = float(hd) ** -0.5 if scale is None else float(scale)
    dstat = torch.empty((B, num_heads, 256), dtype=torch.float32, device=qkv.device)
    dqkv = torch.empty_like(qkv)
    _lib.check(_lib.lib().qvit_attention_train_bwd_prof(_lib.ptr(qkv), _lib.ptr(out), _lib.ptr(dout), _lib.ptr(lse), B, T, num_heads, hd,
                                                        sc, _lib.ptr(dstat), _lib.ptr(dqkv), _lib.ptr(prof), _lib.stream()),
               "qvit_attention_train_bwd")
    return dqkv


def attention_f32_supported(T: int, head_dim: int) -> bool:
    return head_dim == 64 and T <= 208


# ------------------------------------------------------------------------------------------ GEMM
def gemm_i8(a: torch.Tensor, w: torch.Tensor, K: int, N: Optional[int] = None, *, out_kind: int = QVIT_OUT_F32,
            act: int = QVIT_ACT_NONE, scale_a=None, scale_w=None, scale_const: float = 1.0,
            col_scale: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
            residual: Optional[torch.Tensor] = None, next_q=None, flags: Optional[torch.Tensor] = None,
            backend: int = QVIT_GEMM_AUTO, out: Optional[torch.Tensor] = None, ldo: Optional[int] = None,
            acc_abs_max: int = 0) -> torch.Tensor:
    """acc = A[M, :K] @ W[N, :K]^T (int8/uint8 x int8 -> int32) + fused epilogue (include/qvit_b200.h).

    acc_abs_max: optional promise |acc| <= sat_a * sat_w * K (0 = unknown) - lets the epilogue skip the conversion unit.

    a: [M, lda] int8 or uint8 codes, w: [N, ldw] int8 codes; both row-major with the pitch as second dim.
    next_q = (d, q_m, t|None) of the consumer layer for QVIT_OUT_I8."""
    _lib.require_cuda(a, w)
    if a.dtype not in (torch.int8, torch.uint8) or w.dtype != torch.int8:
        raise TypeError("gemm_i8: a must be int8/uint8 and w int8")
    if a.dim() != 2 or w.dim() != 2 or a.stride(1) != 1 or w.stride(1) != 1:
        raise ValueError("gemm_i8: operands must be 2-D with unit inner stride")
    M = a.shape[0]
    N = w.shape[0] if N is None else int(N)
    lda, ldw = (a.stride(0) if M > 1 else a.shape[1]), (w.stride(0) if w.shape[0] > 1 else w.shape[1])
    if K > a.shape[1] or K > w.shape[1]:
        raise ValueError("gemm_i8: K exceeds the operand width")
    dev = a.device
    if out_kind == QVIT_OUT_NONE:
        out, ldo = None, N
    elif out is None:
        if out_kind == QVIT_OUT_F16X2:
            ldo = 2 * N if ldo is None else int(ldo)          # two fp16 planes: hi in [0, N), lo in [ldo/2, ldo/2 + N)
        ldo = N if ldo is None else int(ldo)
        out = torch.empty((M, ldo), dtype=_OUT_DTYPE[out_kind], device=dev)
        if ldo > N and out_kind == QVIT_OUT_I8:
            out[:, N:].zero_()
    else:
        if out.dtype != _OUT_DTYPE[out_kind] or out.dim() != 2 or out.shape[0] != M or out.stride(1) != 1:
            raise ValueError("gemm_i8: bad `out`")
        ldo = out.stride(0) if M > 1 else out.shape[1]
    epi = _lib.Epilogue()
    epi.out_kind, epi.act, epi.scale_const = out_kind, act, float(scale_const)
    epi.acc_abs_max = int(min(max(int(acc_abs_max), 0), 2**31 - 1))
    keep = []

    def sp(v, what):
        if v is None:
            return None
        t_ = _scalar_param(v, dev, what)
        keep.append(t_)
        return t_.data_ptr()

    epi.scale_a, epi.scale_w = sp(scale_a, "scale_a"), sp(scale_w, "scale_w")
    if col_scale is not None:
        col_scale = _f32c(col_scale, "col_scale")
        if col_scale.numel() != N:
            raise ValueError("gemm_i8: col_scale must have N elements")
    if bias is not None:
        bias = _f32c(bias.detach(), "bias")
        if bias.numel() != N:
            raise ValueError("gemm_i8: bias must have N elements")
    epi.col_scale, epi.bias = _lib.ptr(col_scale), _lib.ptr(bias)
    if residual is not None:
        residual = _f32c(residual, "residual")
        if residual.dim() != 2 or residual.shape[0] != M or residual.shape[1] < N:
            raise ValueError("gemm_i8: residual must be [M, >=N] fp32")
        epi.residual, epi.ld_res = residual.data_ptr(), residual.stride(0) if M > 1 else residual.shape[1]
    if out_kind == QVIT_OUT_I8:
        if next_q is None:
            raise ValueError("gemm_i8: QVIT_OUT_I8 needs next_q=(d, q_m, t)")
        epi.next_d, epi.next_qm = sp(next_q[0], "next_d"), sp(next_q[1], "next_qm")
        epi.next_t = sp(next_q[2], "next_t") if len(next_q) > 2 else None
    epi.flags = _lib.ptr(flags)
    _lib.check(_lib.lib().qvit_gemm_i8(_lib.ptr(a), lda, 1 if a.dtype == torch.uint8 else 0, _lib.ptr(w), ldw, M, N, int(K),
                                       _lib.ptr(out), ldo, C.byref(epi), backend, _lib.stream()), "qvit_gemm_i8")
    return out


def _pad64(n: int) -> int:
    return (int(n) + 63) // 64 * 64


def split3_bf16(x: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """Exact 3-way bf16 split of a 2-D fp32 matrix, planes side by side: [R, 3*pad64(C)] or (transposed) [C, 3*pad64(R)]."""
    x = _f32c(x, "split3_bf16")
    R, Cc = x.shape
    pc = _pad64(R if transpose else Cc)
    out = torch.empty((Cc if transpose else R, 3 * pc), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().qvit_split3_bf16(_lib.ptr(x), R, Cc, x.stride(0), 1 if transpose else 0, _lib.ptr(out), pc, _lib.stream()),
               "qvit_split3_bf16")
    return out


def grad_prep(g: torch.Tensor, want_rows: bool = True, want_colsum: bool = True, want_trans: bool = True):
    """One pass over the output gradient g [M, N] of a QAT linear layer: (row planes [M, 3*pad64(N)] | None, transposed planes
    [N, 3*pad64(M)] | None, column sums [N] | None) - the operands of the two gradient GEMMs and the bias gradient."""
    g = _f32c(g, "grad_prep")
    M, N = g.shape
    mp, np_ = _pad64(M), _pad64(N)
    if not (want_rows or want_trans):
        raise ValueError("grad_prep: nothing to produce")
    rows = torch.empty((M, 3 * np_), dtype=torch.bfloat16, device=g.device) if want_rows else None
    trans = torch.empty((N, 3 * mp), dtype=torch.bfloat16, device=g.device) if want_trans else None
    partial = colsum = None
    if want_colsum:
        partial = torch.empty(((mp // 64 + 3) // 4, N), dtype=torch.float32, device=g.device)
        colsum = torch.empty((N,), dtype=torch.float32, device=g.device)
    _lib.check(_lib.lib().qvit_grad_prep(_lib.ptr(g), M, N, g.stride(0), _lib.ptr(rows), np_, _lib.ptr(trans), mp, _lib.ptr(partial),
                                         _lib.ptr(colsum), _lib.stream()), "qvit_grad_prep")
    return rows, trans, colsum


def codes_to_bf16_t(codes: torch.Tensor, cols: int) -> torch.Tensor:
    """int8 codes [R, >=cols] -> bf16 [cols, pad64(R)] (transposed, zero padded)."""
    _lib.require_cuda(codes)
    R = codes.shape[0]
    out = torch.empty((cols, _pad64(R)), dtype=torch.bfloat16, device=codes.device)
    _lib.check(_lib.lib().qvit_codes_to_bf16_t(_lib.ptr(codes), R, int(cols), codes.stride(0), _lib.ptr(out), out.shape[1], _lib.stream()),
               "qvit_codes_to_bf16_t")
    return out


def codes_to_bf16(codes: torch.Tensor, cols: int) -> torch.Tensor:
    """int8 codes [R, >=cols] -> bf16 [R, pad64(cols)] (same orientation, zero padded)."""
    _lib.require_cuda(codes)
    R = codes.shape[0]
    out = torch.empty((R, _pad64(cols)), dtype=torch.bfloat16, device=codes.device)
    _lib.check(_lib.lib().qvit_codes_to_bf16(_lib.ptr(codes), R, int(cols), codes.stride(0), _lib.ptr(out), out.shape[1], _lib.stream()),
               "qvit_codes_to_bf16")
    return out


def gemm_bf16_split_t(g_planes: torch.Tensor, x_bf16: torch.Tensor, N_out: int, K_in: int, planes: int = 3, scale=None) -> torch.Tensor:
    """out[N_out, K_in] fp32 = |scale| * sum_p G_p^T @ X: the weight-gradient GEMM straight from the ROW planes of the output
    gradient ([tokens, 3*pad64(N_out)]) and the bf16 codes [tokens, >= K_in] - both read as MN-major tcgen05 operands."""
    _lib.require_cuda(g_planes, x_bf16)
    tokens = g_planes.shape[0]
    if x_bf16.shape[0] != tokens:
        raise ValueError("gemm_bf16_split_t: operand row counts differ")
    out = torch.empty((N_out, (K_in + 3) // 4 * 4), dtype=torch.float32, device=g_planes.device)
    epi = _lib.Epilogue()
    epi.out_kind, epi.act, epi.scale_const = QVIT_OUT_F32, QVIT_ACT_NONE, 1.0
    keep = None
    if scale is not None:
        keep = _scalar_param(scale, g_planes.device, "scale")
        epi.scale_a = keep.data_ptr()
    # (small outputs: a workspace lets the kernel split the long contraction four ways with a reproducible sum)
    ws = torch.empty_like(out) if out.numel() <= (1 << 20) else None
    _lib.check(_lib.lib().qvit_gemm_bf16_split_t(_lib.ptr(g_planes), g_planes.stride(0), planes, g_planes.shape[1] // 3, _lib.ptr(x_bf16),
                                                 x_bf16.stride(0), tokens, int(N_out), int(K_in), _lib.ptr(out), out.stride(0),
                                                 _lib.ptr(ws), C.byref(epi), _lib.stream()), "qvit_gemm_bf16_split_t")
    return out if out.shape[1] == K_in else out[:, :K_in]


def gemm_bf16_split(a_planes: torch.Tensor, b: torch.Tensor, K: int, planes: int = 3, scale=None,
                    out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                    residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M, N] fp32 = |scale| * sum_p A_p[M, :K] @ B[N, :K]^T (+ bias[n]) (+ residual[m, n]) on tcgen05 kind::f16 (fp32
    accumulation in TMEM)."""
    _lib.require_cuda(a_planes, b)
    M, N = a_planes.shape[0], b.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a_planes.device)
    epi = _lib.Epilogue()
    epi.out_kind, epi.act, epi.scale_const = QVIT_OUT_F32, QVIT_ACT_NONE, 1.0
    keep = None
    if scale is not None:
        keep = _scalar_param(scale, a_planes.device, "scale")
        epi.scale_a = keep.data_ptr()
    if bias is not None:
        bias = _f32c(bias.detach(), "bias")
        epi.bias = bias.data_ptr()
    if residual is not None:
        _lib.require_cuda(residual)
        if residual.dtype != torch.float32 or residual.dim() != 2 or residual.stride(1) != 1:
            raise TypeError("gemm_bf16_split: residual must be a 2-D fp32 tensor with unit inner stride")
        epi.residual, epi.ld_res = residual.data_ptr(), residual.stride(0)
    _lib.check(_lib.lib().qvit_gemm_bf16_split(_lib.ptr(a_planes), a_planes.stride(0), planes, _lib.ptr(b), b.stride(0), M, N, int(K),
                                               _lib.ptr(out), out.stride(0), C.byref(epi), _lib.stream()), "qvit_gemm_bf16_split")
    return out


def matmul_f32_tc(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, *, a_transposed: bool = False,
                  b_transposed: bool = False) -> torch.Tensor:
    """fp32-equivalent  op(a) @ op(b)^T (+ bias)  on the tensor cores, for the layers whose codes do not fit the int8 pipe
    (> 8-bit quantizers, weight-only mode - what GETA trains through before it has walked the bit width down, train.py:247-250).

    Both fp32 operands are split EXACTLY into three bf16 planes (8 + 8 + 8 significant bits) and the six plane products
    with index sum <= 2 are accumulated in fp32 TMEM (dropped terms <= 2^-24 relative): three launches of the split-bf16
    GEMM - B plane 0 against A planes {0, 1, 2}, plane 1 against {0, 1}, plane 2 against {0} - the later ones adding onto the
    earlier result through the epilogue's residual.  op(x) = x^T when *_transposed (the split kernel transposes while it
    splits).  a: [M, K] (or [K, M] transposed), b: [N, K] (or [K, N] transposed) -> [M, N] fp32."""
    ap = split3_bf16(a, transpose=a_transposed)
    bp = split3_bf16(b, transpose=b_transposed)
    K = a.shape[0] if a_transposed else a.shape[1]
    Kb = b.shape[0] if b_transposed else b.shape[1]
    if K != Kb:
        raise ValueError("matmul_f32_tc: contraction lengths differ")
    kp = _pad64(K)
    M, N = ap.shape[0], bp.shape[0]
    buf = torch.empty((M, (N + 3) // 4 * 4), dtype=torch.float32, device=ap.device)      # TMA stores need a 16-byte row pitch
    out = buf[:, :N]
    gemm_bf16_split(ap, bp[:, 0:kp], K, planes=3, out=out, bias=bias)
    gemm_bf16_split(ap, bp[:, kp:2 * kp], K, planes=2, out=out, residual=out)
    gemm_bf16_split(ap, bp[:, 2 * kp:3 * kp], K, planes=1, out=out, residual=out)
    return out if buf.shape[1] == N else out.contiguous()


# ------------------------------------------------------------------------------------------ UltraNet (DoReFa)
def ultra_weight_codes(w: torch.Tensor, w_bit: int, export_rounding: bool = False) -> torch.Tensor:
    """int8 codes of weight_quantize_fn(w_bit).forward (QU:38-56), same shape as w; values = codes / (2^(b-1)-1)."""
    w = _f32c(w.detach(), "ultra_weight_codes")
    mx = torch.empty(1, dtype=torch.float32, device=w.device)
    codes = torch.empty(w.shape, dtype=torch.int8, device=w.device)
    L = _lib.lib()
    _lib.check(L.qvit_ultra_tanh_absmax(_lib.ptr(w), w.numel(), _lib.ptr(mx), _lib.stream()), "qvit_ultra_tanh_absmax")
    _lib.check(L.qvit_ultra_quantize_weight(_lib.ptr(w), w.numel(), int(w_bit), 1 if export_rounding else 0, _lib.ptr(mx), _lib.ptr(codes),
                                            _lib.stream()), "qvit_ultra_quantize_weight")
    return codes


def ultra_act(x: torch.Tensor, a_bit: int, want_codes: bool = True, want_values: bool = False):
    """activation_quantize_fn(a_bit).forward (QU:66-73): (uint8 codes | None, fp32 values | None)."""
    x = _f32c(x, "ultra_act")
    codes = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_codes else None
    vals = torch.empty_like(x) if want_values else None
    _lib.check(_lib.lib().qvit_ultra_quantize_act(_lib.ptr(x), x.numel(), int(a_bit), _lib.ptr(codes), _lib.ptr(vals),
                                                  _lib.stream()), "qvit_ultra_quantize_act")
    return codes, vals


def uniform_quantize(x: torch.Tensor, k: int) -> torch.Tensor:
    """uniform_quantize(k).forward (QU:12-20)."""
    x = _f32c(x, "uniform_quantize")
    out = torch.empty_like(x)
    _lib.check(_lib.lib().qvit_uniform_quantize(_lib.ptr(x), x.numel(), int(k), _lib.ptr(out), _lib.stream()),
               "qvit_uniform_quantize")
    return out


def ultra_bn_act_pool_nchw(x: torch.Tensor, scale, bias, levels: int, pool: bool, ldc: Optional[int] = None) -> torch.Tensor:
    """NCHW fp32 -> NHWC uint8 codes round(clamp(x*scale+bias, 0, 1) * levels), optionally 2x2 max-pooled.
    Channel pitch ldc >= C; padding channels are zero."""
    x = _f32c(x, "ultra_bn_act_pool_nchw")
    B, Cc, H, W = x.shape
    ldc = Cc if ldc is None else int(ldc)
    OH, OW = (H // 2, W // 2) if pool else (H, W)
    alloc = torch.zeros if ldc > Cc else torch.empty
    out = alloc((B, OH, OW, ldc), dtype=torch.uint8, device=x.device)
    _lib.check(_lib.lib().qvit_ultra_bn_act_pool_nchw(_lib.ptr(x), B, Cc, H, W, _lib.ptr(scale), _lib.ptr(bias), int(levels),
                                                      1 if pool else 0, _lib.ptr(out), ldc, _lib.stream()),
               "qvit_ultra_bn_act_pool_nchw")
    return out


def conv2d_f32_wcodes(x, w_codes, w_levels: float, bias, stride, padding, dilation) -> torch.Tensor:
    """Conv2d_Q.forward (QU:85-89) on a raw fp32 input, weights as integer codes (groups == 1)."""
    x = _f32c(x, "conv2d_f32_wcodes")
    _lib.require_cuda(w_codes)
    B, Cc, H, W = x.shape
    O, Ci, kh, kw = w_codes.shape
    if Ci != Cc:
        raise ValueError("conv2d_f32_wcodes: channel mismatch (groups must be 1)")
    sh, sw = stride
    ph, pw = padding
    dh, dw = dilation
    OH = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    OW = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    y = torch.empty((B, O, OH, OW), dtype=torch.float32, device=x.device)
    b = None if bias is None else _f32c(bias.detach(), "bias")
    _lib.check(_lib.lib().qvit_conv2d_f32_wcodes(_lib.ptr(x), B, Cc, H, W, _lib.ptr(w_codes.contiguous()), O, kh, kw, sh, sw,
                                                 ph, pw, dh, dw, float(w_levels), _lib.ptr(b), _lib.ptr(y), _lib.stream()),
               "qvit_conv2d_f32_wcodes")
    return y


def ultra_conv_bn_act(in_codes: torch.Tensor, w_codes_ohwi: torch.Tensor, pad: int, acc_scale: float,
                      bn_scale: Optional[torch.Tensor], bn_bias: Optional[torch.Tensor], out_levels: int, pool: bool,
                      f32_out: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One fused integer UltraNet layer.  in_codes [B,H,W,C] uint8, w_codes [O,kh,kw,C] int8.
    Returns uint8 NHWC codes (pooled 2x2 if `pool`) or, with f32_out, the un-quantised NCHW fp32 map."""
    _lib.require_cuda(in_codes, w_codes_ohwi)
    if in_codes.dtype != torch.uint8 or w_codes_ohwi.dtype != torch.int8:
        raise TypeError("ultra_conv_bn_act: uint8 activations and int8 weights expected")
    in_codes, w_codes_ohwi = in_codes.contiguous(), w_codes_ohwi.contiguous()
    B, H, W, Cc = in_codes.shape
    O, kh, kw, Ci = w_codes_ohwi.shape
    if Ci != Cc:
        raise ValueError("ultra_conv_bn_act: channel mismatch")
    OH, OW = H + 2 * pad - kh + 1, W + 2 * pad - kw + 1
    dev = in_codes.device
    if f32_out:
        if out is None:
            out = torch.empty((B, O, OH, OW), dtype=torch.float32, device=dev)
        oc, of = None, out
    else:
        if out is None:
            out = torch.empty((B, OH // 2, OW // 2, O) if pool else (B, OH, OW, O), dtype=torch.uint8, device=dev)
        oc, of = out, None
    _lib.check(_lib.lib().qvit_ultra_conv_bn_act(_lib.ptr(in_codes), B, H, W, Cc, _lib.ptr(w_codes_ohwi), O, kh, kw, int(pad),
                                                 float(acc_scale), _lib.ptr(bn_scale), _lib.ptr(bn_bias), int(out_levels),
                                                 1 if pool else 0, _lib.ptr(oc), _lib.ptr(of), _lib.stream()),
               "qvit_ultra_conv_bn_act")
    return out


def pack_conv_weights_tc(w_codes_ohwi: torch.Tensor) -> torch.Tensor:
    """[O, kh, kw, C] int8 weight codes -> the [O_pad, K_pad] operand of ultra_conv_tc (flattened (tap, channel) order, O padded to
    16 and K to 128 with zeros).  One-time, per model."""
    O = w_codes_ohwi.shape[0]
    flat = w_codes_ohwi.reshape(O, -1)
    K = flat.shape[1]
    out = torch.zeros(((O + 15) // 16 * 16, (K + 127) // 128 * 128), dtype=torch.int8, device=w_codes_ohwi.device)
    out[:O, :K] = flat
    return out


def ultra_conv_tc_supported(C: int, O: int, kh: int, kw: int) -> bool:
    K_pad = (kh * kw * C + 127) // 128 * 128
    return C in (16, 32, 64, 128) and O <= 256 and (K_pad // 128) * ((O + 15) // 16 * 16 + 128) * 128 + 2048 <= 227 * 1024


def ultra_conv_tc(in_codes: torch.Tensor, w_packed: torch.Tensor, O: int, kh: int, kw: int, pad: int, acc_scale: float,
                  bn_scale: Optional[torch.Tensor], bn_bias: Optional[torch.Tensor], out_levels: int, pool: bool,
                  f32_out: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One fused integer UltraNet layer as an implicit GEMM on tcgen05 (include/qvit_b200.h: qvit_ultra_conv_tc).
    in_codes [B,H,W,C] uint8, w_packed from pack_conv_weights_tc.  Same outputs as ultra_conv_bn_act."""
    _lib.require_cuda(in_codes, w_packed)
    if in_codes.dtype != torch.uint8 or w_packed.dtype != torch.int8:
        raise TypeError("ultra_conv_tc: uint8 activations and int8 weights expected")
    in_codes = in_codes.contiguous()
    B, H, W, Cc = in_codes.shape
    OH, OW = H + 2 * pad - kh + 1, W + 2 * pad - kw + 1
    dev = in_codes.device
    if f32_out:
        if out is None:
            out = torch.empty((B, O, OH, OW), dtype=torch.float32, device=dev)
        oc, of = None, out
    else:
        if out is None:
            out = torch.empty((B, OH // 2, OW // 2, O) if pool else (B, OH, OW, O), dtype=torch.uint8, device=dev)
        oc, of = out, None
    _lib.check(_lib.lib().qvit_ultra_conv_tc(_lib.ptr(in_codes), B, H, W, Cc, _lib.ptr(w_packed), int(O), kh, kw, int(pad),
                                             float(acc_scale), _lib.ptr(bn_scale), _lib.ptr(bn_bias), int(out_levels),
                                             1 if pool else 0, _lib.ptr(oc), _lib.ptr(of), _lib.stream()), "qvit_ultra_conv_tc")
    return out


def conv2d_i8_tc(a_codes_nhwc: torch.Tensor, w_packed: torch.Tensor, O: int, kernel, stride, pad: int, dilation, scale_a, scale_w,
                 bias: Optional[torch.Tensor]) -> torch.Tensor:
    """QuantizeConv2d as an implicit GEMM on tcgen05: signed int8 NHWC codes [B,H,W,C] x packed weight codes -> fp32 NCHW."""
    _lib.require_cuda(a_codes_nhwc, w_packed)
    if a_codes_nhwc.dtype != torch.int8 or not a_codes_nhwc.is_contiguous():
        raise TypeError("conv2d_i8_tc: contiguous int8 NHWC codes expected")
    B, H, W, Cc = a_codes_nhwc.shape
    kh, kw = kernel
    OH = (H + 2 * pad - dilation[0] * (kh - 1) - 1) // stride[0] + 1
    OW = (W + 2 * pad - dilation[1] * (kw - 1) - 1) // stride[1] + 1
    dev = a_codes_nhwc.device
    out = torch.empty((B, O, OH, OW), dtype=torch.float32, device=dev)
    sa, sw_ = _scalar_param(scale_a, dev, "scale_a"), _scalar_param(scale_w, dev, "scale_w")
    b = None if bias is None else _f32c(bias.detach(), "bias")
    _lib.check(_lib.lib().qvit_conv2d_i8_tc(_lib.ptr(a_codes_nhwc), B, H, W, Cc, _lib.ptr(w_packed), int(O), kh, kw, stride[0], stride[1],
                                            int(pad), dilation[0], dilation[1], _lib.ptr(sa), _lib.ptr(sw_), _lib.ptr(b), _lib.ptr(out),
                                            _lib.stream()), "qvit_conv2d_i8_tc")
    return out


def bn_fold(gamma, beta, mean, var, eps: float, mode: int = 0):
    """(scale, bias) of an eval BatchNorm.  mode 0: nn.BatchNorm2d (eps inside sqrt); mode 1: QZ:34-46 (eps outside)."""
    gamma, beta, mean, var = (_f32c(t.detach(), "bn_fold") for t in (gamma, beta, mean, var))
    Cn = gamma.numel()
    scale, bias = torch.empty_like(gamma), torch.empty_like(gamma)
    _lib.check(_lib.lib().qvit_bn_fold(_lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(mean), _lib.ptr(var), float(eps), int(mode), Cn,
                                       _lib.ptr(scale), _lib.ptr(bias), _lib.stream()), "qvit_bn_fold")
    return scale, bias


def bn_act_quantize_int(gamma, beta, mean, var, eps: float, w_bit=2, in_bit=4, out_bit=4, l_shift=4):
    """bn_act_quantize_int (QZ:68-89) -> (inc, bias) int32, computed in the dtype of the inputs (fp32 or fp64)."""
    ts = [gamma, beta, mean, var]
    _lib.require_cuda(*ts)
    is64 = all(t.dtype == torch.float64 for t in ts)
    if not is64:
        ts = [t.to(torch.float32) for t in ts]
    ts = [t.contiguous() for t in ts]
    Cn = ts[0].numel()
    inc = torch.empty(Cn, dtype=torch.int32, device=ts[0].device)
    bias = torch.empty(Cn, dtype=torch.int32, device=ts[0].device)
    _lib.check(_lib.lib().qvit_bn_act_quantize_int(*[_lib.ptr(t) for t in ts], 1 if is64 else 0, float(eps), int(w_bit),
                                                   int(in_bit), int(out_bit), int(l_shift), Cn, _lib.ptr(inc),
                                                   _lib.ptr(bias), _lib.stream()), "qvit_bn_act_quantize_int")
    return inc, bias


def pack_int4(codes: torch.Tensor) -> torch.Tensor:
    """Little-endian nibble pack (array_to_string, qnn_mem_process.py:11-24): last dim must be even."""
    _lib.require_cuda(codes)
    codes = codes.to(torch.int8).contiguous()
    if codes.shape[-1] % 2:
        raise ValueError("pack_int4: last dimension must be even")
    out = torch.empty((*codes.shape[:-1], codes.shape[-1] // 2), dtype=torch.uint8, device=codes.device)
    _lib.check(_lib.lib().qvit_pack_int4(_lib.ptr(codes), codes.numel(), _lib.ptr(out), _lib.stream()), "qvit_pack_int4")
    return out


def unpack_int4(packed: torch.Tensor, signed: bool = True) -> torch.Tensor:
    _lib.require_cuda(packed)
    packed = packed.to(torch.uint8).contiguous()
    out = torch.empty((*packed.shape[:-1], packed.shape[-1] * 2), dtype=torch.int8, device=packed.device)
    _lib.check(_lib.lib().qvit_unpack_int4(_lib.ptr(packed), out.numel(), 1 if signed else 0, _lib.ptr(out), _lib.stream()),
               "qvit_unpack_int4")
    return out


def pack_hls_weights(codes: torch.Tensor, w_bit: int, simd: int, pe: int) -> torch.Tensor:
    """FPGA (HLS) weight layout of the reference exporter (qnn_mem_process.py:84-130, 152-157).

    codes: [O, I, kh, kw] integer weight codes (weight_quantize_int, QZ:24-31).  Returns int64 [pe, tiles] whose bit
    patterns are the reference's words (runs of `simd` codes in (kh, kw, I) order, element e in bits [w_bit*e, +w_bit),
    two's complement; word (oc, j) at [oc % pe, (oc // pe) * runs + j]); `& (2**64 - 1)` gives the unsigned value."""
    _lib.require_cuda(codes)
    if codes.dim() != 4:
        raise ValueError("pack_hls_weights: codes must be [O, I, kh, kw]")
    c8 = codes.to(torch.int8).contiguous()
    O, I, kh, kw = c8.shape
    if O % pe != 0:
        raise AssertionError("out_ch mod pe must 0")             # the reference's own assert (qnn_mem_process.py:86)
    runs = (kh * kw * I + simd - 1) // simd
    out = torch.empty((pe, runs * (O // pe)), dtype=torch.int64, device=c8.device)
    _lib.check(_lib.lib().qvit_pack_hls_weights(_lib.ptr(c8), O, I, kh, kw, int(w_bit), int(simd), int(pe), _lib.ptr(out),
                                                _lib.stream()), "qvit_pack_hls_weights")
    return out
