"""Oracle (TEST INFRASTRUCTURE) for the DoReFa-style fixed-grid path of UltraNet.

Restates:
  * uniform_quantize(k)                       QU:8-27     (torch)  /  QZ:5-9 (numpy)
  * weight_quantize_fn(w_bit).forward         QU:38-56
  * activation_quantize_fn(a_bit).forward     QU:66-73
  * Conv2d_Q.forward / Linear_Q.forward       QU:85-89 / QU:217-220
  * BatchNorm2d_Q.forward (folded, quantised) QU:107-130
  * weight_quantize_int / weight_quantize_float      QZ:24-31 / QZ:13-19   (NumPy, float64)
  * bn_act_w_bias_float                       QZ:34-46    (eps OUTSIDE the sqrt)
  * bn_act_quantize_int                       QZ:68-89
  * array_to_string (little-endian nibble pack)  qnn_mem_process.py:11-24
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------- torch path (QU)

def uniform_quantize(x: torch.Tensor, k: int) -> torch.Tensor:
    """QU:12-20: identity at 32 bits, sign at 1 bit, else round(x*n)/n with n = 2^k - 1."""
    if k == 32:
        return x
    if k == 1:
        return torch.sign(x)
    n = float(2 ** k - 1)
    return torch.round(x * n) / n


def ultra_weight_values(w: torch.Tensor, w_bit: int) -> torch.Tensor:
    """weight_quantize_fn.forward (QU:38-56)."""
    w = w.detach().to(torch.float32).cpu()
    if w_bit == 32:
        return w
    if w_bit == 1:
        e = torch.mean(torch.abs(w))
        # QU:36 builds uniform_quantize(k = w_bit - 1 = 0): n = 2^0 - 1 = 0, so round(x*0)/0 = NaN -
        # the reference's 1-bit weight path returns NaN everywhere (QU:44-46 with QU:18-19).  Kept as is.
        return (uniform_quantize(w / e, 0) + 1) / 2 * e
    v = torch.tanh(w)                                            # QU:50
    v = v / torch.max(torch.abs(v))                              # QU:53 per-tensor max
    return uniform_quantize(v, w_bit - 1)                        # QU:55


def ultra_weight_codes(w: torch.Tensor, w_bit: int) -> torch.Tensor:
    """Signed integer codes in [-(2^(b-1)-1), 2^(b-1)-1] with values == codes / (2^(b-1)-1).

    Reference quirk kept: at w_bit == 2 the torch path calls uniform_quantize(k=1), i.e. sign()
    (QU:15-16), so codes are +-1 (0 only for an exact 0 weight) - whereas the NumPy export
    (QZ:24-31) rounds to {-1,0,1}.  The torch path is the forward the modules run."""
    assert 2 <= w_bit <= 8
    w = w.detach().to(torch.float32).cpu()
    n = float(2 ** (w_bit - 1) - 1)
    v = torch.tanh(w)
    v = v / torch.max(torch.abs(v))
    if w_bit == 2:
        return torch.sign(v).to(torch.int64)
    return torch.round(v * n).to(torch.int64)


def ultra_act_values(x: torch.Tensor, a_bit: int) -> torch.Tensor:
    """activation_quantize_fn.forward (QU:66-73): clamp to [0,1] then the 2^a-1 grid."""
    x = x.detach().to(torch.float32).cpu()
    if a_bit == 32:
        return x
    return uniform_quantize(torch.clamp(x, 0, 1), a_bit)


def ultra_act_codes(x: torch.Tensor, a_bit: int) -> torch.Tensor:
    """Unsigned codes 0..2^a-1 with values == codes / (2^a-1)."""
    x = x.detach().to(torch.float32).cpu()
    n = float(2 ** a_bit - 1)
    return torch.round(torch.clamp(x, 0, 1) * n).to(torch.int64)


def conv2d_q_forward(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], w_bit: int,
                     stride=1, padding=0, dilation=1, groups=1) -> torch.Tensor:
    """Conv2d_Q.forward (QU:85-89).  The input is NOT quantised here."""
    b = None if bias is None else bias.detach().float().cpu()
    return F.conv2d(x.detach().float().cpu(), ultra_weight_values(w, w_bit), b, stride, padding, dilation, groups)


def linear_q_forward(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], w_bit: int) -> torch.Tensor:
    """Linear_Q.forward (QU:217-220)."""
    b = None if bias is None else bias.detach().float().cpu()
    return F.linear(x.detach().float().cpu(), ultra_weight_values(w, w_bit), b)


def batchnorm2d_q_scale_bias(gamma, beta, mean, var, eps: float, w_bit: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """The per-channel (w_q, b_q) BatchNorm2d_Q applies (QU:107-123): fold with sqrt(var)+eps,
    clamp to [-1,1], map to [0,1], quantise on the 2^w_bit-1 grid, map back."""
    w = gamma / (torch.sqrt(var) + eps)
    b = beta - (mean / (torch.sqrt(var) + eps)) * gamma
    w_q = 2 * uniform_quantize(torch.clamp(w, -1, 1) / 2 + 0.5, w_bit) - 1
    b_q = 2 * uniform_quantize(torch.clamp(b, -1, 1) / 2 + 0.5, w_bit) - 1
    return w_q, b_q


def batchnorm2d_q_forward(x, gamma, beta, mean, var, eps: float, w_bit: int) -> torch.Tensor:
    """BatchNorm2d_Q.forward (QU:125-130): batch_norm with zero mean, unit var, eps 0.  The ATen op is called directly:
    torch >= 2 put a Python guard (eps <= 0 raises) in front of it that did not exist when the reference was written."""
    w_q, b_q = batchnorm2d_q_scale_bias(gamma, beta, mean, var, eps, w_bit)
    return torch.batch_norm(x, w_q, b_q, mean * 0, torch.sign(torch.abs(var) + 1), False, 0.1, eps * 0, False)


def batchnorm1d_q_forward(x, gamma, beta, mean, var, eps: float, w_bit: int) -> torch.Tensor:
    """BatchNorm1d_Q.forward (QU:185-206): the folded scale is quantised (w_q, QU:196) but the UN-quantised fold (w, b) is
    what reaches batch_norm (QU:203-206); zero mean, unit var (sign(var + 1)), eps 0."""
    w = gamma / (torch.sqrt(var) + eps)
    b = beta - (mean / (torch.sqrt(var) + eps)) * gamma
    return torch.batch_norm(x, w, b, mean * 0, torch.sign(var + 1), False, 0.1, eps * 0, False)


# ---------------------------------------------------------------- NumPy export path (QZ)

def np_uniform_quantize(v: np.ndarray, bit: int) -> np.ndarray:
    n = float(2 ** bit - 1)
    return np.round(v * n) / n                                   # QZ:5-9


def np_weight_quantize_float(w: np.ndarray, bit: int) -> np.ndarray:
    v = np.tanh(w)
    v = v / np.max(np.abs(v))
    return np_uniform_quantize(v, bit - 1)                       # QZ:13-19


def np_weight_quantize_int(w: np.ndarray, bit: int) -> np.ndarray:
    v = np.tanh(w)
    v = v / np.max(np.abs(v))
    return np.round(v * (2 ** (bit - 1) - 1)).astype(np.int32)   # QZ:24-31


def np_bn_fold(gamma, beta, mean, var, eps) -> Tuple[np.ndarray, np.ndarray]:
    """bn_act_w_bias_float (QZ:34-46): w = gamma/(sqrt(var)+eps), b = beta - mean/(sqrt(var)+eps)*gamma."""
    w = gamma / (np.sqrt(var) + eps)
    b = beta - (mean / (np.sqrt(var) + eps) * gamma)
    return w, b


def np_bn_act_quantize_int(gamma, beta, mean, var, eps, w_bit=2, in_bit=4, out_bit=4, l_shift=4):
    """bn_act_quantize_int (QZ:68-89): integer (inc, bias) thresholds with a 2^l_shift gain."""
    w, b = np_bn_fold(gamma, beta, mean, var, eps)
    n = 2 ** (w_bit - 1 + in_bit + l_shift) / ((2 ** (w_bit - 1) - 1) * (2 ** in_bit - 1))
    inc = np.round((2 ** out_bit - 1) * n * w).astype(np.int32)
    bias = np.round((2 ** (w_bit - 1) - 1) * (2 ** in_bit - 1) * (2 ** out_bit - 1) * n * b).astype(np.int32)
    return inc, bias


def torch_bn_fold(gamma, beta, mean, var, eps):
    """nn.BatchNorm2d eval as (scale, bias): s = gamma/sqrt(var+eps), b = beta - mean*s
    (what MM:74.. applies after each conv; eps INSIDE the sqrt, unlike QZ:43-45)."""
    s = gamma / torch.sqrt(var + eps)
    return s, beta - mean * s


# ---------------------------------------------------------------- pack convention

def pack_words(codes: Sequence[int], elem_bit: int) -> int:
    """array_to_string (qnn_mem_process.py:11-24): element i occupies bits [elem_bit*i, elem_bit*(i+1)),
    negatives in two's complement; returns an unbounded Python int."""
    word = 0
    for i, c in enumerate(codes):
        c = int(c)
        if c < 0:
            c += 1 << elem_bit
        word |= c << (elem_bit * i)
    return word


def pack_int4_bytes(codes: np.ndarray) -> np.ndarray:
    """Same convention, vectorised, to bytes: byte j = (codes[2j] & 15) | (codes[2j+1] & 15) << 4.
    Last axis must be even."""
    c = np.asarray(codes).astype(np.int64)
    assert c.shape[-1] % 2 == 0
    lo = c[..., 0::2] & 0xF
    hi = c[..., 1::2] & 0xF
    return (lo | (hi << 4)).astype(np.uint8)


def unpack_int4_bytes(packed: np.ndarray, signed: bool = True) -> np.ndarray:
    p = np.asarray(packed).astype(np.int64)
    lo = p & 0xF
    hi = (p >> 4) & 0xF
    out = np.stack([lo, hi], axis=-1).reshape(*p.shape[:-1], p.shape[-1] * 2)
    if signed:
        out = np.where(out >= 8, out - 16, out)
    return out.astype(np.int8)


# ---------------------------------------------------------------- FPGA (HLS) parameter layout

def hls_weight_words(codes: np.ndarray, w_bit: int, simd: int, pe: int) -> np.ndarray:
    """QNNLayerMemProcess.conv + w_to_hls_array (qnn_mem_process.py:84-130, 152-157): integer weight codes
    [O, I, kh, kw] -> reordered to (O, kh, kw, I), flattened per output channel, cut into runs of `simd` codes (the last
    run may be shorter), each run packed by pack_words; word (out_ch, j) lands at res[out_ch % pe][(out_ch // pe) * runs + j].
    Returns uint64 [pe, tiles] (simd * w_bit <= 64)."""
    codes = np.asarray(codes)
    o = codes.shape[0]
    assert o % pe == 0 and simd * w_bit <= 64
    flat = codes.transpose(0, 2, 3, 1).reshape(o, -1)
    h = flat.shape[1]
    runs = (h + simd - 1) // simd
    res = np.zeros((pe, runs * (o // pe)), dtype=np.uint64)
    for oc in range(o):
        for j in range(runs):
            res[oc % pe, (oc // pe) * runs + j] = np.uint64(pack_words(flat[oc, j * simd:(j + 1) * simd].tolist(), w_bit))
    return res


def hls_inc_bias(inc: np.ndarray, bias: np.ndarray, pe: int) -> Tuple[np.ndarray, np.ndarray]:
    """inc_bias_to_hls_array (qnn_mem_process.py:133-143): per-channel vectors -> [pe, channels // pe] (channel c at
    [c % pe, c // pe])."""
    return np.asarray(inc).reshape(-1, pe).T, np.asarray(bias).reshape(-1, pe).T
