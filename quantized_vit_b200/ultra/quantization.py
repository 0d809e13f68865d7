"""GPU counterparts of the NumPy export helpers in ``4-bit quantization/quantization.py`` (QZ): integer weight
codes, BN -> (scale, bias) fold and the integer (inc, bias) thresholds.  One-time, offline-style work; it exists
so that the packed integer parameters the FPGA flow consumes can be produced from device-resident weights
(SURVEY.md section 8f rank 3).  All functions take and return torch CUDA tensors."""
from __future__ import annotations

import torch

from .. import ops


def weight_quantize_int(weight: torch.Tensor, bit: int) -> torch.Tensor:
    """QZ:24-31: round(tanh(w) / max|tanh(w)| * (2^(bit-1)-1)) as int32 (always rounds, also at bit == 2 where the
    torch forward takes sign(), QU:15-16)."""
    ops._lib.require_cuda(weight)
    return ops.ultra_weight_codes(weight.float(), bit, export_rounding=True).to(torch.int32)


def weight_quantize_float(weight: torch.Tensor, bit: int) -> torch.Tensor:
    """QZ:13-19."""
    codes = weight_quantize_int(weight, bit).to(torch.float32)
    return codes / torch.full((), float(2 ** (bit - 1) - 1), dtype=torch.float32, device=codes.device)


def bn_act_w_bias_float(gamma, beta, mean, var, eps):
    """QZ:34-46: w = gamma / (sqrt(var) + eps), b = beta - mean / (sqrt(var) + eps) * gamma (eps OUTSIDE the sqrt)."""
    return ops.bn_fold(gamma, beta, mean, var, eps, mode=1)


def bn_act_quantize_int(gamma, beta, mean, var, eps, w_bit=2, in_bit=4, out_bit=4, l_shift=4):
    """QZ:68-89 -> (inc, bias) int32."""
    return ops.bn_act_quantize_int(gamma, beta, mean, var, eps, w_bit, in_bit, out_bit, l_shift)


def w_to_hls_array(weight_codes: torch.Tensor, w_bit: int, simd: int, pe: int) -> torch.Tensor:
    """QNNLayerMemProcess.conv + w_to_hls_array (qnn_mem_process.py:84-130, 152-157): integer weight codes [O, I, kh, kw]
    -> int64 [pe, tiles] whose bit patterns are the reference's words of ``simd * w_bit`` bits (``& (2**64 - 1)`` gives the
    unsigned value the reference prints into param.h)."""
    return ops.pack_hls_weights(weight_codes, w_bit, simd, pe)


def inc_bias_to_hls_array(inc: torch.Tensor, bias: torch.Tensor, pe: int):
    """qnn_mem_process.py:133-143: per-channel thresholds -> [pe, channels // pe] (channel c at [c % pe, c // pe])."""
    return inc.reshape(-1, pe).t().contiguous(), bias.reshape(-1, pe).t().contiguous()
