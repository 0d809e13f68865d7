"""ctypes binding of libqvit_b200.so - the C-ABI boundary of the hot path (include/qvit_b200.h).

There is NO CPU fallback: if the library is missing the import of any product module that needs it raises,
and every entry point raises ``RuntimeError`` with ``qvit_last_error()`` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqvit_b200.so")

QVIT_OUT_I32, QVIT_OUT_F32, QVIT_OUT_BF16, QVIT_OUT_I8, QVIT_OUT_NONE, QVIT_OUT_F16X2 = 0, 1, 2, 3, 4, 5
QVIT_ACT_NONE, QVIT_ACT_GELU, QVIT_ACT_RELU = 0, 1, 2
QVIT_GEMM_AUTO, QVIT_GEMM_TCGEN05, QVIT_GEMM_SIMT = 0, 1, 2
QVIT_FLAG_NAN, QVIT_FLAG_OVERFLOW, QVIT_FLAG_NAN_GRAD = 1, 2, 4

_p = C.c_void_p
_i64 = C.c_int64
_i = C.c_int
_f = C.c_float


class Epilogue(C.Structure):
    """struct qvit_epilogue (include/qvit_b200.h)."""
    _fields_ = [("out_kind", C.c_int32), ("act", C.c_int32), ("scale_a", _p), ("scale_w", _p),
                ("scale_const", C.c_float), ("acc_abs_max", C.c_int32), ("col_scale", _p), ("bias", _p), ("residual", _p),
                ("ld_res", C.c_int64), ("next_d", _p), ("next_qm", _p), ("next_t", _p), ("flags", _p)]


# name -> (restype, argtypes); must list every symbol include/qvit_b200.h declares (tests/test_abi.py checks)
PROTOTYPES = {
    "qvit_abi_version": (_i, []),
    "qvit_last_error": (C.c_char_p, []),
    "qvit_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "qvit_quantize_sym": (_i, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _p, _p]),
    "qvit_fake_quantize_sym": (_i, [_p, _i64, _p, _p, _p, _p, _p]),
    "qvit_sym_backward": (_i, [_p, _p, _i64, _p, _p, _p, _f, _f, _p, _p, _p, _p]),
    "qvit_gelu_quantize_sym": (_i, [_p, _i64, _p, _p, _p, _p, _p, _p]),
    "qvit_gelu_sym_backward": (_i, [_p, _p, _i64, _p, _p, _p, _f, _f, _p, _p, _p, _p]),
    "qvit_embed_assemble": (_i, [_p, _p, _p, _i, _i, _i, _p, _p]),
    "qvit_absmax": (_i, [_p, _i64, _p, _p]),
    "qvit_im2col_quantize_sym": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i64, _p, _p]),
    "qvit_ultra_tanh_absmax": (_i, [_p, _i64, _p, _p]),
    "qvit_ultra_quantize_weight": (_i, [_p, _i64, _i, _i, _p, _p, _p]),
    "qvit_ultra_quantize_act": (_i, [_p, _i64, _i, _p, _p, _p]),
    "qvit_uniform_quantize": (_i, [_p, _i64, _i, _p, _p]),
    "qvit_ultra_bn_act_pool_nchw": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _i, _p, _i, _p]),
    "qvit_conv2d_f32_wcodes": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p]),
    "qvit_ultra_conv_bn_act": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _f, _p, _p, _i, _i, _p, _p, _p]),
    "qvit_ultra_conv_tc": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _f, _p, _p, _i, _i, _p, _p, _p]),
    "qvit_conv2d_i8_tc": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "qvit_bn_fold": (_i, [_p, _p, _p, _p, _f, _i, _i, _p, _p, _p]),
    "qvit_bn_act_quantize_int": (_i, [_p, _p, _p, _p, _i, C.c_double, _i, _i, _i, _i, _i, _p, _p, _p]),
    "qvit_pack_int4": (_i, [_p, _i64, _p, _p]),
    "qvit_unpack_int4": (_i, [_p, _i64, _i, _p, _p]),
    "qvit_pack_hls_weights": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "qvit_geta_quant_step": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _i, _f, _f, _f, _f, C.c_double, C.c_double, _f, _i, _f, _f,
                                  _f, _f, _f, _f, _p, _p]),
    "qvit_quant_sat_levels": (_i, [_p, _i, _p, _p]),
    "qvit_gemm_set_cta_group": (_i, [_i]),
    "qvit_gemm_read_profile": (_i, [_p, _i]),
    "qvit_gemm_i8": (_i, [_p, _i64, _i, _p, _i64, _i, _i, _i, _p, _i64, C.POINTER(Epilogue), _i, _p]),
    "qvit_split3_bf16": (_i, [_p, _i64, _i64, _i64, _i, _p, _i64, _p]),
    "qvit_codes_to_bf16_t": (_i, [_p, _i64, _i64, _i64, _p, _i64, _p]),
    "qvit_gemm_bf16_split": (_i, [_p, _i64, _i, _p, _i64, _i, _i, _i, _p, _i64, C.POINTER(Epilogue), _p]),
    "qvit_layernorm_quantize": (_i, [_p, _i64, _i, _p, _p, _f, _p, _p, _p, _p, _i64, _p, _p, _p]),
    "qvit_layernorm_fwd": (_i, [_p, _i64, _i, _p, _p, _f, _p, _p, _p, _p]),
    "qvit_layernorm_bwd": (_i, [_p, _p, _i64, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "qvit_attention_train_fwd": (_i, [_p, _i, _i, _i, _i, _f, _p, _p, _p, _p]),
    "qvit_attention_train_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _p, _p, _p]),
    "qvit_attention_train_bwd_prof": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _p, _p, _p, _p]),
    "qvit_gemm_bf16_split_t": (_i, [_p, _i64, _i, _i64, _p, _i64, _i64, _i, _i, _p, _i64, _p, _p, _p]),
    "qvit_codes_to_bf16": (_i, [_p, _i64, _i64, _i64, _p, _i64, _p]),
    "qvit_grad_prep": (_i, [_p, _i64, _i64, _i64, _p, _i64, _p, _i64, _p, _p, _p]),
    "qvit_attention_f32": (_i, [_p, _i, _i, _i, _i, _f, _p, _p]),
    "qvit_attention_quantize_sym": (_i, [_p, _i, _i, _i, _i, _f, _p, _p, _p, _p, _i64, _p, _p, _p]),
    "qvit_attention_f32_debug": (_i, [_p, _i, _i, _i, _i, _f, _p, _p, _i, _p]),
    "qvit_attention_f16x2": (_i, [_p, _i64, _i, _i, _i, _i, _i, _f, _i, _i, _i, _p, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "qvit_split2_f16": (_i, [_p, _i64, _i, _i64, _p, _p, _i64, _i, _p, _p]),
    "qvit_quantize_sym_bf16": (_i, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _p, _p]),
}

_lock = threading.Lock()
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib() -> C.CDLL:
    """Load (once) and return the library with typed prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m quantized_vit_b200.build` "
                "(nvcc, sm_100a).  quantized_vit_b200 has no CPU or PyTorch fallback for the hot path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


CALLS = 0     # number of successful C-ABI compute calls (each enqueues >= 1 of our kernels); read by bench.py


def check(status: int, what: str = "") -> None:
    global CALLS
    CALLS += 1
    if status != 0:
        msg = lib().qvit_last_error()
        raise RuntimeError(f"libqvit_b200 {what} failed (status {status}): {msg.decode() if msg else ''}")


def ptr(t) -> int | None:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("quantized_vit_b200: the hot path runs on CUDA (sm_100a) only; got a CPU tensor. "
                               "There is no CPU fallback - move the module and its inputs to the GPU.")
