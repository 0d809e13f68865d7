"""Module swapper and quant-parameter helpers - mirror of
``QViT_with_GETA/only_train_once/quantization/quant_model.py`` (QM), targeting the sm_100a layer classes."""
from __future__ import annotations

import logging
import math
from typing import Dict, Union

import torch.nn as nn

from .quant_layers import LAYER_TO_QUANTLAYER, QuantizationMode, QuantizationType


def _parse_enum(enum_cls, value, what: str):
    if not isinstance(value, str):
        return value
    try:
        return enum_cls(value)
    except ValueError:
        raise ValueError(f"Invalid quantization {what}: {value}. Must be one of {[m.value for m in enum_cls]}")


def model_to_quantize_model(model: nn.Module, d_quant_init: float = 1e-4, t_quant_init: float = 1.0,
                            q_m_init: float = 1.0, quant_init_by_module: bool = True, num_bits: int = 16,
                            quant_type: Union[QuantizationType, str] = QuantizationType.SYMMETRIC_NONLINEAR,
                            quant_mode: Union[QuantizationMode, str] = QuantizationMode.WEIGHT_ONLY) -> nn.Module:
    """QM:15-82: replace every module whose class is named exactly ``Linear`` / ``Conv2d`` (QM:66) by the quantized
    class built with ``from_module``.  Enum arguments may be given as strings; unknown strings raise ValueError
    (QM:48-62).  Returns the same (mutated) model."""
    quant_type = _parse_enum(QuantizationType, quant_type, "type")
    quant_mode = _parse_enum(QuantizationMode, quant_mode, "mode")
    targets = [(name, mod) for name, mod in model.named_modules() if type(mod).__name__ in LAYER_TO_QUANTLAYER]
    for name, mod in targets:
        parent_name, _, leaf = name.rpartition(".")
        parent = model.get_submodule(parent_name)
        qmod = LAYER_TO_QUANTLAYER[type(mod).__name__].from_module(
            module=mod, d_quant_init=d_quant_init, t_quant_init=t_quant_init, q_m_init=q_m_init, quant_type=quant_type,
            quant_mode=quant_mode, quant_init_by_module=quant_init_by_module, num_bits=num_bits)
        setattr(parent, leaf, qmod)
    logging.getLogger(__name__).info(f"Converted {len(targets)} layers to quantized versions")
    return model


def get_quant_param_dict(model: nn.Module) -> Dict[str, Dict[str, float]]:
    """QM:85-101: {layer name: {d_quant_wt: .., q_m_wt: .., ...}}.  One batched device->host read instead of an
    ``.item()`` per parameter."""
    import torch
    names, tensors = [], []
    for name, param in model.named_parameters():
        if any(tag in name for tag in ("d_quant", "t_quant", "q_m")):
            names.append(name)
            tensors.append(param.detach().reshape(-1)[:1].float())
    out: Dict[str, Dict[str, float]] = {}
    if not names:
        return out
    by_dev = {}
    for i, t in enumerate(tensors):
        by_dev.setdefault(str(t.device), []).append(i)
    values = [0.0] * len(names)
    for idxs in by_dev.values():
        vals = torch.cat([tensors[i] for i in idxs]).tolist()
        for i, v in zip(idxs, vals):
            values[i] = v
    for name, v in zip(names, values):
        layer, _, pname = name.rpartition(".")
        out.setdefault(layer, {})[pname] = v
    return out


def get_bitwidth_dict(param_dict: Dict[str, Dict[str, float]]) -> Dict[str, Dict[str, float]]:
    """QM:104-136: b = log2(|q_m|^t / |d| + 1) + 1 per layer for weights and (if present) activations."""

    def bits(d: float, q_m: float, t: float = 1.0) -> float:
        return math.log2(math.exp(t * math.log(abs(q_m))) / abs(d) + 1) + 1

    out: Dict[str, Dict[str, float]] = {}
    for layer, p in param_dict.items():
        out[layer] = {"weight": bits(p["d_quant_wt"], abs(p["q_m_wt"]), p.get("t_quant_wt", 1.0))}
        if "d_quant_act" in p:
            out[layer]["activation"] = bits(p["d_quant_act"], abs(p["q_m_act"]), p.get("t_quant_act", 1.0))
    return out
