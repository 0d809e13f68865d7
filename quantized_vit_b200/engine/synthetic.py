"""Synthetic random-init parameters for benchmarks and smoke tests (no dataset / checkpoint is reachable offline).

``vit_state_dict`` reproduces what the reference pipeline would hand us for a freshly constructed model:
``VisionTransformer`` init (vit_model.py:330-346: trunc_normal(std=.02) linears, LayerNorm (1, 0), kaiming conv)
followed by ``model_to_quantize_model(num_bits, "symmetric+linear", "weight_and_activation")`` whose
``initialize_quant_layer`` (quant_layers.py:413-440) sets q_m = max|W|, d = q_m / (2^(b-1)-1) for the weight AND the
activation quantizer.  ``act_bits`` overrides d_quant_act = q_m_act / (2^(act_bits-1)-1) (the W4A8 configuration)
and ``calibrate_to`` replaces q_m_act by a fixed range (calibrated-range fixture)."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .. import ops


def vit_state_dict(embed_dim=768, depth=12, num_heads=12, patch=16, img=224, classes=1000, mlp_ratio=4, num_bits=4,
                   act_bits: Optional[int] = None, calibrate_to: Optional[float] = None, seed=0,
                   device="cuda") -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    D, hid = embed_dim, int(embed_dim * mlp_ratio)
    n_tok = (img // patch) ** 2 + 1
    sd: Dict[str, torch.Tensor] = {}

    def tn(*shape, std=0.02):
        t = torch.empty(*shape)
        torch.nn.init.trunc_normal_(t, std=std, generator=g)
        return t

    sd["cls_token"] = tn(1, 1, D)
    sd["pos_embed"] = tn(1, n_tok, D)
    fan_out = patch * patch * D
    sd["patch_embed.proj.weight"] = torch.randn(D, 3, patch, patch, generator=g) * math.sqrt(2.0 / fan_out)
    sd["patch_embed.proj.bias"] = torch.zeros(D)
    for i in range(depth):
        p = f"blocks.{i}"
        sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"] = torch.ones(D), torch.zeros(D)
        sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"] = torch.ones(D), torch.zeros(D)
        sd[f"{p}.attn.qkv.weight"], sd[f"{p}.attn.qkv.bias"] = tn(3 * D, D), torch.zeros(3 * D)
        sd[f"{p}.attn.proj.weight"], sd[f"{p}.attn.proj.bias"] = tn(D, D), torch.zeros(D)
        sd[f"{p}.mlp.fc1.weight"], sd[f"{p}.mlp.fc1.bias"] = tn(hid, D), torch.zeros(hid)
        sd[f"{p}.mlp.fc2.weight"], sd[f"{p}.mlp.fc2.bias"] = tn(D, hid), torch.zeros(D)
    sd["norm.weight"], sd["norm.bias"] = torch.ones(D), torch.zeros(D)
    sd["head.weight"], sd["head.bias"] = tn(classes, D), torch.zeros(classes)
    sd = {k: v.to(device) for k, v in sd.items()}
    levels_w = 2 ** (num_bits - 1) - 1
    levels_a = 2 ** ((act_bits or num_bits) - 1) - 1
    for name in [k[:-7] for k in list(sd) if k.endswith(".weight") and sd[k].dim() >= 2]:
        qm = ops.absmax(sd[name + ".weight"])                   # initialize_quant_layer on the device (no host sync)
        sd[name + ".d_quant_wt"], sd[name + ".q_m_wt"] = qm / levels_w, qm.clone()
        qa = qm.clone() if calibrate_to is None else torch.full_like(qm, float(calibrate_to))
        sd[name + ".d_quant_act"], sd[name + ".q_m_act"] = qa / levels_a, qa
    return sd


def ultranet_state_dict(seed=0, device="cuda") -> Dict[str, torch.Tensor]:
    """UltraNetQua (mymodel.py:71-125) with default PyTorch init and non-trivial eval BatchNorm statistics."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    convs = [(0, 3, 16, 3), (4, 16, 32, 3), (8, 32, 64, 3), (12, 64, 64, 3), (16, 64, 64, 3), (19, 64, 64, 3),
             (22, 64, 64, 3), (25, 64, 64, 3), (28, 64, 36, 1)]
    sd: Dict[str, torch.Tensor] = {}
    for idx, cin, cout, k in convs:
        bound = 1.0 / math.sqrt(cin * k * k)
        sd[f"layers.{idx}.weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        if idx == 28:
            sd[f"layers.{idx}.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound
        else:
            b = idx + 1
            sd[f"layers.{b}.weight"] = torch.rand(cout, generator=g) * 0.2 + 0.2
            sd[f"layers.{b}.bias"] = torch.rand(cout, generator=g) * 0.4 + 0.3
            sd[f"layers.{b}.running_mean"] = torch.randn(cout, generator=g) * 0.1
            sd[f"layers.{b}.running_var"] = torch.rand(cout, generator=g) + 0.5
    return {k: v.to(device) for k, v in sd.items()}
