"""Oracle (TEST INFRASTRUCTURE): whole-model CPU forwards over a plain ``state_dict``.

Functional restatements of the two callers of the hot path, so that model-level parity (logits,
top-1) can be checked without the reference's module classes being present on the GPU box:

  * VisionTransformer.forward        VIT:290-328  (PatchEmbed VIT:94-103, ViTAttention VIT:125-153,
                                                  Mlp VIT:170-177, Block VIT:202-208)
    with every nn.Linear / nn.Conv2d replaced by the GETA quantized layer (QM:65-79, QL:495-499,
    QL:575-587)
  * UltraNetQua.layers               MM:71-125    (conv / nn.BatchNorm2d eval / act-quant / max-pool)

plus ``fill_state_dict_`` - a construction-order-independent, seeded parameter fill used by both
``oracle/make_golden.py`` (which feeds the UNMODIFIED reference) and the tests (which feed the
product), so large models never need their weights committed.
"""
from __future__ import annotations

import hashlib
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

from . import ref_geta, ref_ultra


# ------------------------------------------------------------------------------------------
# deterministic parameter fill
# ------------------------------------------------------------------------------------------

def _seed_for(name: str, seed: int) -> int:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def fill_state_dict_(sd: Dict[str, torch.Tensor], seed: int = 0, weight_std: float = 0.02,
                     skip_substrings=("d_quant", "q_m", "t_quant", "num_batches_tracked")) -> Dict[str, torch.Tensor]:
    """In-place, per-tensor-seeded fill (depends only on the tensor's NAME and shape):
    *.weight of rank>=2 ~ N(0, weight_std^2); LayerNorm/BatchNorm weight ~ U(0.5,1.5); biases ~ N(0, 0.02^2)
    (N(0,0.1^2)+0.5 for BatchNorm); running_mean ~ N(0,0.1^2); running_var ~ U(0.5,1.5);
    cls_token/pos_embed ~ N(0,0.02^2).  Quantizer parameters are left alone."""
    for name, ten in sd.items():
        if any(s in name for s in skip_substrings) or not ten.is_floating_point():
            continue
        g = torch.Generator().manual_seed(_seed_for(name, seed))
        if name.endswith("running_var"):
            v = torch.rand(ten.shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            v = torch.randn(ten.shape, generator=g) * 0.1
        elif name.endswith("weight") and ten.dim() >= 2:
            v = torch.randn(ten.shape, generator=g) * weight_std
        elif name.endswith("weight"):
            v = torch.rand(ten.shape, generator=g) + 0.5
        elif name.endswith("bias"):
            v = torch.randn(ten.shape, generator=g) * 0.02
        else:
            v = torch.randn(ten.shape, generator=g) * 0.02
        ten.copy_(v.to(ten.dtype))
    return sd


# ------------------------------------------------------------------------------------------
# ViT with GETA quantized layers
# ------------------------------------------------------------------------------------------

def _qparams(sd, prefix: str, which: str) -> Optional[Dict[str, torch.Tensor]]:
    key = f"{prefix}.d_quant_{which}"
    if key not in sd:
        return None
    q = {"d": sd[key], "q_m": sd[f"{prefix}.q_m_{which}"]}
    if f"{prefix}.t_quant_{which}" in sd:
        q["t"] = sd[f"{prefix}.t_quant_{which}"]
    return q


def _qlinear(sd, prefix: str, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    r = ref_geta.quantize_linear_forward(x, sd[f"{prefix}.weight"], sd.get(f"{prefix}.bias"),
                                         _qparams(sd, prefix, "wt"), _qparams(sd, prefix, "act"),
                                         want_int=False)
    if taps is not None and taps.get("__layers__"):
        taps[f"{prefix}.in"], taps[f"{prefix}.y"] = x, r["y"]
    return r["y"]


def vit_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, depth: int, num_heads: int,
                patch: int = 16, ln_eps: float = 1e-6, taps: Optional[dict] = None) -> torch.Tensor:
    """Logits of the GETA-quantized VisionTransformer (no dist token, no pre_logits).

    VIT:290-328; LayerNorm eps 1e-6 (VIT:241); attention scale head_dim^-0.5 (VIT:119); GELU(erf)."""
    sd = {k: v.detach().float().cpu() for k, v in sd.items()}
    x = x.detach().float().cpu()
    r = ref_geta.quantize_conv2d_forward(x, sd["patch_embed.proj.weight"], sd.get("patch_embed.proj.bias"),
                                         _qparams(sd, "patch_embed.proj", "wt"), _qparams(sd, "patch_embed.proj", "act"),
                                         stride=patch, padding=0, want_int=False)
    h = r["y"].flatten(2).transpose(1, 2)                                   # VIT:100
    B = h.shape[0]
    h = torch.cat((sd["cls_token"].expand(B, -1, -1), h), dim=1) + sd["pos_embed"]   # VIT:295-305
    D = h.shape[-1]
    if taps is not None:
        taps["embed"] = h
    for i in range(depth):
        p = f"blocks.{i}"
        y = F.layer_norm(h, (D,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], ln_eps)
        qkv = _qlinear(sd, f"{p}.attn.qkv", y, taps)                              # VIT:133
        N = qkv.shape[1]
        qkv = qkv.reshape(B, N, 3, num_heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) * (q.shape[-1] ** -0.5)            # VIT:141
        attn = attn.softmax(dim=-1)
        y = (attn @ v).transpose(1, 2).reshape(B, N, -1)                    # VIT:149
        if taps is not None:
            taps[f"{p}.attn.proj.in"] = y
        h = h + _qlinear(sd, f"{p}.attn.proj", y, taps)                           # VIT:151, 206
        y = F.layer_norm(h, (D,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], ln_eps)
        y = F.gelu(_qlinear(sd, f"{p}.mlp.fc1", y, taps))                         # VIT:172-173
        h = h + _qlinear(sd, f"{p}.mlp.fc2", y, taps)                             # VIT:175, 207
        if taps is not None:
            taps[f"{p}.out"] = h
    h = F.layer_norm(h, (D,), sd["norm.weight"], sd["norm.bias"], ln_eps)   # VIT:309
    return _qlinear(sd, "head", h[:, 0], taps)                                    # VIT:312, 327


def vit_forward_autograd(params: Dict[str, torch.Tensor], x: torch.Tensor, depth: int, num_heads: int,
                         patch: int = 16, ln_eps: float = 1e-6, taps: Optional[dict] = None) -> torch.Tensor:
    """vit_forward under autograd (QAT parity, config 3): ``params`` are fp32 CPU leaf tensors (requires_grad as the
    caller set it); the quantizers differentiate through ref_geta.SymQuantFn, i.e. the reference's autograd.Functions
    (QL:71-125, 163-205).  ``taps`` receives every quantized layer's input and output with retain_grad(), so that after
    backward() ``taps[name + ".in"].grad`` / ``taps[name + ".y"].grad`` are the reference's per-layer gradients."""
    def q(prefix, which):
        return _qparams(params, prefix, which)

    def qlin(prefix, y):
        if taps is not None and y.requires_grad:
            y.retain_grad()
        out = ref_geta.quantize_linear_autograd(y, params[f"{prefix}.weight"], params.get(f"{prefix}.bias"), q(prefix, "wt"),
                                                q(prefix, "act"))
        if taps is not None:
            out.retain_grad()
            taps[f"{prefix}.in"], taps[f"{prefix}.y"] = y, out
        return out

    h = ref_geta.quantize_conv2d_autograd(x, params["patch_embed.proj.weight"], params.get("patch_embed.proj.bias"),
                                          q("patch_embed.proj", "wt"), q("patch_embed.proj", "act"), stride=patch, padding=0)
    if taps is not None:
        h.retain_grad()
        taps["patch_embed.proj.in"], taps["patch_embed.proj.y"] = x, h
    h = h.flatten(2).transpose(1, 2)
    B = h.shape[0]
    h = torch.cat((params["cls_token"].expand(B, -1, -1), h), dim=1) + params["pos_embed"]
    D = h.shape[-1]
    for i in range(depth):
        p = f"blocks.{i}"
        y = F.layer_norm(h, (D,), params[f"{p}.norm1.weight"], params[f"{p}.norm1.bias"], ln_eps)
        qkv = qlin(f"{p}.attn.qkv", y)
        N = qkv.shape[1]
        qkv = qkv.reshape(B, N, 3, num_heads, -1).permute(2, 0, 3, 1, 4)
        attn = ((qkv[0] @ qkv[1].transpose(-2, -1)) * (qkv.shape[-1] ** -0.5)).softmax(dim=-1)
        y = (attn @ qkv[2]).transpose(1, 2).reshape(B, N, -1)
        h = h + qlin(f"{p}.attn.proj", y)
        y = F.layer_norm(h, (D,), params[f"{p}.norm2.weight"], params[f"{p}.norm2.bias"], ln_eps)
        y = F.gelu(qlin(f"{p}.mlp.fc1", y))
        h = h + qlin(f"{p}.mlp.fc2", y)
    h = F.layer_norm(h, (D,), params["norm.weight"], params["norm.bias"], ln_eps)
    return qlin("head", h[:, 0])


def grad_digest(g: torch.Tensor, n_samples: int = 256):
    """(stats[3] = sum, abs-sum, l2 in float64; samples at a fixed pseudo-random stride) of a gradient tensor - the
    committed fingerprint of tensors too large to commit (oracle/make_golden.py::golden_qat and its tests)."""
    flat = g.detach().double().reshape(-1).cpu()
    n = flat.numel()
    idx = (torch.arange(min(n_samples, n), dtype=torch.int64) * 7919) % n
    stats = torch.stack([flat.sum(), flat.abs().sum(), flat.pow(2).sum().sqrt()])
    return stats.numpy(), flat[idx].float().numpy()


# ------------------------------------------------------------------------------------------
# UltraNet
# ------------------------------------------------------------------------------------------

# (conv index in nn.Sequential, bn index or None, maxpool after?) - MM:71-125
ULTRANET_LAYERS = [(0, 1, True), (4, 5, True), (8, 9, True), (12, 13, True),
                   (16, 17, False), (19, 20, False), (22, 23, False), (25, 26, False), (28, None, False)]


def ultranet_features(sd: Dict[str, torch.Tensor], x: torch.Tensor, w_bit: int = 4, a_bit: int = 4,
                      bn_eps: float = 1e-5, taps: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """UltraNetQua.layers(x) in eval mode (MM:71-125 / MM:134): the [B,36,H/16,W/16] map fed to the
    YOLO head (the head itself, MM:32-60, is fp32 glue and out of scope)."""
    sd = {k: v.detach().float().cpu() for k, v in sd.items() if v.is_floating_point()}
    h = x.detach().float().cpu()
    for conv_i, bn_i, pool in ULTRANET_LAYERS:
        w = sd[f"layers.{conv_i}.weight"]
        pad = 1 if w.shape[-1] == 3 else 0
        h = ref_ultra.conv2d_q_forward(h, w, sd.get(f"layers.{conv_i}.bias"), w_bit, stride=1, padding=pad)
        if bn_i is not None:
            h = F.batch_norm(h, sd[f"layers.{bn_i}.running_mean"], sd[f"layers.{bn_i}.running_var"],
                             sd[f"layers.{bn_i}.weight"], sd[f"layers.{bn_i}.bias"], False, 0.0, bn_eps)
            h = ref_ultra.ultra_act_values(h, a_bit)
            if pool:
                h = F.max_pool2d(h, 2, 2)
        if taps is not None:
            taps.append(h)
    return h


def yolo_decode(p: torch.Tensor, img_size, anchors=((20, 20),) * 6):
    """YOLOLayer.forward in eval mode (MM:32-60) - fp32 glue, restated only so a whole-model golden
    can be compared; returns the `io` tensor [B, na*ny*nx, 6]."""
    bs, _, ny, nx = p.shape
    na, no = len(anchors), 6
    stride = max(img_size) / max(nx, ny)
    yv, xv = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing="ij")
    grid = torch.stack((xv, yv), 2).float().view(1, 1, ny, nx, 2)
    anchor_wh = (torch.tensor(anchors, dtype=torch.float32) / stride).view(1, na, 1, 1, 2)
    p = p.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
    io = p.clone()
    io[..., :2] = torch.sigmoid(io[..., :2]) + grid
    io[..., 2:4] = torch.exp(io[..., 2:4]) * anchor_wh
    io[..., :4] *= stride
    torch.sigmoid_(io[..., 4:])
    return io.view(bs, -1, no)
