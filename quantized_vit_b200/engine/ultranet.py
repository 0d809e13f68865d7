"""Fused integer inference engine for UltraNet W4A4 (``4-bit quantization/mymodel.py`` UltraNetQua.layers,
MM:71-125): 8 x [Conv2d_Q 3x3 -> nn.BatchNorm2d(eval) -> activation_quantize_fn(4) (-> MaxPool 2x2)] + 1x1 head conv.

One-time preparation (K1' + K5): weight codes via the fused tanh/max/round kernels, re-laid out as
[O, kh, kw, C] int8; BatchNorm folded to per-channel (scale, bias) with the nn.BatchNorm2d formula
(eps inside the sqrt; ``fold="export"`` selects quantization.py:34-46 instead).
Per image: every layer is ONE kernel (integer conv on activation codes + BN + clamp/round + pool on codes) - layers 1..8
as implicit GEMMs on the tcgen05 int8 tensor-core pipe (weights resident in shared memory, im2col tile gathered in shared
memory, int32 accumulators in TMEM), layer 0 (3 input channels) on the CUDA-core dp4a kernel -,
NHWC uint8 codes between layers, and the whole chain is replayed from a CUDA graph (batch-1 latency is
launch-bound: 0.4 GOP, 105 KB of weights).

First layer: the reference feeds the raw fp32 image to Conv2d_Q (QU:85-89).  ``input_bits=8`` (default) snaps the
image to the 8-bit grid the deployment flow uses (ultranet_param_gen.py:15) and runs layer 0 on the integer
path too; ``input_bits=None`` keeps the fp32-input semantics (fp32 conv kernel + BN/act/pool kernel).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .. import ops

# (conv index, bn index | None, pool) in nn.Sequential order - MM:71-125
ULTRANET_LAYERS = [(0, 1, True), (4, 5, True), (8, 9, True), (12, 13, True), (16, 17, False), (19, 20, False),
                   (22, 23, False), (25, 26, False), (28, None, False)]


class UltraNetEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], w_bit: int = 4, a_bit: int = 4, bn_eps: float = 1e-5,
                 input_bits: Optional[int] = 8, fold: str = "torch", device="cuda", conv: str = "tc"):
        """conv: "tc" = layers with 16 / 32 / 64 / 128 input channels (L1..L8) as implicit GEMMs on the tcgen05 int8 pipe
        (qvit_ultra_conv_tc), "simt" = every layer on the CUDA-core dp4a kernel (layer 0 with its 3 input channels always)."""
        if conv not in ("tc", "simt"):
            raise ValueError("conv must be 'tc' or 'simt'")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("UltraNetEngine runs on CUDA (sm_100a) only - there is no CPU fallback")
        if not (2 <= w_bit <= 8 and 1 <= a_bit <= 8):
            raise ValueError("UltraNetEngine: 2 <= w_bit <= 8 and 1 <= a_bit <= 8")
        self.w_bit, self.a_bit, self.input_bits = w_bit, a_bit, input_bits
        self.w_levels = float(2 ** (w_bit - 1) - 1)
        self.a_levels = 2 ** a_bit - 1
        sd = {k: v.detach().to(self.device) for k, v in state_dict.items() if v.is_floating_point()}
        self.layers: List[dict] = []
        for conv_i, bn_i, pool in ULTRANET_LAYERS:
            w = sd[f"layers.{conv_i}.weight"].float()
            codes = ops.ultra_weight_codes(w, w_bit)                                  # [O, C, kh, kw] int8
            L = dict(codes_oihw=codes, codes_ohwi=codes.permute(0, 2, 3, 1).contiguous(), pad=1 if w.shape[-1] == 3 else 0,
                     pool=pool, O=w.shape[0], C=w.shape[1], conv_bias=sd.get(f"layers.{conv_i}.bias"))
            if bn_i is not None:
                L["scale"], L["bias"] = ops.bn_fold(sd[f"layers.{bn_i}.weight"], sd[f"layers.{bn_i}.bias"],
                                                    sd[f"layers.{bn_i}.running_mean"], sd[f"layers.{bn_i}.running_var"],
                                                    bn_eps, mode=0 if fold == "torch" else 1)
            else:
                L["scale"], L["bias"] = None, (None if L["conv_bias"] is None else L["conv_bias"].float().contiguous())
            kh, kw = w.shape[2], w.shape[3]
            L["kh"], L["kw"] = kh, kw
            L["w_tc"] = ops.pack_conv_weights_tc(L["codes_ohwi"]) if (conv == "tc" and ops.ultra_conv_tc_supported(L["C"], L["O"], kh, kw)) else None
            self.layers.append(L)
        if input_bits is not None:
            # layer-0 input channels padded 3 -> 4 so the dp4a kernel reads whole words
            L0 = self.layers[0]
            c_pad = (L0["C"] + 3) // 4 * 4
            w0 = torch.zeros((L0["O"], L0["codes_ohwi"].shape[1], L0["codes_ohwi"].shape[2], c_pad), dtype=torch.int8,
                             device=self.device)
            w0[..., :L0["C"]] = L0["codes_ohwi"]
            L0["codes_ohwi_pad"], L0["c_pad"] = w0, c_pad
        self._graphs = {}

    def weight_bytes(self) -> int:
        return sum(L["codes_ohwi"].numel() for L in self.layers)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, taps: Optional[list] = None) -> torch.Tensor:
        """x [B, 3, H, W] fp32 in [0, 1] -> the [B, 36, H/16, W/16] fp32 map fed to the YOLO head (MM:134).
        ``taps`` (optional list) receives the uint8 NHWC activation codes after each of the 8 quantized layers."""
        ops._lib.require_cuda(x)
        L0 = self.layers[0]
        if self.input_bits is not None:
            in_levels = 2 ** self.input_bits - 1
            codes = ops.ultra_bn_act_pool_nchw(x, None, None, in_levels, False, ldc=L0["c_pad"])      # image -> u8 NHWC
            h = ops.ultra_conv_bn_act(codes, L0["codes_ohwi_pad"], L0["pad"], 1.0 / (in_levels * self.w_levels), L0["scale"],
                                      L0["bias"], self.a_levels, L0["pool"])
        else:
            y = ops.conv2d_f32_wcodes(x, L0["codes_oihw"], self.w_levels, None, (1, 1), (L0["pad"],) * 2, (1, 1))
            h = ops.ultra_bn_act_pool_nchw(y, L0["scale"], L0["bias"], self.a_levels, L0["pool"])
        if taps is not None:
            taps.append(h)
        acc_scale = 1.0 / (self.a_levels * self.w_levels)
        for L in self.layers[1:-1]:
            h = self._layer(L, h, acc_scale)
            if taps is not None:
                taps.append(h)
        return self._layer(self.layers[-1], h, acc_scale, f32_out=True)

    def _layer(self, L: dict, h: torch.Tensor, acc_scale: float, f32_out: bool = False) -> torch.Tensor:
        scale = None if f32_out else L["scale"]
        pool = False if f32_out else L["pool"]
        if L["w_tc"] is not None:
            return ops.ultra_conv_tc(h, L["w_tc"], L["O"], L["kh"], L["kw"], L["pad"], acc_scale, scale, L["bias"], self.a_levels, pool,
                                     f32_out=f32_out)
        return ops.ultra_conv_bn_act(h, L["codes_ohwi"], L["pad"], acc_scale, scale, L["bias"], self.a_levels, pool, f32_out=f32_out)

    __call__ = forward

    def capture(self, batch: int = 1, h: int = 160, w: int = 320):
        key = (batch, h, w)
        if key in self._graphs:
            return self._graphs[key]
        x = torch.zeros((batch, 3, h, w), dtype=torch.float32, device=self.device)
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self.forward(x)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = self.forward(x)
        self._graphs[key] = (x, y, g)
        return self._graphs[key]

    @staticmethod
    def macs_per_image(h: int = 160, w: int = 320) -> int:
        shapes = [(3, 16, 3, 1), (16, 32, 3, 2), (32, 64, 3, 4), (64, 64, 3, 8), (64, 64, 3, 16), (64, 64, 3, 16),
                  (64, 64, 3, 16), (64, 64, 3, 16), (64, 36, 1, 16)]
        return sum(ci * co * k * k * (h // s) * (w // s) for ci, co, k, s in shapes)
