"""CPU restatement of the quantizer-scalar half of GETA.step() - TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows QViT_with_GETA/only_train_once/optimizer/geta.py and base_optimizer.py with torch CPU ops on (1,) fp32 tensors,
operation for operation, so that the values round exactly as the reference's do:
  grad clamp                geta.py:160-165
  compute_grad_variant      base_optimizer.py:17-86   (first buffer = grad, bias corrections 1 - beta^t, safe_guard 1e-8)
  gradient_descent_step     geta.py:571-596
  ..._range_wt / _range_act geta.py:598-665 / 667-721 (incl. the `else` branch of range_wt that also moves the
                                                       activation scalars with the model learning rate)
  ..._fix                   geta.py:723-772
  _d_quant_helper           geta.py:787-804
Pinned against the real GETA class by tests/golden/geta_step.npz (oracle/make_golden.py::golden_geta_step)."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

QUANT_TAGS = ("d_quant", "t_quant", "q_m")


def d_quant_helper(bit_width, q_m, t_quant):
    """geta.py:787-804."""
    if t_quant is None:
        t_quant = 1.0
    q_m = torch.max(torch.abs(q_m)).item() if isinstance(q_m, torch.Tensor) else abs(q_m)
    q_m = max(abs(q_m), 1e-10)
    return math.exp(t_quant * math.log(abs(q_m))) / (2 ** (bit_width - 1) - 1)


class GetaQuantStepRef:
    """State (moment buffers, step counter) + one step over a dict {param name: (1,) tensor with .grad}."""

    def __init__(self, variant="sgd", lr=0.1, lr_quant=1e-3, first_momentum=None, second_momentum=None, dampening=None,
                 weight_decay=None, min_bit_wt=2, max_bit_wt=16, min_bit_act=2, max_bit_act=16, grad_clip=None):
        self.hp = dict(variant=variant, lr=lr, lr_quant=lr_quant, first_momentum=first_momentum or 0.0,
                       second_momentum=second_momentum or 0.0, dampening=dampening or 0.0, weight_decay=weight_decay)
        self.min_bit_wt, self.max_bit_wt, self.min_bit_act, self.max_bit_act = min_bit_wt, max_bit_wt, min_bit_act, max_bit_act
        self.grad_clip = grad_clip
        self.num_steps, self.safe_guard = 0, 1e-8
        self.m1: Dict[str, torch.Tensor] = {}
        self.m2: Dict[str, torch.Tensor] = {}

    def _grad_variant(self, params):
        hp = self.hp
        is_adam = hp["variant"] in ("adam", "adamw")
        bc1 = 1.0 - hp["first_momentum"] ** self.num_steps if is_adam else None
        bc2 = 1.0 - hp["second_momentum"] ** self.num_steps if is_adam else None
        gv = {}
        for name, p in params.items():
            if p.grad is None:
                continue
            g = torch.clone(p.grad.data).detach()
            if hp["weight_decay"] is not None and hp["variant"] != "adamw":
                g += hp["weight_decay"] * p.data
            if not is_adam:
                if hp["first_momentum"] > 0.0 or hp["dampening"] > 0.0:
                    if hp["first_momentum"] > 0:
                        if name not in self.m1:
                            self.m1[name] = g
                        else:
                            self.m1[name].mul_(hp["first_momentum"]).add_(g, alpha=(1.0 - hp["dampening"]))
                        g = self.m1[name]
                gv[name] = g
            else:
                if hp["first_momentum"] > 0:
                    if name not in self.m1:
                        self.m1[name] = g
                    else:
                        self.m1[name].mul_(hp["first_momentum"]).add_(g, alpha=(1.0 - hp["first_momentum"]))
                    f = self.m1[name]
                else:
                    f = g
                if hp["second_momentum"] > 0:
                    if name not in self.m2:
                        self.m2[name] = g * g
                    else:
                        self.m2[name].mul_(hp["second_momentum"]).add_(g * g, alpha=(1.0 - hp["second_momentum"]))
                    v = self.m2[name]
                else:
                    v = g * g
                denom = (v / bc2).sqrt().add_(self.safe_guard)
                gv[name] = (f / bc1) / denom
        return gv

    def step(self, params: Dict[str, torch.Tensor], stage: str, bit_dict: Optional[dict] = None):
        hp = self.hp
        if self.grad_clip is not None:
            for p in params.values():
                if p.grad is not None:
                    p.grad = p.grad.clamp(min=self.grad_clip[0], max=self.grad_clip[1])
        self.num_steps += 1
        gv = self._grad_variant(params)

        def descend(names_pred, lr_key):
            for name, p in params.items():
                if name not in gv or not names_pred(name):
                    continue
                if hp["weight_decay"] is not None and hp["variant"] == "adamw":
                    p.data.add_(hp["weight_decay"] * p.data, alpha=-hp[lr_key])
                p.data.add_(gv[name], alpha=-hp[lr_key])

        is_quant = lambda n: any(t in n for t in QUANT_TAGS)                       # noqa: E731
        is_wt = lambda n: any(t in n for t in ("d_quant_wt", "t_quant_wt", "q_m_wt"))   # noqa: E731
        is_act = lambda n: any(t in n for t in ("d_quant_act", "t_quant_act", "q_m_act"))   # noqa: E731
        layers = []
        for n in params:
            layer = ".".join(n.split(".")[:-1])
            if layer not in layers:
                layers.append(layer)

        def get(layer, leaf):
            return params.get(f"{layer}.{leaf}")

        if stage == "descent":
            descend(is_quant, "lr_quant")
        elif stage == "range":
            descend(is_wt, "lr_quant")
            descend(lambda n: not is_wt(n), "lr")          # range_wt's `else` branch: everything else with the model lr
            for layer in layers:                          # projection (weights)
                d, qm, t = get(layer, "d_quant_wt"), get(layer, "q_m_wt"), get(layer, "t_quant_wt")
                if d is None:
                    continue
                lo = d_quant_helper(self.max_bit_wt, qm.data, None if t is None else t.data)
                hi = d_quant_helper(self.min_bit_wt, qm.data, None if t is None else t.data)
                d.data.clamp_(min=float(lo), max=float(hi))
            descend(is_act, "lr_quant")
            for layer in layers:                          # projection (activations)
                d, qm, t = get(layer, "d_quant_act"), get(layer, "q_m_act"), get(layer, "t_quant_act")
                if d is None:
                    continue
                lo = d_quant_helper(self.max_bit_act, qm.data, None if t is None else t.data)
                hi = d_quant_helper(self.min_bit_act, qm.data, None if t is None else t.data)
                d.data = torch.clip(d.data, min=lo, max=hi)
        elif stage == "fix":
            descend(is_quant, "lr_quant")
            for layer in bit_dict:
                for side, key in (("wt", "weight"), ("act", "activation")):
                    d, qm, t = get(layer, f"d_quant_{side}"), get(layer, f"q_m_{side}"), get(layer, f"t_quant_{side}")
                    if d is None:
                        continue
                    v = d_quant_helper(bit_dict[layer][key], qm.data, None if t is None else t.data)
                    d.data = torch.clip(d.data, min=v, max=v)
        else:
            raise ValueError(stage)
