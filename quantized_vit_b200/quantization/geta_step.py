"""Fused update of every quantizer scalar of a model - the quant-parameter half of ``GETA.step()``
(``QViT_with_GETA/only_train_once/optimizer/geta.py:571-772, 787-804``, ``base_optimizer.py:17-86``; SURVEY.md
section 8f rank 2).

The reference walks every ``d_quant_* / q_m_* / t_quant_*`` parameter (six (1,) tensors per layer, ~300 for ViT-B) in
nested Python loops with substring matching, runs a handful of ATen kernels per scalar and calls ``.item()`` for every
projection bound, every step.  ``GetaQuantParamStepper`` keeps a device table of pointers to those parameters and
launches ONE kernel (``qvit_geta_quant_step``) per step; nothing synchronises with the host.

It covers the three stages GETA applies to quantizer scalars of groups without active pruning:
``"descent"`` (geta.py:571-596), ``"range"`` (598-665 followed by 667-721) and ``"fix"`` (723-772).  Model weights,
pruning and the stage schedule stay with the caller's optimizer: this is the consumer of the step-size gradients the
fused backward produces, not a re-implementation of GETA."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Tuple

import torch

from .. import _lib

_SLOTS = ("d_quant_wt", "q_m_wt", "t_quant_wt", "d_quant_act", "q_m_act", "t_quant_act")
_VARIANTS = {"sgd": 0, "adam": 1, "adamw": 2}
_MODES = {"descent": 0, "range": 1, "fix": 2}


class GetaQuantParamStepper:
    def __init__(self, named_parameters: Iterable[Tuple[str, torch.nn.Parameter]], *, variant: str = "sgd", lr: float,
                 lr_quant: float = 1e-3, first_momentum: Optional[float] = None, second_momentum: Optional[float] = None,
                 dampening: Optional[float] = None, weight_decay: Optional[float] = None, min_bit_wt: float = 2,
                 max_bit_wt: float = 16, min_bit_act: float = 2, max_bit_act: float = 16,
                 grad_clip: Optional[Tuple[float, float]] = None, safe_guard: float = 1e-8):
        if variant not in _VARIANTS:
            raise ValueError(f"variant must be one of {sorted(_VARIANTS)}")
        self.variant, self.lr, self.lr_quant = variant, float(lr), float(lr_quant)
        self.first_momentum = 0.0 if first_momentum is None else float(first_momentum)
        self.second_momentum = 0.0 if second_momentum is None else float(second_momentum)
        self.dampening = 0.0 if dampening is None else float(dampening)
        self.weight_decay = weight_decay
        self.min_bit_wt, self.max_bit_wt = float(min_bit_wt), float(max_bit_wt)
        self.min_bit_act, self.max_bit_act = float(min_bit_act), float(max_bit_act)
        self.grad_clip, self.safe_guard = grad_clip, float(safe_guard)
        self.num_steps = 0
        layers: Dict[str, Dict[str, torch.nn.Parameter]] = {}
        for name, p in named_parameters:
            layer, _, leaf = name.rpartition(".")
            if leaf in _SLOTS:
                layers.setdefault(layer, {})[leaf] = p
        self.layer_names = list(layers)
        self._params = [[layers[n].get(s) for s in _SLOTS] for n in self.layer_names]
        flat = [p for row in self._params for p in row if p is not None]
        if not flat:
            raise ValueError("no quantizer parameters (d_quant_* / q_m_* / t_quant_*) found")
        self.device = flat[0].device
        for p in flat:
            _lib.require_cuda(p)
            if p.dtype != torch.float32 or p.numel() != 1 or p.device != self.device:
                raise ValueError("quantizer parameters must be (1,) fp32 tensors on one CUDA device")
        n = len(self.layer_names) * 6
        self._m1 = torch.zeros(n, dtype=torch.float32, device=self.device)
        self._m2 = torch.zeros(n, dtype=torch.float32, device=self.device)
        self._inited = torch.zeros(n, dtype=torch.uint8, device=self.device)
        self._ptab = torch.zeros(n, dtype=torch.int64, device=self.device)
        self._gtab = torch.zeros(n, dtype=torch.int64, device=self.device)
        # the gradient pointers change whenever autograd re-allocates .grad (zero_grad(set_to_none=True)): they are staged
        # through a RING of pinned tables, each guarded by an event recorded after its host->device copy, so that a host
        # running several steps ahead never overwrites a table whose DMA has not executed yet
        self._gtab_ring = [torch.zeros(n, dtype=torch.int64).pin_memory() for _ in range(4)]
        self._gtab_events = [None] * 4
        self._gtab_slot = 0
        self._gtab_key = None
        self._ptab_key = None
        self.flags = torch.zeros(1, dtype=torch.int32, device=self.device)

    def _refresh_tables(self):
        # parameter storage can be replaced (OTO pruning re-creates Parameters, .data may be reassigned): re-read the pointers
        ptrs = [0 if p is None else p.data_ptr() for row in self._params for p in row]
        key = tuple(ptrs)
        if key != self._ptab_key:
            self._ptab.copy_(torch.tensor(ptrs, dtype=torch.int64), non_blocking=False)
            self._ptab_key = key
        gptrs = tuple(0 if (p is None or p.grad is None) else p.grad.data_ptr() for row in self._params for p in row)
        if gptrs != self._gtab_key:                      # persistent .grad buffers (zero_grad(set_to_none=False)): nothing to copy
            j = self._gtab_slot
            if self._gtab_events[j] is not None:
                self._gtab_events[j].synchronize()       # four steps old: complete long ago unless the host is far ahead
            self._gtab_ring[j].copy_(torch.tensor(gptrs, dtype=torch.int64))
            self._gtab.copy_(self._gtab_ring[j], non_blocking=True)
            ev = self._gtab_events[j] = self._gtab_events[j] or torch.cuda.Event()
            ev.record()
            self._gtab_slot = (j + 1) % len(self._gtab_ring)
            self._gtab_key = gptrs

    def _poll_flags(self):
        """Deferred, sync-free NaN report: the flag word of step N is copied to pinned memory asynchronously and examined
        at step N+1 (or later) once its event has completed - GETA's scalars turning NaN raises NanInGradientError by
        default, one step late, without a host synchronisation in the training loop."""
        from .quant_layers import NanInGradientError
        st = self.__dict__.setdefault("_poll", {"host": torch.zeros(1, dtype=torch.int32).pin_memory(), "ev": None})
        if torch.cuda.is_current_stream_capturing():
            return
        if st["ev"] is not None and st["ev"].query():
            bits = int(st["host"][0])
            st["ev"] = None
            if bits & _lib.QVIT_FLAG_NAN_GRAD:
                raise NanInGradientError("Error: NaN appears in gradient! (quantizer-scalar gradients, reported by qvit_geta_quant_step)")
        if st["ev"] is None:
            st["host"].copy_(self.flags, non_blocking=True)     # the device word stays sticky: `flags` / explicit checks still see it
            st["ev"] = torch.cuda.Event()
            st["ev"].record()

    @torch.no_grad()
    def step(self, stage: str = "range", bit_dict: Optional[Dict[str, Dict[str, float]]] = None) -> None:
        """One optimizer step for all quantizer scalars.  stage: "descent" | "range" | "fix" (bit_dict as returned by
        GETA.get_bitwidth_dict: {layer: {"weight": b, "activation": b}})."""
        if stage not in _MODES:
            raise ValueError(f"stage must be one of {sorted(_MODES)}")
        self.num_steps += 1
        self._refresh_tables()
        is_adam = self.variant in ("adam", "adamw")
        bc1 = 1.0 - self.first_momentum ** self.num_steps if is_adam else 1.0
        bc2 = 1.0 - self.second_momentum ** self.num_steps if is_adam else 1.0
        fb_w = fb_a = None
        if stage == "fix":
            if bit_dict is None:
                raise ValueError('stage "fix" needs bit_dict')
            fb_w = torch.tensor([float(bit_dict[n]["weight"]) for n in self.layer_names], dtype=torch.float32).to(self.device)
            if all("activation" in bit_dict[n] for n in self.layer_names):
                fb_a = torch.tensor([float(bit_dict[n]["activation"]) for n in self.layer_names], dtype=torch.float32).to(self.device)
        clip = self.grad_clip is not None
        cmin, cmax = (self.grad_clip if clip else (0.0, 0.0))
        L = _lib.lib()
        _lib.check(L.qvit_geta_quant_step(
            _lib.ptr(self._ptab), _lib.ptr(self._gtab), _lib.ptr(self._m1), _lib.ptr(self._m2), _lib.ptr(self._inited),
            _lib.ptr(fb_w), _lib.ptr(fb_a), len(self.layer_names), _VARIANTS[self.variant], _MODES[stage], self.lr, self.lr_quant,
            0 if self.weight_decay is None else 1, 0.0 if self.weight_decay is None else float(self.weight_decay),
            self.first_momentum, self.second_momentum, self.dampening if not is_adam else self.first_momentum,
            C.c_double(bc1), C.c_double(bc2), self.safe_guard, 1 if clip else 0, float(cmin), float(cmax), self.min_bit_wt,
            self.max_bit_wt, self.min_bit_act, self.max_bit_act, _lib.ptr(self.flags), _lib.stream()), "qvit_geta_quant_step")
        self._poll_flags()
