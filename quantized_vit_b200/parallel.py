"""Data parallelism for the hot path - one process per GPU, torch.distributed plumbing.

* Inference shards by batch: every rank holds a full replica (43 MB of packed ViT-B weights) and a contiguous slice of
  the batch; there is NO collective on the forward path (``shard_batch`` / ``gather_outputs`` only).
* QAT exchanges exactly one thing per step: gradients.  ``GradientAllReducer`` averages (a) all weight/bias grads in
  flat fp32 buckets and (b) every layer's step-size / range / exponent gradients (d_quant_*, q_m_*, t_quant_*: <= 300
  scalars for ViT-B) packed into ONE small buffer, with NCCL over NVLink (gloo on CPU for tests).  It must run
  after ``loss.backward()`` and BEFORE ``optimizer.grad_clipping()`` (reference order utils.py:291-292) so that an
  N-GPU step equals the 1-GPU step on the concatenated batch.  The reference itself has no distributed code at all.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

QUANT_PARAM_TAGS = ("d_quant", "q_m", "t_quant")        # geta.py:253-272 finds them by the same substrings


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of n items for `rank` (first n % world ranks get one extra)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_outputs(y: torch.Tensor, total: int, group=None) -> Optional[torch.Tensor]:
    """Host-side gather of per-rank outputs (logits) on rank 0, in batch order.  Not on the timed forward path."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return y
    rank = dist.get_rank(group)
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn, *y.shape[1:]), dtype=y.dtype, device=y.device)
    pad[: y.shape[0]] = y
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


class GradientAllReducer:
    """Bucketed gradient averaging for QAT, overlapped with backward.  Build once per model.

    * Every bucket owns ONE persistent flat fp32 buffer and the parameters' ``.grad`` tensors are VIEWS into it
      (``zero_grad()`` below zeroes the buffers; a ``.grad`` that the training loop replaced or set to None is copied back
      in and re-pointed), so no flatten / unflatten copy runs per step.
    * ``register_post_accumulate_grad_hook`` counts the parameters of a bucket as autograd finishes them and launches the
      bucket's asynchronous all-reduce as soon as the last one is in - while the rest of backward is still running.
      Buckets are laid out in reverse parameter order (the order backward produces them).  All quantizer scalars
      (d_quant_* / q_m_* / t_quant_*) travel in one small extra bucket: "gradients and step sizes" (SURVEY.md 8e).
    * ``reduce()`` after ``backward()`` launches whatever has not been launched, waits, and applies the weight.  Weighting:
      every rank's loss is the mean over ITS shard; with ``local_batch`` / ``global_batch`` given each rank pre-scales by
      n_r / N and the collective sums, which is the gradient of the mean over the concatenated batch also for uneven
      shards; without them the shards are assumed equal and the collective averages.
    * Parameters that received no gradient on this rank stay ``None`` when ``keep_none`` (the autograd graph is the same on
      every data-parallel rank, so all ranks agree on which those are); their slice of the buffer travels as zeros.
    * Under CUDA-graph capture (forward + backward replayed from one graph) no collective is captured: the hook of a bucket's
      last parameter records an EXTERNAL event instead, and ``reduce()`` after ``graph.replay()`` makes a side stream wait for
      each bucket's event before launching its all-reduce there - the exchange still overlaps the rest of the replayed
      backward, and NCCL never runs inside a capture.
    Must run BEFORE gradient clipping (reference order utils.py:291-292)."""

    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], bucket_bytes: int = 64 << 20, group=None,
                 overlap: bool = True, keep_none: bool = True):
        self.group, self.keep_none = group, keep_none
        named = [(n, p) for n, p in named_params if p.requires_grad]
        quant = [p for n, p in named if any(tag in n for tag in QUANT_PARAM_TAGS)]
        rest = [p for n, p in named if not any(tag in n for tag in QUANT_PARAM_TAGS)]
        groups: List[List[torch.nn.Parameter]] = []
        cur, cur_bytes = [], 0
        for p in reversed(rest):                                   # backward reaches the last layers first
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > bucket_bytes:
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            groups.append(cur)
        self.quant = quant
        self.buckets = groups
        self._all = groups + ([quant] if quant else [])
        self._flat: List[torch.Tensor] = []
        self._views: List[List[torch.Tensor]] = []
        self._bucket_of = {}
        for bi, params in enumerate(self._all):
            dev = params[0].device
            flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
            views, off = [], 0
            for pi, p in enumerate(params):
                v = flat[off:off + p.numel()].view_as(p)
                views.append(v)
                off += p.numel()
                self._bucket_of[id(p)] = (bi, pi)
            self._flat.append(flat)
            self._views.append(views)
        self._ready = [0] * len(self._all)
        self._seen = [set() for _ in self._all]
        self._work = [None] * len(self._all)
        self._scale = None
        self._events = None          # per-bucket external events recorded by a captured backward
        self._graph_seen = None
        self._comm = None
        self._hooks = []
        if overlap and hasattr(torch.Tensor, "register_post_accumulate_grad_hook"):
            for params in self._all:
                for p in params:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # ------------------------------------------------------------------ persistent gradient storage
    def zero_grad(self) -> None:
        """Zero the flat buffers and (re-)point every .grad at its view: use instead of model.zero_grad(set_to_none=True)."""
        for flat, params, views in zip(self._flat, self._all, self._views):
            flat.zero_()
            for p, v in zip(params, views):
                p.grad = v

    def _adopt(self, bi: int, pi: int) -> None:
        """Make sure parameter (bi, pi)'s gradient lives in its view (the loop may have replaced .grad or left it None)."""
        p, v = self._all[bi][pi], self._views[bi][pi]
        g = p.grad
        if g is None:
            v.zero_()
        elif g.data_ptr() != v.data_ptr():
            v.copy_(g)
            p.grad = v

    # ------------------------------------------------------------------ collectives
    def _world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def set_batch(self, local_batch: int, global_batch: int) -> None:
        """Weight of this rank's (shard-mean) gradients in the global mean: n_r / N.  Call when shards are uneven."""
        self._scale = float(local_batch) / float(global_batch)

    def _launch(self, bi: int) -> None:
        if self._work[bi] is not None or self._world() == 1:
            return
        for pi in range(len(self._all[bi])):
            if pi not in self._seen[bi] or self._all[bi][pi].grad is None or \
                    self._all[bi][pi].grad.data_ptr() != self._views[bi][pi].data_ptr():
                self._adopt(bi, pi)
        flat = self._flat[bi]
        if self._scale is not None:
            flat.mul_(self._scale)
            op = dist.ReduceOp.SUM
        else:
            op = dist.ReduceOp.SUM          # divided by the world size in reduce() (gloo has no AVG)
            if dist.get_backend(self.group) == "nccl":
                op = dist.ReduceOp.AVG
        self._work[bi] = (dist.all_reduce(flat, op=op, group=self.group, async_op=True), op)

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if self._world() == 1:
            return
        bi, pi = self._bucket_of[id(p)]
        if pi in self._seen[bi]:
            return
        self._seen[bi].add(pi)
        if len(self._seen[bi]) == len(self._all[bi]):
            if p.is_cuda and torch.cuda.is_current_stream_capturing():
                if self._events is None:
                    self._events = [None] * len(self._all)
                    self._comm = torch.cuda.Stream(device=p.device)
                if self._events[bi] is None:
                    self._events[bi] = torch.cuda.Event(external=True)
                self._events[bi].record()    # becomes an event-record node of the graph: fires on every replay
            else:
                self._launch(bi)             # the whole bucket is final: its all-reduce runs under the rest of backward

    def reduce(self) -> int:
        """Finish the step's gradient exchange in place; returns the number of collectives issued."""
        world = self._world()
        if world == 1:
            for s in self._seen:
                s.clear()
            return 0
        none_mask = [[(p.grad is None) for p in params] for params in self._all]
        n = 0
        graph_mode = self._events is not None and not any(self._seen)
        if self._events is not None and any(self._seen):     # the step that was captured: remember which parameters it touched
            self._graph_seen = [set(s) for s in self._seen]
        if graph_mode:
            # a replayed step: hooks did not run; every bucket's all-reduce goes to the side stream behind the bucket's event
            cur = torch.cuda.current_stream()
            for bi in range(len(self._all)):
                if self._events[bi] is not None:
                    self._comm.wait_event(self._events[bi])
                else:
                    self._comm.wait_stream(cur)
                with torch.cuda.stream(self._comm):
                    self._seen[bi] = set(self._graph_seen[bi]) if self._graph_seen is not None else set(range(len(self._all[bi])))
                    self._launch(bi)
        for bi in range(len(self._all)):
            self._launch(bi)
        for bi, params in enumerate(self._all):
            work, op = self._work[bi]
            work.wait()
            if self._scale is None and op == dist.ReduceOp.SUM:
                self._flat[bi].div_(world)
            for pi, p in enumerate(params):
                untouched = (pi not in self._seen[bi]) if self._hooks else none_mask[bi][pi]
                if self.keep_none and untouched:
                    p.grad = None            # untouched on every rank: optimizers must keep skipping it
                else:
                    p.grad = self._views[bi][pi]
            self._work[bi] = None
            self._seen[bi].clear()
            n += 1
        return n

    def clip_(self, bound: float = 1.0) -> None:
        """GETA.grad_clipping (geta.py:160-165) on the flat buffers: one clamp per bucket instead of one per parameter."""
        for flat in self._flat:
            flat.clamp_(-bound, bound)


def clip_gradients_(params: Iterable[torch.nn.Parameter], bound: float = 1.0) -> None:
    """GETA.grad_clipping (geta.py:160-165): clamp every .grad to [-bound, bound]; runs AFTER the all-reduce."""
    for p in params:
        if p.grad is not None:
            p.grad.clamp_(-bound, bound)
