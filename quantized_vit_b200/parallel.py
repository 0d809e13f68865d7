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
    """Bucketed gradient averaging for QAT.  Build once per model; call ``reduce()`` after every backward."""

    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], bucket_bytes: int = 64 << 20, group=None):
        self.group = group
        self.quant: List[torch.nn.Parameter] = []
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, cur_bytes = [], 0
        for name, p in named_params:
            if not p.requires_grad:
                continue
            if any(tag in name for tag in QUANT_PARAM_TAGS):
                self.quant.append(p)
                continue
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)

    @staticmethod
    def _flatten(params: Sequence[torch.nn.Parameter]) -> torch.Tensor:
        dev = params[0].device
        flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            n = p.numel()
            if p.grad is None:
                flat[off:off + n].zero_()          # a rank that did not touch a parameter contributes zero
            else:
                flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        return flat

    @staticmethod
    def _unflatten(flat: torch.Tensor, params: Sequence[torch.nn.Parameter]) -> None:
        off = 0
        for p in params:
            n = p.numel()
            g = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n

    def reduce(self) -> int:
        """Average gradients over the group in place; returns the number of collectives issued."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return 0
        work = []
        groups = list(self.buckets) + ([self.quant] if self.quant else [])
        for params in groups:                       # issue everything first: NVSwitch collectives overlap each other
            flat = self._flatten(params)
            work.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), flat, params))
        for w, flat, params in work:
            w.wait()
            flat.div_(world)
            self._unflatten(flat, params)
        return len(work)


def clip_gradients_(params: Iterable[torch.nn.Parameter], bound: float = 1.0) -> None:
    """GETA.grad_clipping (geta.py:160-165): clamp every .grad to [-bound, bound]; runs AFTER the all-reduce."""
    for p in params:
        if p.grad is not None:
            p.grad.clamp_(-bound, bound)
