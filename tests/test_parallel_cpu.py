"""world_size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding/gather and the QAT gradient
all-reduce (weight grads in buckets + packed step-size grads), N-rank average == 1-rank gradient of the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quantized_vit_b200 import parallel


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 256, 1025):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


class _Toy(torch.nn.Module):
    """Parameter names mimic a QuantizeLinear (weight, bias, d_quant_wt, q_m_wt, d_quant_act, q_m_act)."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.weight = torch.nn.Parameter(torch.randn(6, 5, generator=g))
        self.bias = torch.nn.Parameter(torch.randn(6, generator=g))
        self.d_quant_wt = torch.nn.Parameter(torch.tensor([0.3]))
        self.q_m_wt = torch.nn.Parameter(torch.tensor([1.7]))
        self.d_quant_act = torch.nn.Parameter(torch.tensor([0.2]))
        self.q_m_act = torch.nn.Parameter(torch.tensor([1.1]))
        self.unused = torch.nn.Parameter(torch.zeros(3))

    def forward(self, x):
        # smooth stand-in for the quantized layer: every parameter (incl. the scalar "step sizes") gets a batch-dependent grad
        w = self.weight * self.d_quant_wt + torch.tanh(self.weight) * self.q_m_wt
        return torch.nn.functional.linear(x * self.d_quant_act + torch.sin(x) * self.q_m_act, w, self.bias)


def _worker(rank, world, port, q, n=10):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(1)
        x = torch.randn(n, 5)
        y = torch.randn(n, 6)
        model = _Toy()
        red = parallel.GradientAllReducer(model.named_parameters(), bucket_bytes=64)      # tiny buckets -> several collectives
        xs, ys = parallel.shard_batch(x, rank, world), parallel.shard_batch(y, rank, world)
        if n % world:
            red.set_batch(xs.shape[0], n)          # uneven shards: weight n_r / N instead of 1 / world
        # step 1 the way an unchanged loop does it (fresh .grad tensors), step 2 on the persistent bucket views
        for it in range(2):
            if it == 0:
                model.zero_grad(set_to_none=True)
            else:
                red.zero_grad()
                assert model.weight.grad.data_ptr() == red._views[red._bucket_of[id(model.weight)][0]][red._bucket_of[id(model.weight)][1]].data_ptr()
            # mean over the GLOBAL batch = weighted average over ranks of the per-rank mean
            loss = ((model(xs) - ys) ** 2).mean()
            loss.backward()
            n_coll = red.reduce()
        red.clip_(1.0)
        out = parallel.gather_outputs(model(xs).detach(), x.shape[0])
        # plain numpy payloads: torch tensors travel through a Queue as shared-memory handles that die with this process
        grads = {n: (p.grad.numpy().copy() if p.grad is not None else None) for n, p in model.named_parameters()}
        q.put((rank, n_coll, grads, None if out is None else out.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11])
def test_two_rank_allreduce_equals_single_process(n):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, n)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process reference on the concatenated batch
    torch.manual_seed(1)
    x = torch.randn(n, 5)
    y = torch.randn(n, 6)
    model = _Toy()
    ((model(x) - y) ** 2).mean().backward()
    parallel.clip_gradients_(model.parameters(), 1.0)
    for rank, n_coll, grads, out in res:
        assert n_coll >= 3                                      # >= 2 weight buckets + the packed quant-scalar buffer
        for n, p in model.named_parameters():
            if n == "unused":
                assert grads[n] is None                          # untouched on every rank: stays None (optimizers skip it)
                continue
            assert torch.allclose(torch.from_numpy(grads[n]), p.grad, rtol=1e-5, atol=1e-6), n
    assert torch.allclose(torch.from_numpy(res[0][3]), model(x).detach(), rtol=1e-5, atol=1e-6)       # gathered in batch order on rank 0
    assert res[1][3] is None
