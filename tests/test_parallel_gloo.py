"""world_size-2 gloo tests (CPU) for the multi-GPU host logic: sharding arithmetic, the flat-bucket gradient
all-reduce used by batch sharding, and the dB/dC all-reduce wiring of channel sharding (with a pure-torch stand-in
for the scan, since the real one needs a GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gfe_mamba_b200 import parallel as par


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world=2):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn), nprocs=world, join=True)


def _entry(rank, world, port, fn):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    for total in (1, 7, 16, 17, 256):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.shard_range(4, 2, 2)


def test_shard_channels_requires_warp_multiple():
    t = torch.zeros(1, 4, 1024)
    assert par.shard_channels(t, 3, 8).shape[-1] == 128
    with pytest.raises(ValueError):
        par.shard_channels(torch.zeros(1, 4, 96), 0, 2)


def _grad_sync(rank, world):
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4))
    frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)
    for i, p in enumerate(model.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    list(model.parameters())[1].grad = None if rank == 1 else list(model.parameters())[1].grad   # a rank without this grad
    calls = par.allreduce_gradients(list(model.parameters()) + [frozen], average=True, bucket_bytes=16 * 4 * 8)
    assert calls >= 2        # small bucket size forces several buckets
    for i, p in enumerate(model.parameters()):
        want = (1 + 2) * (i + 1) / 2 if i != 1 else (1 * 2) / 2
        assert torch.allclose(p.grad, torch.full_like(p, want)), (i, p.grad.flatten()[0].item(), want)
    assert frozen.grad is None


def test_flat_bucket_allreduce_gloo():
    _run(_grad_sync)


def _toy_scan(u, delta, A_log, Bm, Cm, D, z=None, dt_bias=None):
    # pure-torch stand-in with the same sharding structure: per-channel math, B/C shared over channels
    y = (u * delta).unsqueeze(-1) * Bm.unsqueeze(2) * Cm.unsqueeze(2)        # (B, L, ED, N)
    return y.sum(-1) * torch.exp(-torch.exp(A_log)).sum(-1) + D * u


def _channel_shard(rank, world):
    torch.manual_seed(1)
    B, L, ED, N = 2, 5, 64, 4
    u, delta = torch.randn(B, L, ED), torch.rand(B, L, ED)
    A_log, D = torch.randn(ED, N) * 0.1, torch.randn(ED)
    Bm, Cm = torch.randn(B, L, N, requires_grad=True), torch.randn(B, L, N, requires_grad=True)
    dout = torch.randn(B, L, ED)
    # single-process truth
    _toy_scan(u, delta, A_log, Bm, Cm, D).backward(dout)
    dB_full, dC_full = Bm.grad.clone(), Cm.grad.clone()
    Bm.grad = Cm.grad = None
    sl = lambda t, dim=-1: par.shard_channels(t, rank, world, dim=dim, multiple=32)
    out = par.channel_sharded_scan(_toy_scan, sl(u), sl(delta), sl(A_log, 0), Bm, Cm, sl(D))
    out.backward(sl(dout))
    assert torch.allclose(Bm.grad, dB_full, atol=1e-5) and torch.allclose(Cm.grad, dC_full, atol=1e-5)
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out.detach())
    assert torch.allclose(torch.cat(gathered, -1), _toy_scan(u, delta, A_log, Bm.detach(), Cm.detach(), D), atol=1e-5)


def test_channel_sharded_scan_allreduces_bc_grads_gloo():
    _run(_channel_shard)


def _batch_shard(rank, world):
    x = torch.arange(10 * 3, dtype=torch.float32).view(10, 3)
    mine = par.shard_batch(x, rank, world)
    sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.shape[0]]))
    assert sum(int(s) for s in sizes) == 10
    lo, hi = par.shard_range(10, rank, world)
    assert torch.equal(mine, x[lo:hi])


def test_batch_sharding_gloo():
    _run(_batch_shard)
