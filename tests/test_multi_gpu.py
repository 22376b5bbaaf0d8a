"""Two-GPU NCCL tests of the sharded selective scan: channel sharding (BASELINE config 4: one long sequence,
rank g owns ED/G channels, dB/dC all-reduced in backward) and batch sharding (configs 3/5: rank g owns B/G rows, parameter
gradients all-reduced) against the same computation on one GPU.  Skipped on a box with fewer than two GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(B, L, ED, N, dev):
    g = torch.Generator(device="cpu").manual_seed(123)
    rn = lambda *s: torch.randn(*s, generator=g)
    d = dict(u=rn(B, L, ED), delta=0.5 * rn(B, L, ED), z=rn(B, L, ED), Bm=rn(B, L, N), Cm=rn(B, L, N), dout=rn(B, L, ED),
             A_log=torch.log(torch.arange(1, N + 1).float()).repeat(ED, 1) + 0.1 * rn(ED, N), D=1 + 0.1 * rn(ED), bias=-4 + rn(ED))
    return {k: v.to(dev) for k, v in d.items()}


def _full(d):
    from gfe_mamba_b200 import selective_scan_fn
    leaves = {k: d[k].clone().requires_grad_() for k in ("u", "delta", "z", "Bm", "Cm", "A_log", "D", "bias")}
    out = selective_scan_fn(leaves["u"], leaves["delta"], leaves["A_log"], leaves["Bm"], leaves["Cm"], leaves["D"],
                            z=leaves["z"], dt_bias=leaves["bias"])
    out.backward(d["dout"])
    return out.detach(), {k: v.grad for k, v in leaves.items()}


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def _worker(rank, world, port, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from gfe_mamba_b200 import parallel as par, selective_scan_fn
        if mode == "channel":
            B, L, ED, N = 1, 2048, 256, 16
            d = _inputs(B, L, ED, N, dev)
            ref_out, ref_g = _full(d)
            sh = lambda t, dim=-1: par.shard_channels(t, rank, world, dim=dim).contiguous().clone().requires_grad_()
            u, delta, z = sh(d["u"]), sh(d["delta"]), sh(d["z"])
            A_log, D, bias = sh(d["A_log"], 0), sh(d["D"], 0), sh(d["bias"], 0)
            Bm, Cm = d["Bm"].clone().requires_grad_(), d["Cm"].clone().requires_grad_()
            out = par.channel_sharded_scan(selective_scan_fn, u, delta, A_log, Bm, Cm, D, z=z, dt_bias=bias)
            out.backward(par.shard_channels(d["dout"], rank, world).contiguous())
            lo, hi = rank * ED // world, (rank + 1) * ED // world
            assert _rel(out, ref_out[..., lo:hi]) < 1e-5
            assert _rel(u.grad, ref_g["u"][..., lo:hi]) < 1e-4 and _rel(delta.grad, ref_g["delta"][..., lo:hi]) < 1e-4
            assert _rel(A_log.grad, ref_g["A_log"][lo:hi]) < 1e-4 and _rel(D.grad, ref_g["D"][lo:hi]) < 1e-4
            assert _rel(Bm.grad, ref_g["Bm"]) < 1e-4 and _rel(Cm.grad, ref_g["Cm"]) < 1e-4   # summed over both ranks' channels
        else:
            B, L, ED, N = 4, 512, 128, 16
            d = _inputs(B, L, ED, N, dev)
            ref_out, ref_g = _full(d)
            rows = lambda t: par.shard_batch(t, rank, world).contiguous().clone().requires_grad_()
            u, delta, z, Bm, Cm = (rows(d[k]) for k in ("u", "delta", "z", "Bm", "Cm"))
            params = [torch.nn.Parameter(d[k].clone()) for k in ("A_log", "D", "bias")]
            out = selective_scan_fn(u, delta, params[0], Bm, Cm, params[1], z=z, dt_bias=params[2])
            out.backward(par.shard_batch(d["dout"], rank, world).contiguous())
            assert par.allreduce_gradients(params, average=False) == 1
            lo, hi = par.shard_range(B, rank, world)
            assert _rel(out, ref_out[lo:hi]) < 1e-5 and _rel(u.grad, ref_g["u"][lo:hi]) < 1e-4
            for p, k in zip(params, ("A_log", "D", "bias")):
                assert _rel(p.grad, ref_g[k]) < 1e-4, k
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["channel", "batch"])
def test_two_gpu_sharding_matches_single_gpu(mode):
    mp.spawn(_worker, args=(2, _free_port(), mode), nprocs=2, join=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_process_two_devices_same_shape():
    """One host thread runs the same shapes on cuda:0 and then cuda:1: the launch caches (dynamic-shared-memory opt-in,
    persistent grid size, workspace sizes) are per device, so the second device must not inherit the first one's."""
    from gfe_mamba_b200 import Mamba, MambaConfig
    shapes = [(24, 203, 512), (2, 300, 64)]          # chained kernels (57-88 KB of dynamic shared memory) and the L-split pair
    ref = {}
    for di in (0, 1):
        dev = torch.device("cuda", di)
        for (B, L, ED) in shapes:
            d = _inputs(B, L, ED, 16, dev)
            out, g = _full(d)
            torch.cuda.synchronize(dev)
            key = (B, L, ED)
            if di == 0:
                ref[key] = (out.cpu(), {k: v.cpu() for k, v in g.items()})
            else:
                assert _rel(out.cpu(), ref[key][0]) < 1e-6
                for k, v in g.items():
                    assert _rel(v.cpu(), ref[key][1][k]) < 1e-5, (key, k)
        torch.manual_seed(0)
        m = Mamba(MambaConfig(d_model=64, n_layers=2)).to(dev)
        y = m(torch.randn(2, 50, 64, device=dev))
        assert torch.isfinite(y).all()
