"""Multi-GPU plumbing for the selective-scan path: one process per GPU, torch.distributed (NCCL on GPUs,
gloo in the CPU unit tests).  The recurrence is independent per (batch row, channel), so the data path never
needs a collective; only gradients are exchanged:

  batch sharding    each rank runs its own B/G sequences; after backward ONE flat-bucket all-reduce sums the
                    parameter gradients (A_log, D, conv1d, dt_proj, x_proj, in_proj, out_proj, norm).
  channel sharding  (one long sequence, BASELINE config 4) rank g owns channels [g*ED/G, (g+1)*ED/G) of
                    u, delta, z, A_log, D, dt_bias and out; B and C are shared, so their gradients -- sums over
                    ALL channels -- are all-reduced in backward (2 * B * L * N values).

The reference has no distributed code on this path (classify_mamba.py:69-73 is commented out); this module is
what its Accelerate/DDP use elsewhere (main_gan_vit.py:31,54) would provide.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of ``total`` items for ``rank`` (first ``total % world`` ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rows of dim 0 owned by ``rank`` under batch sharding."""
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


def shard_channels(t: torch.Tensor, rank: int, world: int, dim: int = -1, multiple: int = 32) -> torch.Tensor:
    """Channel slice owned by ``rank`` under ED sharding.  ED / world must be a multiple of ``multiple`` (32 keeps every
    rank on the fast kernels: one warp serves 32 adjacent channels)."""
    ED = t.shape[dim]
    if ED % (world * multiple) != 0:
        raise ValueError(f"ED={ED} is not divisible into {world} shards of a multiple of {multiple} channels")
    per = ED // world
    return t.narrow(dim, rank * per, per)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                        average: bool = True, bucket_bytes: int = 256 << 20) -> int:
    """Sum (or average) the gradients of ``params`` over the group with as few collectives as possible: gradients are
    packed into flat fp32 buckets (one bucket covers a whole Mamba stack), all-reduced, and unpacked in place.
    Parameters without a gradient contribute zeros so that every rank issues identical collectives.
    Returns the number of all-reduce calls issued."""
    params = [p for p in params if p.requires_grad]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    calls, i = 0, 0
    while i < len(params):
        bucket: List[torch.nn.Parameter] = []
        size = 0
        while i < len(params) and (not bucket or (size + params[i].numel()) * 4 <= bucket_bytes):
            bucket.append(params[i])
            size += params[i].numel()
            i += 1
        dev = bucket[0].device
        flat = torch.zeros(size, dtype=torch.float32, device=dev)
        off = 0
        for p in bucket:
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        calls += 1
        if average:
            flat.div_(world)
        off = 0
        for p in bucket:
            g = flat[off:off + p.numel()].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
    return calls


class _AllReduceGrad(torch.autograd.Function):
    """Identity in forward; sums the incoming gradient over the group in backward (for tensors replicated across
    ranks whose consumers are sharded: B and C under channel sharding)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        if dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            g = g.clone()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


def channel_sharded_scan(scan_fn: Callable, u, delta, A_log, Bm, Cm, D, z=None, dt_bias=None,
                         group: Optional[dist.ProcessGroup] = None, **kw):
    """Run ``scan_fn`` (gfe_mamba_b200.selective_scan_fn) on THIS rank's channel shard.

    ``u, delta, z`` are already the local shards (B, L, ED/G); ``A_log, D, dt_bias`` the local rows; ``Bm, Cm`` are the
    full, replicated (B, L, N) tensors.  The output is the local (B, L, ED/G) shard.  In backward the gradients of
    Bm and Cm are all-reduced (they sum over every rank's channels); nothing else is exchanged."""
    Bm = _AllReduceGrad.apply(Bm, group)
    Cm = _AllReduceGrad.apply(Cm, group)
    return scan_fn(u, delta, A_log, Bm, Cm, D, z=z, dt_bias=dt_bias, **kw)


def parameter_gradient_names(model: torch.nn.Module) -> Sequence[str]:
    """Names of the parameters whose gradients the batch-sharded step all-reduces (everything trainable)."""
    return [n for n, p in model.named_parameters() if p.requires_grad]
