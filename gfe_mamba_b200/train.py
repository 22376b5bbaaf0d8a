"""Training-step plumbing around the Mamba stack (host side of SURVEY 8e / 8f rank 3; the reference's loop is
classify_mamba.py:95-109): forward, backward, data-parallel gradient all-reduce OVERLAPPED with backward, multi-tensor
clip + Adam, and an optional CUDA-graph capture of the whole step for launch-bound shapes (the production shape, B = 2).

    sync = LayerGradSync(model)                       # one flat fp32 bucket per layer, all-reduced as its gradients land
    opt = ClipAdam(model.parameters(), lr=1e-4, max_norm=1.0, zero_grad=True)
    step = TrainStep(model, opt, grad_sync=sync, autocast_dtype=torch.bfloat16)
    loss = step(x)                                    # or GraphedTrainStep(step, x) ; loss = graphed(x)

PyTorch stays the plumbing (autograd, streams, NCCL through torch.distributed); every kernel on the Mamba path is this
library's or cuBLAS.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def default_loss(y: torch.Tensor) -> torch.Tensor:
    return y.float().square().mean()


class _Bucket:
    __slots__ = ("params", "flat", "pending", "handle", "numel")

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.numel = sum(p.numel() for p in params)
        dev = params[0].device
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        off = 0
        for p in params:   # gradients live INSIDE the bucket: autograd accumulates in place, nothing is packed or unpacked
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.pending = len(params)
        self.handle = None


def layer_buckets(model: torch.nn.Module) -> List[List[torch.nn.Parameter]]:
    """Parameters grouped per layer in BACKWARD order (last layer first): ``model.layers[i]`` when the model has layers
    (gfe_mamba_b200.Mamba), one bucket for everything else."""
    seen, groups = set(), []
    layers = list(getattr(model, "layers", []))
    for layer in reversed(layers):
        ps = [p for p in layer.parameters() if p.requires_grad and id(p) not in seen]
        seen.update(id(p) for p in ps)
        if ps:
            groups.append(ps)
    rest = [p for p in model.parameters() if p.requires_grad and id(p) not in seen]
    if rest:
        groups.append(rest)
    return groups


class LayerGradSync:
    """Sum (average) parameter gradients over the data-parallel group, one flat fp32 bucket per layer, each bucket
    all-reduced (async) the moment its last gradient has been accumulated -- so the collective of layer i runs under the
    backward kernels of layers i-1, i-2, ...  The all-reduce of the A_log / D / projection gradients is the ONLY
    collective of batch sharding; cfg 5: 8 buckets of ~6.8 MB."""

    def __init__(self, model: torch.nn.Module, group: Optional[dist.ProcessGroup] = None, average: bool = True,
                 groups: Optional[Sequence[Sequence[torch.nn.Parameter]]] = None, enabled: bool = True):
        self.group, self.average, self.enabled = group, average, enabled
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        for ps in (groups or layer_buckets(model)):
            for p in ps:
                if p.dtype != torch.float32:
                    raise TypeError("LayerGradSync: fp32 master parameters expected")
        self.buckets = [_Bucket(list(ps)) for ps in (groups or layer_buckets(model))]
        self._hooks = []
        for b in self.buckets:
            for p in b.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))

    def _make_hook(self, b: _Bucket) -> Callable:
        def hook(_p):
            b.pending -= 1
            if b.pending == 0 and self.enabled and self.world > 1:
                # issued from the autograd thread on the backward stream: NCCL's stream waits for the gradients written so
                # far and runs concurrently with the rest of backward
                b.handle = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return hook

    @property
    def bucket_bytes(self) -> List[int]:
        return [4 * b.numel for b in self.buckets]

    def finish(self) -> int:
        """Wait for the outstanding collectives (the current stream waits, the host does not), apply the average, and
        re-arm the buckets for the next backward.  Returns the number of all-reduce calls that were issued."""
        calls = 0
        for b in self.buckets:
            if b.pending != 0 and b.pending != len(b.params):
                raise RuntimeError("LayerGradSync.finish: a bucket received only part of its gradients (unused parameters?)")
            if b.handle is not None:
                b.handle.wait()
                b.handle = None
                calls += 1
                if self.average:
                    b.flat.div_(self.world)
            b.pending = len(b.params)
        return calls

    def zero(self) -> None:
        for b in self.buckets:
            b.flat.zero_()

    def allreduce_now(self) -> None:
        """The same collectives without any backward to hide under (measurement: total all-reduce time)."""
        if self.world > 1:
            for b in self.buckets:
                dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


class TrainStep:
    """forward -> loss -> backward (gradient buckets all-reduced underneath) -> clip + Adam.  ``optimizer`` is expected to
    leave the gradients zeroed in place (ClipAdam(zero_grad=True)); otherwise they are zeroed here before backward."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable = default_loss,
                 grad_sync: Optional[LayerGradSync] = None, autocast_dtype: Optional[torch.dtype] = None):
        self.model, self.optimizer, self.loss_fn, self.grad_sync, self.autocast_dtype = model, optimizer, loss_fn, grad_sync, autocast_dtype
        self._self_zeroing = all(g.get("zero_grad", False) for g in optimizer.param_groups)
        if grad_sync is None:   # fixed gradient buffers all the same (CUDA graphs, no per-step allocation)
            for p in model.parameters():
                if p.requires_grad and p.grad is None:
                    p.grad = torch.zeros_like(p)

    def forward_backward(self, x: torch.Tensor, *loss_args) -> torch.Tensor:
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            y = self.model(x)
        loss = self.loss_fn(y, *loss_args)
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync.finish()
        return loss.detach()

    def __call__(self, x: torch.Tensor, *loss_args) -> torch.Tensor:
        if not self._self_zeroing:
            self.optimizer.zero_grad(set_to_none=False)
        loss = self.forward_backward(x, *loss_args)
        self.optimizer.step()
        return loss


class GraphedTrainStep:
    """One whole training step (forward, backward, clip + Adam) captured in a CUDA graph and replayed: at the production
    shape (B = 2, L = 1858, 6 layers) the step is launch-bound -- ~70 % of its 8.3 ms was host overhead between ~600 small
    launches.  Static shapes only; single GPU (NCCL inside a capture is not attempted here)."""

    def __init__(self, step: TrainStep, example_x: torch.Tensor, *example_loss_args, warmup: int = 3):
        if step.grad_sync is not None and step.grad_sync.world > 1:
            raise RuntimeError("GraphedTrainStep: capture the single-GPU step; multi-GPU steps overlap NCCL eagerly")
        self.step = step
        self.x = example_x.detach().clone()
        self.loss_args = tuple(a.detach().clone() for a in example_loss_args)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up outside the capture: builds the optimiser tables, the launch caches, cuBLAS handles
            for _ in range(warmup):
                step(self.x, *self.loss_args)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = step(self.x, *self.loss_args)

    def __call__(self, x: torch.Tensor, *loss_args) -> torch.Tensor:
        self.x.copy_(x)
        for dst, src in zip(self.loss_args, loss_args):
            dst.copy_(src)
        self.graph.replay()
        return self.loss
