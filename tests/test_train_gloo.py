"""world_size-2 gloo tests (CPU) of the training-step plumbing (gfe_mamba_b200/train.py): per-layer gradient buckets whose
all-reduce is issued from backward hooks, against plain data-parallel averaging."""
import torch
import torch.distributed as dist

from test_parallel_gloo import _run

from gfe_mamba_b200.train import LayerGradSync, layer_buckets


class _Toy(torch.nn.Module):
    """Stand-in with the attribute layout LayerGradSync keys on (``.layers``); the real Mamba needs a GPU."""

    def __init__(self):
        super().__init__()
        self.layers = torch.nn.ModuleList(torch.nn.Linear(6, 6) for _ in range(3))
        self.head = torch.nn.Linear(6, 2)

    def forward(self, x):
        for l in self.layers:
            x = torch.tanh(l(x)) + x
        return self.head(x)


def test_layer_buckets_are_in_backward_order():
    m = _Toy()
    groups = layer_buckets(m)
    assert len(groups) == 4
    assert groups[0][0] is m.layers[2].weight and groups[2][0] is m.layers[0].weight and groups[3][0] is m.head.weight
    assert sum(len(g) for g in groups) == len(list(m.parameters()))


def _sync(rank, world):
    torch.manual_seed(0)
    model, ref = _Toy(), _Toy()
    ref.load_state_dict(model.state_dict())
    sync = LayerGradSync(model)
    assert all(p.grad is not None and p.grad.data_ptr() >= b.flat.data_ptr() for b in sync.buckets for p in b.params)
    for it in range(2):                               # twice: the buckets re-arm, gradients are zeroed in place in between
        x = torch.randn(5, 6, generator=torch.Generator().manual_seed(10 * it + rank))
        model(x).square().sum().backward()
        calls = sync.finish()
        assert calls == 4
        ref.zero_grad(set_to_none=True)
        ref(x).square().sum().backward()
        for p, q in zip(model.parameters(), ref.parameters()):
            g = q.grad.clone()
            dist.all_reduce(g)
            assert torch.allclose(p.grad, g / world, rtol=1e-5, atol=1e-6)
        sync.zero()
    sync.remove()


def test_layer_grad_sync_matches_data_parallel_average_gloo():
    _run(_sync)


def test_layer_grad_sync_single_process_is_a_no_op():
    model = _Toy()
    sync = LayerGradSync(model)
    model(torch.randn(3, 6)).sum().backward()
    assert sync.finish() == 0 and all(p.grad.abs().sum() > 0 for p in model.layers[0].parameters())
