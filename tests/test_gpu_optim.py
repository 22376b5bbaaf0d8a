"""GPU parity of the multi-tensor clip + Adam step (csrc/optim.cu) against the reference training loop's own calls:
per-parameter torch.nn.utils.clip_grad_norm_ followed by torch.optim.Adam.step (classify_mamba.py:104-109)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(seed, shapes, scale):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    gs = [[torch.randn(s, device="cuda", generator=g) * sc for s, sc in zip(shapes, scale)] for _ in range(4)]
    return ps, gs


@pytest.mark.parametrize("per_param", [True, False])
@pytest.mark.parametrize("max_norm", [1.0, None])
def test_clip_adam_matches_torch(per_param, max_norm):
    from gfe_mamba_b200.optim import ClipAdam
    shapes = [(1024, 16), (1024,), (2048, 512), (1024, 1, 4), (7,), (48 + 32, 1024), (4097,), (3, 5, 7)]
    scale = [10.0, 1e-3, 1.0, 0.1, 5.0, 1e-2, 1.0, 100.0]   # some tensors clipped hard, some not at all
    ours, grads = _params(3, shapes, scale)
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt_ref = torch.optim.Adam(ref, lr=1e-2)
    opt = ClipAdam(ours, lr=1e-2, max_norm=max_norm, per_parameter_clip=per_param, zero_grad=False)
    for step in range(4):
        for p, q, g in zip(ours, ref, grads[step]):
            p.grad = g.clone()
            q.grad = g.clone()
        if max_norm is not None:
            if per_param:
                for q in ref:   # the reference's loop, classify_mamba.py:106-107
                    torch.nn.utils.clip_grad_norm_(q, max_norm=max_norm)
            else:
                torch.nn.utils.clip_grad_norm_(ref, max_norm=max_norm)
        opt_ref.step()
        opt.step()
        torch.cuda.synchronize()
        for p, q in zip(ours, ref):
            assert torch.allclose(p, q, rtol=2e-6, atol=2e-7), (step, p.shape, (p - q).abs().max().item())
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-8), "clipped gradients differ"
    st = opt.state[ours[2]]
    assert torch.allclose(st["exp_avg"], opt_ref.state[ref[2]]["exp_avg"], rtol=1e-5, atol=1e-6)   # lerp vs fma rounding near zero
    assert torch.allclose(st["exp_avg_sq"], opt_ref.state[ref[2]]["exp_avg_sq"], rtol=1e-4, atol=1e-8)


def test_clip_adam_zero_grad_and_graph_capture():
    """zero_grad=True leaves the gradients zeroed in place (fixed addresses); a captured step replays with the right bias
    corrections because the step counter lives on the device."""
    from gfe_mamba_b200.optim import ClipAdam
    shapes = [(300, 16), (300,), (64, 300)]
    ours, grads = _params(5, shapes, [1.0, 1.0, 1.0])
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt_ref = torch.optim.Adam(ref, lr=1e-3)
    opt = ClipAdam(ours, lr=1e-3, max_norm=1.0, zero_grad=True)
    static_g = [torch.zeros_like(p) for p in ours]
    for p, g in zip(ours, static_g):
        p.grad = g
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):   # warm-up outside the capture (builds the tables)
        for g, src in zip(static_g, grads[0]):
            g.copy_(src)
        opt.step()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    for q, src in zip(ref, grads[0]):
        q.grad = src.clone()
    for q in ref:
        torch.nn.utils.clip_grad_norm_(q, max_norm=1.0)
    opt_ref.step()
    for step in range(1, 4):
        for g, src in zip(static_g, grads[step]):
            assert float(g.abs().max()) == 0.0, "gradients were not zeroed by the previous step"
            g.copy_(src)
        graph.replay()
        for q, src in zip(ref, grads[step]):
            q.grad = src.clone()
        for q in ref:
            torch.nn.utils.clip_grad_norm_(q, max_norm=1.0)
        opt_ref.step()
    torch.cuda.synchronize()
    for p, q in zip(ours, ref):
        assert torch.allclose(p, q, rtol=2e-6, atol=2e-7), (p - q).abs().max().item()
