"""Host logic of the multi-tensor optimiser (no GPU): the chunk table that describes a parameter list to the kernels."""
import pytest
import torch


def test_chunk_plan_covers_every_element_once():
    from gfe_mamba_b200.optim import ClipAdam
    numels = [1, 4096, 4097, 10000, 3]
    ct, cs, tc0 = ClipAdam.plan_chunks(numels, 4096)
    assert tc0 == [0, 1, 2, 4, 7, 8] and len(ct) == len(cs) == 8
    for t, n in enumerate(numels):
        starts = [s for tt, s in zip(ct, cs) if tt == t]
        assert starts == list(range(0, n, 4096))
        assert ct[tc0[t]:tc0[t + 1]] == [t] * (tc0[t + 1] - tc0[t])


def test_clip_adam_rejects_cpu_parameters():
    from gfe_mamba_b200.optim import ClipAdam
    p = torch.nn.Parameter(torch.ones(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ClipAdam([p], lr=1e-3).step()
    with pytest.raises(ValueError):
        ClipAdam([p], lr=-1.0)
