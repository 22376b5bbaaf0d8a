"""fwd+bwd time of one rank's channel shard of cfg4 (B=1, L=65536, ED=1024/N) on one GPU."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gfe_mamba_b200 import selective_scan_fn, _native
for ED in (1024, 512, 256, 128):
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    u, dl, z, dout = (rn(1, 65536, ED).requires_grad_() for _ in range(4))
    Bm, Cm = rn(1, 65536, 16).requires_grad_(), rn(1, 65536, 16).requires_grad_()
    A = (torch.log(torch.arange(1, 17, device="cuda").float()).repeat(ED, 1)).requires_grad_()
    D, b = torch.ones(ED, device="cuda", requires_grad=True), (rn(ED) * 0.3 - 3).requires_grad_()
    def step():
        out = selective_scan_fn(u, dl * 0.5, A, Bm, Cm, D, z=z, dt_bias=b)
        torch.autograd.grad(out, (u, dl, z, Bm, Cm, A, D, b), dout.detach())
    for _ in range(3): step()
    _native.timing_enable(True); _native.timing_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    k = _native.timing_collect(); _native.timing_enable(False)
    print("ED", ED, "ms/step", round(e0.elapsed_time(e1) / 10, 3), {n: round(v[0] / v[1], 3) for n, v in k.items()})
