"""Cycles per phase of the chained kernels (library built with -DGFE_PHASE_CLOCKS, GFE_LIB_VARIANT=clk).
Sums over warp leaders; printed as clocks per (warp, 16-step chunk)."""
import ctypes, os, sys, torch
os.environ["GFE_LIB_VARIANT"] = "clk"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gfe_mamba_b200 import selective_scan_fn, _native
lib = ctypes.CDLL(_native.LIB_PATH)
B, L, ED = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (16, 4096, 1536)
dt = torch.bfloat16 if (len(sys.argv) <= 4 or sys.argv[4] == "bf16") else torch.float32
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, device="cuda", generator=g) * sc).to(dt)
u, dl, z, dout = rn(B, L, ED).requires_grad_(), rn(B, L, ED, sc=0.5).requires_grad_(), rn(B, L, ED).requires_grad_(), rn(B, L, ED)
Bm, Cm = rn(B, L, 16).requires_grad_(), rn(B, L, 16).requires_grad_()
A = torch.log(torch.arange(1, 17, device="cuda").float()).repeat(ED, 1).requires_grad_()
D, b = torch.ones(ED, device="cuda", requires_grad=True), (torch.randn(ED, device="cuda", generator=g) * 0.3 - 3).requires_grad_()
def step():
    out = selective_scan_fn(u, dl, A, Bm, Cm, D, z=z, dt_bias=b)
    torch.autograd.grad(out, (u, dl, z, Bm, Cm, A, D, b), dout)
for _ in range(3): step()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 8)()
lib.gfe_debug_fwd_phase_clocks(None, 1); lib.gfe_debug_bwd_phase_clocks(None, 1)
n = 5
for _ in range(n): step()
torch.cuda.synchronize()
warp_chunks = n * B * (ED // 16) * ((L + 15) // 16)     # one warp serves 16 channels
for name, fn, labels in (("forward", lib.gfe_debug_fwd_phase_clocks, ["unit set-up + chain wait", "barrier (1)", "recurrence (16 steps)", "cp.async wait + barrier (2)", "epilogue", "phase A (items of the next chunk)", "unit tail", "refill (cp.async issue)"]),
                         ("backward", lib.gfe_debug_bwd_phase_clocks, ["unit set-up + chain wait", "barrier (1)", "fwd sweep 8..15 + (rev 8..15 | fwd 0..7)", "cp.async wait + barrier (2)", "row sums", "phase A (items of the next chunk)", "phase C 8..15 + rev sweep 0..7 + phase C 0..7", "refill (cp.async issue)"])):
    fn(buf, 0)
    tot = sum(buf)
    print(f"{name}: {tot / warp_chunks:.0f} clk per (warp, chunk)")
    for lab, v in zip(labels, buf):
        print(f"   {lab:36s} {v / warp_chunks:8.1f} clk  {100 * v / tot:5.1f} %")
