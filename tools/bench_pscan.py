"""pscan(A, X): device time and achieved HBM GB/s against the measured peak, forward and backward.
usage: python tools/bench_pscan.py [B L D N]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gfe_mamba_b200 import pscan, _native
B, L, D, N = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (4, 1024, 1024, 16)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
dev = "cuda"
sets = [dict(A=torch.rand(B, L, D, N, device=dev).mul_(0.5).add_(0.4).requires_grad_(), X=torch.randn(B, L, D, N, device=dev).requires_grad_(),
             dH=torch.randn(B, L, D, N, device=dev)) for _ in range(2)]
def ours(d):
    H = pscan(d["A"], d["X"])
    torch.autograd.grad(H, (d["A"], d["X"]), d["dH"])
def timed(fn, n=6):
    for i in range(2): fn(sets[i % 2])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for i in range(n): fn(sets[i % 2])
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for i in range(3): ours(sets[i % 2])
torch.cuda.synchronize()   # first launches (module load, allocator growth) stay out of the per-kernel averages
_native.timing_enable(True); _native.timing_collect()
t = timed(ours)
kern = _native.timing_collect(); _native.timing_enable(False)
el = B * L * D * N * 4
alg = {"pscan_fwd": 3 * el, "pscan_bwd": 5 * el}
out = {"op": "pscan fwd+bwd", "B": B, "L": L, "D": D, "N": N, "ms_per_step": round(t, 4), "peak_gbs": peak, "kernels": []}
for name, (ms, cnt) in kern.items():
    if not name.startswith("pscan"): continue
    avg = ms / cnt; a = alg.get(name)
    out["kernels"].append({"kernel": name, "avg_ms": round(avg, 4), "algorithmic_GBps": None if a is None else round(a / avg / 1e6, 1),
                           "frac_of_measured_peak": None if a is None else round(a / avg / 1e6 / peak, 3)})
print(json.dumps(out))
