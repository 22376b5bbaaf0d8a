"""Fused residual add + RMSNorm: device time and achieved HBM GB/s against the measured peak, beside the
torch ops it replaces (add, pow, mean, rsqrt, mul, mul).  usage: python tools/bench_addnorm.py [rows D dtype]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gfe_mamba_b200 import add_rmsnorm, _native

rows, D, dts = (int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]) if len(sys.argv) > 3 else (65536, 768, "bf16")
dt = {"f32": torch.float32, "bf16": torch.bfloat16}[dts]
s = torch.empty((), dtype=dt).element_size()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = "cuda"
sets = [dict(x=torch.randn(rows, D, device=dev).to(dt).requires_grad_(), a=torch.randn(rows, D, device=dev).to(dt).requires_grad_(),
             dres=torch.randn(rows, D, device=dev).to(dt), dy=torch.randn(rows, D, device=dev).to(dt)) for _ in range(3)]
w = torch.ones(D, device=dev, requires_grad=True)

def ours(d):
    resid, y = add_rmsnorm(d["x"], d["a"], w, 1e-5)
    torch.autograd.grad((resid, y), (d["x"], d["a"], w), (d["dres"], d["dy"]))

def torch_ops(d):
    resid = d["x"] + d["a"]
    y = resid * torch.rsqrt(resid.pow(2).mean(-1, keepdim=True) + 1e-5) * w
    torch.autograd.grad((resid, y), (d["x"], d["a"], w), (d["dres"], d["dy"]))

def timed(fn, n=20):
    for i in range(3): fn(sets[i % 3])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for i in range(n): fn(sets[i % 3])
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for i in range(3): ours(sets[i % 3])
torch.cuda.synchronize()   # first launches (module load, allocator growth) stay out of the per-kernel averages
_native.timing_enable(True); _native.timing_collect()
t_ours = timed(ours)
kern = _native.timing_collect(); _native.timing_enable(False)
t_torch = timed(torch_ops)
fwd_b, bwd_b = 4 * rows * D * s, 4 * rows * D * s
out = {"op": "add_rmsnorm fwd+bwd", "rows": rows, "D": D, "dtype": dts, "ms_per_step": round(t_ours, 4), "torch_ops_ms_per_step": round(t_torch, 4),
       "speedup_vs_torch_ops": round(t_torch / t_ours, 2), "peak_gbs": peak, "kernels": []}
for name, (ms, cnt) in kern.items():
    if not name.startswith("add_rmsnorm"): continue
    avg = ms / cnt
    alg = {"add_rmsnorm_fwd": fwd_b, "add_rmsnorm_bwd": bwd_b}.get(name)
    out["kernels"].append({"kernel": name, "avg_ms": round(avg, 4), "algorithmic_GBps": None if alg is None else round(alg / avg / 1e6, 1),
                           "frac_of_measured_peak": None if alg is None else round(alg / avg / 1e6 / peak, 3)})
print(json.dumps(out))
