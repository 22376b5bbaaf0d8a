"""Causal depthwise conv1d + SiLU: device time and achieved HBM GB/s against the measured peak, forward and
backward, on the strided first half of an in_proj output (as MambaBlock.forward calls it).
usage: python tools/bench_conv1d.py [B L ED dtype]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gfe_mamba_b200 import causal_conv1d_silu, _native

B, L, ED, dts = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]) if len(sys.argv) > 4 else (16, 4096, 1536, "bf16")
dt = {"f32": torch.float32, "bf16": torch.bfloat16}[dts]
s = torch.empty((), dtype=dt).element_size()
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
dev = "cuda"
sets = [dict(xz=torch.randn(B, L, 2 * ED, device=dev).to(dt).requires_grad_(), du=torch.randn(B, L, ED, device=dev).to(dt)) for _ in range(3)]
w = torch.randn(ED, 1, 4, device=dev, requires_grad=True)
bias = torch.randn(ED, device=dev, requires_grad=True)

def ours(d):
    u = causal_conv1d_silu(d["xz"][..., :ED], w, bias)
    torch.autograd.grad(u, (d["xz"], w, bias), d["du"])

def torch_ops(d):
    x = d["xz"][..., :ED].transpose(1, 2)
    u = torch.nn.functional.silu(torch.nn.functional.conv1d(x, w.to(dt), bias.to(dt), padding=3, groups=ED)[:, :, :L].transpose(1, 2))
    torch.autograd.grad(u, (d["xz"], w, bias), d["du"])

def timed(fn, n=10):
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
tok = B * L
alg = {"conv1d_silu_fwd": 2 * tok * ED * s, "conv1d_silu_bwd": 3 * tok * ED * s}
out = {"op": "causal conv1d + SiLU fwd+bwd", "B": B, "L": L, "ED": ED, "dtype": dts, "ms_per_step": round(t_ours, 4),
       "torch_ops_ms_per_step": round(t_torch, 4), "speedup_vs_torch_ops": round(t_torch / t_ours, 2), "peak_gbs": peak, "kernels": []}
for name, (ms, cnt) in kern.items():
    if not name.startswith("conv1d"): continue
    avg = ms / cnt
    a = alg.get(name)
    out["kernels"].append({"kernel": name, "avg_ms": round(avg, 4), "algorithmic_GBps": None if a is None else round(a / avg / 1e6, 1),
                           "frac_of_measured_peak": None if a is None else round(a / avg / 1e6 / peak, 3)})
print(json.dumps(out))
