"""Whole-stack training step (BASELINE config 5 shape: Mamba(d_model=512, n_layers=8), B x L tokens, fp32 or bf16 autocast):
forward + backward through gfe_mamba_b200.Mamba, device time per step and the share spent in this library's kernels
(the rest is cuBLAS GEMMs and torch glue).  usage: python tools/bench_model.py [B L d_model n_layers dtype]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gfe_mamba_b200 import Mamba, MambaConfig, _native
B, L, D, NL = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (64, 1024, 512, 8)
dts = sys.argv[5] if len(sys.argv) > 5 else "f32"
torch.manual_seed(0)
model = Mamba(MambaConfig(d_model=D, n_layers=NL)).cuda()
x = torch.randn(B, L, D, device="cuda")
def step():
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dts == "bf16"):
        y = model(x)
    y.float().square().mean().backward()
for _ in range(3): step()
torch.cuda.synchronize()
_native.timing_enable(True); _native.timing_collect()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
a.record()
for _ in range(n): step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
kern = _native.timing_collect(); _native.timing_enable(False)
ours = {k: round(v[0] / n, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])}
tot = sum(ours.values())
print(json.dumps({"op": "Mamba stack fwd+bwd", "B": B, "L": L, "d_model": D, "n_layers": NL, "dtype": dts, "ms_per_step": round(ms, 3),
                  "tokens_per_s": round(B * L / ms * 1e3), "library_kernels_ms": round(tot, 3), "library_share": round(tot / ms, 3),
                  "kernels_ms_per_step": ours}))
