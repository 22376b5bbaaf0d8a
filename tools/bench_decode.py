"""Decode throughput of Mamba.step: eager loop against the CUDA-graph loop (gfe_mamba_b200.decode.GraphedDecoder).
usage: python tools/bench_decode.py [batch d_model n_layers tokens]"""
import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gfe_mamba_b200 import Mamba, MambaConfig
from gfe_mamba_b200.decode import GraphedDecoder
B, D, NL, T = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (16, 768, 24, 64)
torch.manual_seed(0)
cfg = MambaConfig(d_model=D, n_layers=NL)
model = Mamba(cfg).cuda().eval()
x = torch.randn(B, T, D, device="cuda")
def eager():
    caches = [(None, torch.zeros(B, cfg.d_inner, cfg.d_conv - 1, device="cuda")) for _ in range(NL)]
    with torch.no_grad():
        for t in range(T):
            y, caches = model.step(x[:, t], caches)
    return y
dec = GraphedDecoder(model, B)
def graphed():
    for t in range(T):
        y = dec.step(x[:, t])
    return y
def timed(fn):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    return time.perf_counter() - t0
te, tg = timed(eager), timed(graphed)
print(json.dumps({"op": "Mamba.step decode", "batch": B, "d_model": D, "n_layers": NL, "tokens": T,
                  "eager_ms_per_token": round(te / T * 1e3, 4), "graph_ms_per_token": round(tg / T * 1e3, 4),
                  "eager_tokens_per_s": round(B * T / te), "graph_tokens_per_s": round(B * T / tg), "speedup": round(te / tg, 2)}))
