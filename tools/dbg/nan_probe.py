import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from test_gpu_parity import make_scan_inputs, run_fused
for (B, L, ED) in ((20, 523, 512), (20, 512, 512), (20, 128, 512), (24, 64, 512)):
    d = make_scan_inputs(B, L, ED, seed=12)
    g = run_fused(d, torch.bfloat16)
    for k, v in g.items():
        if v is None: continue
        bad = ~torch.isfinite(v.float())
        if bad.any():
            idx = bad.nonzero()
            print(B, L, ED, k, "non-finite:", int(bad.sum()), "first", idx[0].tolist(), "last", idx[-1].tolist(),
                  "t values", sorted(set(idx[:, 1].tolist()))[:20] if idx.shape[1] > 1 else "")
        else:
            print(B, L, ED, k, "ok", float(v.float().abs().max()))
