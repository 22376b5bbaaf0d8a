"""Pinned host <-> device copy bandwidth, one direction at a time and both at once (the floor of bench.py's e2e leg)."""
import time, torch
n = 512 << 20
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9
run(True, True, 1)
print(f"H2D alone {run(True, False):.1f} GB/s   D2H alone {run(False, True):.1f} GB/s   both at once {run(True, True):.1f} GB/s per direction")
