"""Pinned host <-> device copy bandwidth, one direction at a time and both at once (the floor of bench.py's e2e leg).

Single process:  python tools/pcie_probe.py
N concurrent processes, one per GPU (what the N-GPU e2e leg does to the host):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py
Under torchrun every rank copies at the same time (barrier before each phase); rank 0 prints per-GPU min / mean over the ranks
and the aggregate, plus the host's NUMA layout -- the evidence for whether the e2e number is bound by the host."""
import os, time, torch
rank, world, lrank = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(lrank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
n = 512 << 20
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
def run(h2d, d2h, reps=4):
    sync(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9
run(True, True, 1)
res = torch.tensor([run(True, False), run(False, True), run(True, True)], device="cuda", dtype=torch.float64)
if world > 1:
    allr = [torch.empty_like(res) for _ in range(world)]
    dist.all_gather(allr, res)
    allr = torch.stack(allr)
else:
    allr = res[None]
if rank == 0:
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    except OSError:
        nodes = []
    names = ("H2D alone", "D2H alone", "both at once (per direction)")
    print(f"{world} concurrent process(es), one per GPU; host: {os.cpu_count()} CPUs, NUMA nodes: {len(nodes) or 'unknown'}")
    for i, nm in enumerate(names):
        col = allr[:, i]
        print(f"  {nm:30s} per GPU min {col.min():5.1f}  mean {col.mean():5.1f} GB/s   aggregate {col.sum():6.1f} GB/s")
if world > 1:
    dist.destroy_process_group()
