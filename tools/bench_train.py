"""Whole training step of a Mamba stack through gfe_mamba_b200.train (forward, backward, per-layer gradient all-reduce
overlapped with backward, multi-tensor clip + Adam) -- BASELINE configs[4] as specified: B=256 per GPU, L=1024,
d_model=512, n_layers=8, NCCL all-reduce of every parameter gradient (A_log, D, projections, conv, norms) at 1/2/4/8 GPUs.

    python tools/bench_train.py [--batch 256 --seq 1024 --d-model 512 --layers 8 --dtype bf16|tf32|f32 --steps 5 --warmup 3] [--graph]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_train.py ...

Prints one JSON line (rank 0): tokens/s over all ranks (device time, max over ranks), the step split into
forward+backward / optimiser, and the all-reduce time total vs exposed (step with the collectives minus step without).
--graph: single GPU, the whole step captured in a CUDA graph (production shape: --batch 2 --seq 1858 --layers 6 --dtype f32)."""
import argparse, json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gfe_mamba_b200 import Mamba, MambaConfig, _native
from gfe_mamba_b200.optim import ClipAdam
from gfe_mamba_b200.train import GraphedTrainStep, LayerGradSync, TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256); ap.add_argument("--seq", type=int, default=1024)
ap.add_argument("--d-model", type=int, default=512); ap.add_argument("--layers", type=int, default=8)
ap.add_argument("--dtype", default="bf16", choices=["bf16", "tf32", "f32"])
ap.add_argument("--steps", type=int, default=5); ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--graph", action="store_true"); ap.add_argument("--torch-optim", action="store_true")
a = ap.parse_args()
rank, world, lrank = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
real_stdout = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
torch.cuda.set_device(lrank); dev = torch.device("cuda", lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = a.dtype == "tf32"
torch.backends.cudnn.allow_tf32 = a.dtype == "tf32"
torch.manual_seed(0)                                   # identical replicas
model = Mamba(MambaConfig(d_model=a.d_model, n_layers=a.layers)).to(dev)
nparam = sum(p.numel() for p in model.parameters())
x = torch.randn(a.batch, a.seq, a.d_model, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
sync = LayerGradSync(model) if world > 1 else None
if a.torch_optim:   # the reference loop's own tail, for comparison (classify_mamba.py:104-109)
    class _RefOpt:
        param_groups = [{"zero_grad": False}]
        def __init__(s): s.o = torch.optim.Adam(model.parameters(), lr=1e-4); s.ps = list(model.parameters())
        def step(s):
            for p in s.ps: torch.nn.utils.clip_grad_norm_(p, max_norm=1.0)
            s.o.step()
        def zero_grad(s, set_to_none=False): s.o.zero_grad(set_to_none=set_to_none)
    opt = _RefOpt()
else:
    opt = ClipAdam(model.parameters(), lr=1e-4, max_norm=1.0, zero_grad=True)
step = TrainStep(model, opt, grad_sync=sync, autocast_dtype=torch.bfloat16 if a.dtype == "bf16" else None)
runner = GraphedTrainStep(step, x, warmup=a.warmup) if a.graph else step

def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize(dev)

def timed(fn, n):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); barrier()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

for _ in range(a.warmup): runner(x)
_native.timing_enable(True); _native.timing_collect()
ms_step = timed(lambda: runner(x), a.steps)
kern = _native.timing_collect(); _native.timing_enable(False)
lib_ms = {k: round(v[0] / a.steps, 3) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])}
res = {"op": "Mamba stack training step (fwd + bwd + grad all-reduce + clip + Adam)", "n_gpus": world, "batch_per_gpu": a.batch,
       "seq_len": a.seq, "d_model": a.d_model, "n_layers": a.layers, "dtype": a.dtype, "graph": a.graph, "params": nparam,
       "optimizer": "torch Adam + per-parameter clip_grad_norm_ loop" if a.torch_optim else "ClipAdam (3 launches)",
       "ms_per_step": round(ms_step, 3), "tokens_per_s": round(world * a.batch * a.seq / ms_step * 1e3),
       "library_kernels_ms": round(sum(lib_ms.values()), 3), "library_kernels_ms_per_step": lib_ms}
if world > 1:   # replicas must stay identical: same parameters on every rank after the synchronised steps (checked BEFORE the
    # measurement below that deliberately runs steps without the collectives)
    chk = torch.stack([p.detach().double().abs().sum() for p in model.parameters()]).sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res["replicas_identical"] = bool((hi - lo).abs().item() <= 1e-9 * max(1.0, abs(hi.item())))
if not a.graph:
    ms_fb = timed(lambda: (step.forward_backward(x), sync.zero() if sync else [p.grad.zero_() for p in model.parameters()]), a.steps)
    res["fwd_bwd_ms"] = round(ms_fb, 3)
if world > 1:
    sync.enabled = False                                # the same step without the collectives
    ms_nosync = timed(lambda: runner(x), a.steps)
    sync.enabled = True
    ms_ar = timed(sync.allreduce_now, 10)               # the collectives alone, nothing to hide under
    res.update({"allreduce_bytes": sum(sync.bucket_bytes), "allreduce_buckets": len(sync.bucket_bytes),
                "allreduce_total_ms": round(ms_ar, 3), "step_without_allreduce_ms": round(ms_nosync, 3),
                "allreduce_exposed_ms": round(max(ms_step - ms_nosync, 0.0), 3),
                "allreduce_hidden_ms": round(max(ms_ar - max(ms_step - ms_nosync, 0.0), 0.0), 3)})
if rank == 0:
    print(json.dumps(res), file=real_stdout, flush=True)
if world > 1:
    dist.destroy_process_group()
