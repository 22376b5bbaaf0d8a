"""Host-buffer entry point of the fused selective scan: pinned host tensors in, pinned host tensors out.

The scan itself moves ~22 B per (token, channel) through HBM in about a millisecond; a caller whose tensors live in host
memory is bound by PCIe instead (0.8 GB each way per cfg3 step).  ``HostScanPipeline`` hides as much of that as the
hardware allows: the batch is cut into row chunks and three streams run concurrently,

    copy-in stream   H2D of chunk i + 1          (PCIe, host -> device)
    compute stream   fused fwd + bwd of chunk i  (selscan kernels through the C ABI)
    copy-out stream  D2H of chunk i - 1          (PCIe, device -> host, full duplex with the H2D)

with double-buffered device inputs and event hand-offs, so the step costs max(H2D, D2H) plus one chunk of ramp instead of
H2D + compute + D2H.  Rows of the batch are independent recurrences (mamba.py:265-286), so chunking changes nothing in
out, du, ddelta, dz, dB, dC; the parameter gradients (dA_log, dD, ddt_bias) are summed over the chunks on the device.
There is no CPU fallback: a CUDA device and the native library are required.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .ops import selective_scan_fn

_IN_KEYS = ("u", "delta", "z", "Bm", "Cm", "dout")
_OUT_KEYS = ("out", "du", "ddelta", "dz", "dBm", "dCm")


def bind_host_to_gpu_node(device: torch.device) -> Optional[str]:
    """Pin this process to the CPUs that are local to ``device``'s PCIe root (sysfs ``local_cpulist``) so that pinned host
    buffers allocated afterwards are first-touched on the GPU's own NUMA node.  With one process per GPU this keeps eight
    ranks from streaming their 1.6 GB per step through one memory controller.  Returns the CPU list, or None when the
    topology cannot be read (nothing is changed then)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = torch.cuda.get_device_properties(device).pci_domain_id
        dev = torch.cuda.get_device_properties(device).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        with open(path) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return text
    except Exception:
        return None


def chunk_schedule(B: int, big: int):
    """Rows per chunk: 1, 1, 2, 4, ... up to ``big``, flat in the middle, mirrored at the end.  The first copy-in and the
    last copy-out are the only transfers nothing overlaps with, so they are kept one row long; the middle chunks are
    large enough for the kernels to fill the GPU."""
    up, r = [], 1
    while r < big and sum(up) + r <= B // 4:
        up.append(r)
        if len(up) >= 2:
            r *= 2
    rest = B - 2 * sum(up)
    mid = [big] * (rest // big) + ([rest % big] if rest % big else [])
    sched = up + mid + up[::-1]
    assert sum(sched) == B and all(0 < n <= big for n in sched), (B, big, sched)
    return sched


class HostScanPipeline:
    """Fused selective-scan forward + backward for HOST tensors of one fixed shape.

    host_in:  u, delta, z (optional), dout: (B, L, ED); Bm, Cm: (B, L, N)  -- pinned, activation dtype
    params:   A_log (ED, N), D (ED), dt_bias (ED) or None                 -- device tensors (weights stay resident)
    host_out: out, du, ddelta, dz: (B, L, ED); dBm, dCm: (B, L, N)        -- pinned, written in place
    returns   dA_log, dD, ddt_bias (device fp32 tensors, summed over the batch)
    """

    def __init__(self, B: int, L: int, ED: int, N: int, dtype: torch.dtype, device: torch.device,
                 rows_per_chunk: Optional[int] = None, has_z: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("gfe_mamba_b200: HostScanPipeline needs a CUDA device (no CPU fallback)")
        self.B, self.L, self.ED, self.N, self.dtype, self.dev, self.has_z = B, L, ED, N, dtype, torch.device(device), has_z
        self.rows = max(1, min(B, rows_per_chunk if rows_per_chunk else max(1, B // 8)))   # largest chunk (device buffers); B/8 measured best (cfg3: 21.4 ms vs 27.7 at B/4)
        self.schedule = chunk_schedule(B, self.rows)
        self.nchunks = len(self.schedule)
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        shp = {"u": ED, "delta": ED, "z": ED, "dout": ED, "Bm": N, "Cm": N}
        keys = [k for k in _IN_KEYS if has_z or k != "z"]
        self.dev_in = [{k: torch.empty((self.rows, L, shp[k]), dtype=dtype, device=self.dev) for k in keys} for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]      # inputs of the slot have landed
        self.ev_free = [torch.cuda.Event() for _ in range(2)]    # compute is done reading the slot's inputs
        self.h2d_bytes = sum(B * L * shp[k] * torch.empty((), dtype=dtype).element_size() for k in keys)
        outs = [k for k in _OUT_KEYS if has_z or k != "dz"]
        self.d2h_bytes = sum(B * L * (N if k in ("dBm", "dCm") else ED) * torch.empty((), dtype=dtype).element_size() for k in outs)

    @torch.no_grad()
    def run(self, host_in: Dict[str, torch.Tensor], A_log: torch.Tensor, D: torch.Tensor, dt_bias: Optional[torch.Tensor],
            host_out: Dict[str, torch.Tensor], delta_softplus: bool = True):
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_comp, self.s_out):
            s.wait_stream(cur)
        pacc = None
        keys = list(self.dev_in[0].keys())
        r1 = 0
        for i, n in enumerate(self.schedule):
            r0, r1 = r1, r1 + n
            slot = i & 1
            with torch.cuda.stream(self.s_in):
                if i >= 2:
                    self.s_in.wait_event(self.ev_free[slot])
                for k in keys:
                    self.dev_in[slot][k][:n].copy_(host_in[k][r0:r1], non_blocking=True)
                self.ev_in[slot].record(self.s_in)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(self.ev_in[slot])
                d = self.dev_in[slot]
                with torch.enable_grad():
                    leaves = {k: d[k][:n].detach().requires_grad_() for k in keys if k != "dout"}
                    pl = [A_log.detach().requires_grad_(), D.detach().requires_grad_()]
                    if dt_bias is not None:
                        pl.append(dt_bias.detach().requires_grad_())
                    out = selective_scan_fn(leaves["u"], leaves["delta"], pl[0], leaves["Bm"], leaves["Cm"], pl[1],
                                            z=leaves.get("z"), dt_bias=pl[2] if dt_bias is not None else None,
                                            delta_softplus=delta_softplus)
                    order = ["u", "delta"] + (["z"] if self.has_z else []) + ["Bm", "Cm"]
                    grads = torch.autograd.grad(out, [leaves[k] for k in order] + pl, d["dout"][:n])
                self.ev_free[slot].record(self.s_comp)
                res = {"out": out.detach()}
                for k, g in zip(order, grads):
                    res["d" + k] = g
                pg = [g.float() for g in grads[len(order):]]
                pacc = pg if pacc is None else [a + g for a, g in zip(pacc, pg)]
                ev_done = torch.cuda.Event()
                ev_done.record(self.s_comp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_done)
                for k, t in res.items():
                    t.record_stream(self.s_out)          # allocated on the compute stream, read by the copy-out stream
                    host_out[k][r0:r1].copy_(t, non_blocking=True)
        cur.wait_stream(self.s_comp)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_in)
        dA_log, dD = pacc[0], pacc[1]
        return dA_log, dD, (pacc[2] if dt_bias is not None else None)


def selective_scan_host(host_in: Dict[str, torch.Tensor], A_log: torch.Tensor, D: torch.Tensor,
                        dt_bias: Optional[torch.Tensor] = None, host_out: Optional[Dict[str, torch.Tensor]] = None,
                        rows_per_chunk: Optional[int] = None, delta_softplus: bool = True):
    """One-shot convenience wrapper: builds a pipeline for the shapes of ``host_in`` (pinning the tensors if needed), runs
    forward + backward and returns (host_out, dA_log, dD, ddt_bias) after synchronising."""
    u = host_in["u"]
    B, L, ED = u.shape
    N = host_in["Bm"].shape[-1]
    dev = A_log.device
    has_z = host_in.get("z") is not None
    hin = {k: (v if v.is_pinned() else v.pin_memory()) for k, v in host_in.items() if v is not None}
    if host_out is None:
        host_out = {k: torch.empty((B, L, N if k in ("dBm", "dCm") else ED), dtype=u.dtype).pin_memory()
                    for k in _OUT_KEYS if has_z or k != "dz"}
    pipe = HostScanPipeline(B, L, ED, N, u.dtype, dev, rows_per_chunk, has_z)
    dA_log, dD, dbias = pipe.run(hin, A_log, D, dt_bias, host_out, delta_softplus)
    torch.cuda.current_stream(dev).synchronize()
    return host_out, dA_log, dD, dbias
