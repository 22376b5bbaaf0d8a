"""CPU oracle for the GFE-Mamba selective-scan hot path -- numpy front end.

TEST INFRASTRUCTURE ONLY.  Allowed importers: ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product
(``gfe_mamba_b200/``, ``cross_atten/``) never imports this module and has no CPU
fallback.

Parity status: PINNED against fixtures generated from the unmodified reference
(``tests/golden/make_golden.py``; checked by ``tests/test_oracle_golden.py``).

The heavy loops live in ``gfe_oracle.c`` (built by ``oracle/Makefile`` into
``libgfe_oracle.so``); this file only marshals numpy arrays and restates the
thin torch glue of ``cross_atten/mamba.py`` (projections, RMSNorm, gating).
Every function cites the reference lines it follows (paths relative to the
reference checkout).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgfe_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)


def build(force: bool = False) -> str:
    """Compile ``gfe_oracle.c`` -> ``libgfe_oracle.so`` (idempotent)."""
    src = os.path.join(_HERE, "gfe_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B" if force else "-s"], check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_npo2.restype = ctypes.c_int
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def _c(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: Optional[np.ndarray]):
    if a is None:
        return _f32p()
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


def _chk(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed rc={rc}")


# --------------------------------------------------------------------------- pscan
def npo2(length: int) -> int:
    """cross_atten/pscan.py:13-18."""
    return int(lib().orc_npo2(ctypes.c_int(int(length))))


def pscan_fwd(A: np.ndarray, X: np.ndarray) -> np.ndarray:
    """PScan.forward, cross_atten/pscan.py:152-186.  A, X: (B, L, D, N) -> H."""
    A, X = _c(A), _c(X)
    B, L, D, N = A.shape
    H = np.empty_like(X)
    _chk(lib().orc_pscan_fwd(_p(A), _p(X), _p(H), B, L, D, N), "pscan_fwd")
    return H


def pscan_bwd(A: np.ndarray, H: np.ndarray, dH: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """PScan.backward, cross_atten/pscan.py:189-224 -> (dA, dX)."""
    A, H, dH = _c(A), _c(H), _c(dH)
    B, L, D, N = A.shape
    dA, dX = np.empty_like(A), np.empty_like(A)
    _chk(lib().orc_pscan_bwd(_p(A), _p(H), _p(dH), _p(dA), _p(dX), B, L, D, N), "pscan_bwd")
    return dA, dX


def pscan_seq64(A: np.ndarray, X: np.ndarray) -> np.ndarray:
    """fp64 sequential H[t] = A[t] H[t-1] + X[t] (truth for tolerance checks)."""
    A64, X64 = A.astype(np.float64), X.astype(np.float64)
    H = np.empty_like(X64)
    h = np.zeros_like(X64[:, 0])
    for t in range(A.shape[1]):
        h = A64[:, t] * h + X64[:, t]
        H[:, t] = h
    return H


# ------------------------------------------------------------------ selective scan
def selscan_ref_fwd(x, delta, A, Bm, Cm, D) -> np.ndarray:
    """MambaBlock.selective_scan, cross_atten/mamba.py:265-286 (materialising, Blelloch pscan)."""
    x, delta, A, Bm, Cm, D = map(_c, (x, delta, A, Bm, Cm, D))
    Bsz, L, ED = x.shape
    N = A.shape[1]
    y = np.empty_like(x)
    _chk(lib().orc_selscan_ref_fwd(_p(x), _p(delta), _p(A), _p(Bm), _p(Cm), _p(D), _p(y), _f32p(), _f32p(),
                                   Bsz, L, ED, N), "selscan_ref_fwd")
    return y


def selscan_ref_bwd(x, delta, A, Bm, Cm, D, dy) -> Dict[str, np.ndarray]:
    """Autograd of selective_scan (mamba.py:265-286) with PScan.backward (pscan.py:189-224)."""
    x, delta, A, Bm, Cm, D, dy = map(_c, (x, delta, A, Bm, Cm, D, dy))
    Bsz, L, ED = x.shape
    N = A.shape[1]
    out = dict(dx=np.empty_like(x), ddelta=np.empty_like(x), dA=np.empty_like(A),
               dB=np.empty_like(Bm), dC=np.empty_like(Cm), dD=np.empty_like(D))
    _chk(lib().orc_selscan_ref_bwd(_p(x), _p(delta), _p(A), _p(Bm), _p(Cm), _p(D), _p(dy),
                                   _p(out["dx"]), _p(out["ddelta"]), _p(out["dA"]), _p(out["dB"]),
                                   _p(out["dC"]), _p(out["dD"]), Bsz, L, ED, N), "selscan_ref_bwd")
    return out


def selscan_seq_fwd(u, delta_raw, A_log, Bm, Cm, D, z=None, dt_bias=None, softplus=True,
                    return_last_state=False):
    """Fused form of mamba.py:232,255-256 (softplus+bias), :288-318 (sequential scan), :220-222 (gate)."""
    u, delta_raw, A_log, Bm, Cm, D = map(_c, (u, delta_raw, A_log, Bm, Cm, D))
    z = None if z is None else _c(z)
    dt_bias = None if dt_bias is None else _c(dt_bias)
    Bsz, L, ED = u.shape
    N = A_log.shape[1]
    out = np.empty_like(u)
    hl = np.empty((Bsz, ED, N), np.float32) if return_last_state else None
    _chk(lib().orc_selscan_seq_fwd(_p(u), _p(delta_raw), _p(z), _p(A_log), _p(Bm), _p(Cm), _p(D), _p(dt_bias),
                                   _p(out), _p(hl), Bsz, L, ED, N, int(bool(softplus))), "selscan_seq_fwd")
    return (out, hl) if return_last_state else out


def selscan_seq_bwd(u, delta_raw, A_log, Bm, Cm, D, dout, z=None, dt_bias=None, softplus=True) -> Dict[str, np.ndarray]:
    """Closed-form backward of the fused form, fp64 accumulation."""
    u, delta_raw, A_log, Bm, Cm, D, dout = map(_c, (u, delta_raw, A_log, Bm, Cm, D, dout))
    z = None if z is None else _c(z)
    dt_bias = None if dt_bias is None else _c(dt_bias)
    Bsz, L, ED = u.shape
    N = A_log.shape[1]
    g = dict(du=np.empty_like(u), ddelta=np.empty_like(u), dz=None if z is None else np.empty_like(u),
             dB=np.empty_like(Bm), dC=np.empty_like(Cm), dA_log=np.empty_like(A_log), dD=np.empty_like(D),
             ddt_bias=None if dt_bias is None else np.empty_like(D))
    _chk(lib().orc_selscan_seq_bwd(_p(u), _p(delta_raw), _p(z), _p(A_log), _p(Bm), _p(Cm), _p(D), _p(dt_bias),
                                   _p(dout), _p(g["du"]), _p(g["ddelta"]), _p(g["dz"]), _p(g["dB"]), _p(g["dC"]),
                                   _p(g["dA_log"]), _p(g["dD"]), _p(g["ddt_bias"]),
                                   Bsz, L, ED, N, int(bool(softplus))), "selscan_seq_bwd")
    return g


# --------------------------------------------------------------------- conv + step
def conv1d_silu_fwd(xin: np.ndarray, w: np.ndarray, bias: Optional[np.ndarray]) -> np.ndarray:
    """nn.Conv1d(groups=ED, k=K, padding=K-1)(x^T)[:, :, :L]^T then F.silu, mamba.py:128-131,208-212.
    xin: (B, L, ED) possibly a strided half of xz; w: (ED, 1, K) or (ED, K)."""
    Bsz, L, ED = xin.shape
    w2 = _c(w.reshape(ED, -1))
    K = w2.shape[1]
    xin = _c(xin)
    u = np.empty((Bsz, L, ED), np.float32)
    _chk(lib().orc_conv1d_silu_fwd(_p(xin), ctypes.c_long(ED), _p(w2), _p(None if bias is None else _c(bias)),
                                   _p(u), Bsz, L, ED, K), "conv1d_silu_fwd")
    return u


def conv1d_silu_bwd(xin, w, bias, du):
    Bsz, L, ED = xin.shape
    w2 = _c(w.reshape(ED, -1))
    K = w2.shape[1]
    xin, du = _c(xin), _c(du)
    dxin = np.empty_like(xin)
    dw = np.empty_like(w2)
    db = None if bias is None else np.empty(ED, np.float32)
    _chk(lib().orc_conv1d_silu_bwd(_p(xin), ctypes.c_long(ED), _p(w2), _p(None if bias is None else _c(bias)),
                                   _p(du), _p(dxin), _p(dw), _p(db), Bsz, L, ED, K), "conv1d_silu_bwd")
    return dxin, dw.reshape(w.shape), db


def ssm_step(x, delta, A_log, Bm, Cm, D, h):
    """Recurrence of MambaBlock.ssm_step, mamba.py:391-403.  Returns (y, h_new)."""
    x, delta, A_log, Bm, Cm, D = map(_c, (x, delta, A_log, Bm, Cm, D))
    Bsz, ED = x.shape
    N = A_log.shape[1]
    h = np.zeros((Bsz, ED, N), np.float32) if h is None else _c(h).copy()   # mamba.py:396-397
    y = np.empty_like(x)
    _chk(lib().orc_ssm_step(_p(x), _p(delta), _p(A_log), _p(Bm), _p(Cm), _p(D), _p(h), _p(y), Bsz, ED, N), "ssm_step")
    return y, h


# ------------------------------------------------------------ block-level glue (numpy)
def _softplus(v: np.ndarray) -> np.ndarray:
    v64 = v.astype(np.float64)
    return np.where(v64 > 20.0, v64, np.log1p(np.exp(np.minimum(v64, 20.0)))).astype(np.float32)


def _silu(v: np.ndarray) -> np.ndarray:
    v64 = v.astype(np.float64)
    return (v64 / (1.0 + np.exp(-v64))).astype(np.float32)


def rmsnorm(x: np.ndarray, w: np.ndarray, eps: float) -> np.ndarray:
    """RMSNorm.forward, mamba.py:415-418."""
    x64 = x.astype(np.float64)
    return (x64 / np.sqrt((x64 ** 2).mean(-1, keepdims=True) + eps) * w).astype(np.float32)


def block_forward(p: Dict[str, np.ndarray], x: np.ndarray, d_state: int, dt_rank: int) -> np.ndarray:
    """MambaBlock.forward on the pscan path, mamba.py:197-225 + ssm :227-263.  ``p`` holds the
    state-dict entries of one ``mixer``.  bias=False / inner_layernorms=False."""
    ED = p["A_log"].shape[0]
    xz = x.astype(np.float32) @ p["in_proj.weight"].T                       # :204
    xin, z = xz[..., :ED], xz[..., ED:]                                     # :205
    u = conv1d_silu_fwd(xin, p["conv1d.weight"], p.get("conv1d.bias"))      # :208-212
    dbc = u @ p["x_proj.weight"].T                                          # :235
    dt, Bm, Cm = dbc[..., :dt_rank], dbc[..., dt_rank:dt_rank + d_state], dbc[..., dt_rank + d_state:]
    delta_raw = dt @ p["dt_proj.weight"].T                                  # :238 (bias added inside softplus, :256)
    y = selscan_seq_fwd(u, delta_raw, p["A_log"], Bm, Cm, p["D"], z=z, dt_bias=p["dt_proj.bias"], softplus=True)
    return y @ p["out_proj.weight"].T                                       # :223


def mamba_forward(sd: Dict[str, np.ndarray], x: np.ndarray, n_layers: int, d_state: int, dt_rank: int,
                  eps: float = 1e-5) -> np.ndarray:
    """Mamba.forward / ResidualBlock.forward, mamba.py:69-77,98-104."""
    for i in range(n_layers):
        pre = f"layers.{i}."
        p = {k[len(pre) + len("mixer."):]: v for k, v in sd.items() if k.startswith(pre + "mixer.")}
        x = block_forward(p, rmsnorm(x, sd[pre + "norm.weight"], eps), d_state, dt_rank) + x
    return x
