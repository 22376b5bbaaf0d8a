"""Build the sm_100a shared library (C ABI in include/gfe_mamba_b200.h) in-tree with nvcc.

    python -m gfe_mamba_b200.build [--force] [--verbose]

Output: gfe_mamba_b200/lib/libgfe_mamba_b200.so (git-ignored, travels to the GPU box with the snapshot).
nvcc cross-compiles without a GPU.  cudart is linked statically, so the library loads on a machine
without a driver (symbol checks in the CPU test-suite rely on that).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(PKG, "build")
LIB = os.path.join(LIBDIR, "libgfe_mamba_b200.so")
SOURCES = ["api.cu", "selscan.cu", "selscan_chain_host.cu", "selscan_v4_fwd.cu", "selscan_chain_bwd.cu", "selscan_seg.cu", "pscan.cu", "conv1d.cu", "step.cu", "addnorm.cu", "optim.cu", "pool.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the gfe_mamba_b200 CUDA library cannot be built")


def _host_cxx() -> list:
    # the image exports CC/CXX pointing at a toolchain without all runtime pieces; prefer the system g++
    for cand in ("/usr/bin/g++",):
        if os.path.exists(cand):
            return ["-ccbin", cand]
    return []


def _stale(target: str, deps: list) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defs: tuple = (), tag: str = "") -> str:
    """`defs`/`tag` build an A/B variant (extra -D flags) into lib/libgfe_mamba_b200.<tag>.so; development only."""
    global OBJDIR, LIB
    if tag:
        OBJDIR = os.path.join(PKG, "build", tag)
        LIB = os.path.join(LIBDIR, f"libgfe_mamba_b200.{tag}.so")
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "gfe_mamba_b200.h"))
    nvcc = _nvcc()
    ccbin = _host_cxx()

    def compile_one(src: str) -> str:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ccbin + NVCC_FLAGS + [f"-D{d}" for d in defs] + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            with open(o + ".log", "w") as f:
                f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ccbin + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    _defs = tuple(a[2:] for a in sys.argv if a.startswith("-D"))
    _tag = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--tag=")), "")
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defs=_defs, tag=_tag)
    print(path)
