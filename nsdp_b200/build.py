"""Builds libnsdp_b200.so (the C-ABI library, include/nsdp_b200.h) in-tree with nvcc for sm_100a.

    python -m nsdp_b200.build [--force] [-v]

One object per .cu under nsdp_b200/csrc (compiled in parallel, cached by mtime), linked into
nsdp_b200/lib/libnsdp_b200.so. The .so is git-ignored but travels to the GPU box with the gpurun
snapshot. No torch headers are involved: the library is plain CUDA runtime + extern "C".
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
# NSDP_BUILD_VARIANT=<name> + NSDP_BUILD_DEFS="-DX -DY" builds an A/B copy (lib/libnsdp_b200_<name>.so) next to the product
VARIANT = os.environ.get("NSDP_BUILD_VARIANT", "")
OBJ = os.path.join(PKG, "lib", "obj" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(PKG, "lib", "libnsdp_b200" + ("_" + VARIANT if VARIANT else "") + ".so")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-Xptxas", "-v", "-I", INCLUDE] + (os.environ.get("NSDP_BUILD_DEFS", "").split() if VARIANT else [])


def _deps_mtime() -> float:
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src: str, force: bool, verbose: bool, hdr_mtime: float) -> str:
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src)
            and os.path.getmtime(obj) >= hdr_mtime):
        return obj
    cmd = [NVCC, *ARCH, *FLAGS, "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, os.path.basename(src)[:-3] + ".ptxas.log")
    with open(log, "w") as f:
        f.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdr_mtime = _deps_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose, hdr_mtime), srcs))
    if (force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
