"""Builds ``csrc/libdia_b200.so`` in-tree with nvcc for sm_100a (no JIT cache, no torch
cpp_extension: the library is a plain C ABI loaded through ctypes)."""

from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = CSRC / "libdia_b200.so"
SOURCES = ["step_kernel.cu", "aux_kernels.cu", "gemm_tcgen05.cu", "prefill_kernels.cu", "batch_kernel.cu", "engine.cu"]
HEADERS = ["common.cuh", "engine_internal.h", "sampler.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES + HEADERS] + [INCLUDE / "dia_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    extra = os.environ.get("DIA_NVCC_EXTRA", "").split()          # e.g. -DDIA_BATCH_TRACE for a bring-up build
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, f"-I{INCLUDE}", "-o", str(LIB), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
