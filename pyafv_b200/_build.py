"""Compile the CUDA library in-tree: pyafv_b200/libafv_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc" / "afv_api.cu"
DEPS = [SRC, *sorted((PKG / "csrc").glob("*.cuh")), *sorted((PKG / "csrc").glob("*.inc")),
        PKG.parent / "include" / "afv_b200.h"]
LIB = PKG / "libafv_b200.so"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-cudart", "shared"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the AFV CUDA library cannot be built")


def up_to_date() -> bool:
    return LIB.exists() and all(LIB.stat().st_mtime >= d.stat().st_mtime for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if up_to_date() and not force:
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(LIB), str(SRC)]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
