"""Builds the CUDA extension in-tree with nvcc for sm_100a (no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "heom_kernels.cu")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libpyqed_heom.so")
DEPS = [SRC, os.path.join(HERE, "csrc", "heom_core.cuh"),
        os.path.join(os.path.dirname(HERE), "include", "pyqed_heom.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_extension(force: bool = False, verbose: bool = False, defines=None, out=None) -> str:
    """Compile ``csrc/heom_kernels.cu`` into ``lib/libpyqed_heom.so`` (or ``out``,
    with extra ``-D`` defines, for tuning variants)."""
    target = out or LIB
    if out is None and not force and not is_stale():
        return LIB
    os.makedirs(os.path.dirname(target), exist_ok=True)
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
           "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3"]
    cmd += [f"-D{d}" for d in (defines or [])]
    cmd += ["-o", target, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return target


if __name__ == "__main__":
    import sys
    print(build_extension(force=True, verbose="-v" in sys.argv))
