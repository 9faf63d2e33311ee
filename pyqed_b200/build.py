"""Builds the CUDA extension in-tree with nvcc for sm_100a (no JIT cache).

The translation units under ``csrc/`` are compiled to objects side by side (one
nvcc process each) and linked into ``lib/libpyqed_heom.so``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = [os.path.join(CSRC, "heom_kernels.cu"), os.path.join(CSRC, "heom_stage_sym.cu")]
SRC = SOURCES[0]
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libpyqed_heom.so")
DEPS = SOURCES + [os.path.join(CSRC, h) for h in ("heom_core.cuh", "heom_device.cuh", "heom_hierarchy.cuh", "heom_stage_rows.cuh", "heom_stage_async.cuh",
                                                 "heom_resident.cuh", "heom_stage_generic.cuh", "heom_stage_sym.cuh")] + \
    [os.path.join(os.path.dirname(HERE), "include", "pyqed_heom.h")]


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


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return res.stderr


def build_extension(force: bool = False, verbose: bool = False, defines=None, out=None) -> str:
    """Compile ``csrc/*.cu`` into ``lib/libpyqed_heom.so`` (or ``out``, with extra
    ``-D`` defines, for tuning variants)."""
    target = out or LIB
    if out is None and not force and not is_stale():
        return LIB
    os.makedirs(os.path.dirname(target), exist_ok=True)
    objdir = os.path.join(os.path.dirname(target), "obj_" + os.path.basename(target))
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
             "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3"]
    flags += [f"-D{d}" for d in (defines or [])]
    objs = [os.path.join(objdir, os.path.splitext(os.path.basename(s))[0] + ".o") for s in SOURCES]
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        logs = list(pool.map(lambda so: _run([nvcc] + flags + ["-c", "-o", so[1], so[0]]), zip(SOURCES, objs)))
    _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", target] + objs)
    if verbose:
        print("\n".join(logs))
    return target


if __name__ == "__main__":
    import sys
    print(build_extension(force=True, verbose="-v" in sys.argv))
