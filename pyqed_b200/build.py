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
SOURCES = [os.path.join(CSRC, "heom_kernels.cu"), os.path.join(CSRC, "heom_stage_sym.cu"),
           os.path.join(CSRC, "heom_inst.cu"), os.path.join(CSRC, "heom_shard.cu")]
SRC = SOURCES[0]
SIZES = range(2, 9)   # system sizes N with templated kernels (one object file each)


def compile_units():
    """(source, object name, extra defines): the C ABI / builder unit, the common part of kernels
    6 / 7, and one unit per system size N for each of the two N-templated kernel families."""
    units = [(SOURCES[0], "heom_kernels.o", []), (SOURCES[1], "heom_stage_sym.o", ["HEOM_SYM_SPLIT"]),
             (SOURCES[3], "heom_shard.o", [])]
    for n in SIZES:
        units.append((SOURCES[1], f"heom_stage_sym_n{n}.o", [f"HEOM_SYM_INST_N={n}"]))
        units.append((SOURCES[2], f"heom_inst_n{n}.o", [f"HEOM_INST_N={n}"]))
    # the slowest units first so that the pool finishes evenly
    units.sort(key=lambda u: 0 if ("_n7" in u[1] or "_n8" in u[1] or "_n5" in u[1]) else 1)
    return units
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libpyqed_heom.so")
DEPS = SOURCES + [os.path.join(CSRC, h) for h in ("heom_core.cuh", "heom_device.cuh", "heom_plan.cuh", "heom_hierarchy.cuh", "heom_stage_rows.cuh", "heom_stage_async.cuh",
                                                 "heom_resident.cuh", "heom_stage_generic.cuh", "heom_stage_sym.cuh", "heom_plan.cuh", "heom_dataflow.cuh", "heom_dataflow_tma.cuh")] + \
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


# headers each source depends on (coarse, for incremental rebuilds during development)
_COMMON = ["heom_core.cuh", "heom_device.cuh"]
HEADER_DEPS = {
    "heom_stage_sym.cu": _COMMON + ["heom_stage_sym.cuh"],
    "heom_shard.cu": _COMMON + ["heom_plan.cuh", "heom_stage_sym.cuh", "../../include/pyqed_heom.h"],
    "heom_inst.cu": _COMMON + ["heom_plan.cuh", "heom_stage_sym.cuh", "heom_stage_async.cuh", "heom_stage_rows.cuh",
                               "heom_resident.cuh", "../../include/pyqed_heom.h"],
    "heom_kernels.cu": _COMMON + ["heom_plan.cuh", "heom_stage_sym.cuh", "heom_stage_async.cuh", "heom_hierarchy.cuh",
                                  "heom_resident.cuh", "heom_stage_generic.cuh", "heom_dataflow.cuh", "heom_dataflow_tma.cuh", "../../include/pyqed_heom.h"],
}


def _object_stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(CSRC, h) for h in HEADER_DEPS.get(os.path.basename(src), [])]
    return any(os.path.getmtime(d) > t for d in deps)


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
    units = compile_units()
    objs = [os.path.join(objdir, u[1]) for u in units]
    workers = max(1, min(len(units), os.cpu_count() or 1))
    # force / tuning variants rebuild everything; otherwise only the objects whose sources changed
    todo = [(u, o) for u, o in zip(units, objs) if force or defines or _object_stale(u[0], o)]
    with ThreadPoolExecutor(max_workers=workers) as pool:
        logs = list(pool.map(lambda uo: _run([nvcc] + flags + [f"-D{d}" for d in uo[0][2]] +
                                             ["-c", "-o", uo[1], uo[0][0]]), todo))
    _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", target] + objs)
    if verbose:
        print("\n".join(logs))
    return target


if __name__ == "__main__":
    import sys
    print(build_extension(force="-f" in sys.argv, verbose="-v" in sys.argv))
