"""Kernel 3 (``stage_rows_async_kernel``, the default stage kernel for N <= 8 with
diagonal coupling operators) on the CPU.

``pyqed_b200/csrc/heom_stage_async.cuh`` is compiled unchanged with g++ against
``tests/_shim/cuda_emu.h`` (see ``tests/test_sym_kernel_emu.py``) and its RK4
trajectory and final ADOs are compared with the oracle: the Hermitian-symmetric
path, the Hermitian row-fetch path, the general (non-Hermitian) path with its
out-of-line column loads, several-entry diagonal operators (sigma_z,
occupation numbers), real and complex H, every N, owned ranges and the rotated
visiting order - under both timing extremes of the asynchronous copies.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from pyqed_b200 import workloads as W
from test_sym_kernel_emu import host_tables, projector_problem, C128, ROOT


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu3") / "libasync_emu.so"
    src = os.path.join(ROOT, "tests", "_shim", "async_emu.cpp")
    subprocess.check_call(["g++", "-O0", "-shared", "-fPIC", "-std=c++17", "-pthread", "-x", "c++",
                           "-o", str(out), src])
    lib = ctypes.CDLL(str(out))
    lib.emu_async_run.restype = ctypes.c_int
    return lib


@pytest.fixture(params=["copies land at issue", "copies land at the wait"])
def emu(request, emu_lib):
    emu_lib.emu_set_async_late(ctypes.c_int(int(request.param.endswith("wait"))))
    return emu_lib


def run(emu, w, nt, sym, herm, sm_count=3, warps=2, parts=None, scramble=0):
    o, t = host_tables(w, single_support=False)
    N = t["N"]
    state = np.zeros((4, o.nmax, N, N), dtype=C128)
    state[0, 0] = w["rho0"]
    state[1:] = np.nan
    traj = np.zeros((nt + 1, N, N), dtype=C128)
    parts = np.array(parts if parts is not None else [0, o.nmax], dtype=np.int64)
    H = np.ascontiguousarray(o.H0)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = emu.emu_async_run(
        ctypes.c_int(N), ctypes.c_int(t["K"]), ctypes.c_int(t["M"]), ctypes.c_int(o.lmax),
        ctypes.c_longlong(o.nmax), p(H), p(t["ops"]), p(t["cbase"]), p(t["kmode"]), p(t["supp"]),
        p(t["damp"]), p(t["link_ptr"]), p(t["links"]), p(state), ctypes.c_double(w["dt"]),
        ctypes.c_int(nt), ctypes.c_int(int(np.all(H.imag == 0))), ctypes.c_int(sym), ctypes.c_int(herm),
        ctypes.c_int(sm_count), ctypes.c_int(warps), p(parts), ctypes.c_int(len(parts) // 2),
        ctypes.c_longlong(0), ctypes.c_int(scramble), p(traj))
    assert rc == 0
    _, ref = o.run(w["rho0"], w["dt"], nt)
    scale = max(1.0, float(np.abs(o.ddos).max()))
    assert np.isfinite(state[0]).all()
    assert np.abs(traj - np.array(ref)).max() < 1e-12
    assert np.abs(state[0] - o.ddos).max() < 1e-12 * scale


def general_problem(n, lmax, seed, herm, multi):
    """Diagonal couplings with one or several non-zero entries; Hermitian-preserving
    bath or a general one (complex exponents, unrelated etal / etar, any rho0)."""
    rng = np.random.default_rng(seed)
    w = projector_problem(n, 1, lmax, seed, complex_h=bool(seed & 1))
    if multi:
        Q = np.zeros_like(w["coupling"])
        for m in range(Q.shape[0]):
            k = int(rng.integers(2, min(3, n) + 1))
            idx = rng.choice(n, size=k, replace=False)
            Q[m, idx, idx] = rng.uniform(0.5, 1.5, k) * rng.choice([-1, 1], k)
        w["coupling"] = Q
    if not herm:
        K = len(w["expn"])
        w["expn"] = (rng.uniform(0.5, 2.0, K) + 1j * rng.uniform(-1, 1, K)).astype(C128)
        w["etal"] = ((rng.normal(size=K) + 1j * rng.normal(size=K)) * 0.3).astype(C128)
        w["etar"] = ((rng.normal(size=K) + 1j * rng.normal(size=K)) * 0.3).astype(C128)
        w["etaa"] = rng.uniform(0.1, 0.5, K).astype(C128)
        r = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        w["rho0"] = (r / np.trace(r)).astype(C128)
    return w


def test_headline_shape_symmetric_path(emu):
    """N=7, projector couplings, real H: the instantiation bench.py times."""
    run(emu, W.fmo(lmax=3, n_matsubara=0), nt=2, sym=1, herm=1)


def test_hermitian_row_fetch_without_symmetric_shortcuts(emu):
    run(emu, W.fmo(lmax=2, n_matsubara=2), nt=2, sym=0, herm=1, sm_count=2, warps=3)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 8])
def test_every_system_size_symmetric(emu, n):
    run(emu, projector_problem(n, 1, 3, seed=n, complex_h=bool(n & 1)), nt=2, sym=1, herm=1)


@pytest.mark.parametrize("n,herm,multi", [(2, 1, True), (3, 0, False), (4, 0, True), (5, 1, True),
                                          (7, 0, True), (8, 0, False)])
def test_general_paths(emu, n, herm, multi):
    """sigma_z-like / occupation-like operators (the multi-row branch) and problems
    whose ADOs are not Hermitian (column entries loaded out of line)."""
    run(emu, general_problem(n, 3, seed=10 + n, herm=bool(herm), multi=multi), nt=2, sym=0, herm=herm)


def test_owned_ranges_and_rotation(emu):
    run(emu, W.fmo(lmax=3, n_matsubara=0), nt=2, sym=1, herm=1, parts=[0, 57, 57, 120], scramble=1,
        sm_count=2, warps=1)
