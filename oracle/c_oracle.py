"""Builder and ctypes binding of ``oracle/heom_oracle.c`` (C/OpenMP restatement).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  The library is built by
``gcc`` into ``oracle/_build/libheom_oracle.so`` (git-ignored; it travels to the
GPU box with the snapshot so ``bench.py`` never needs a compiler there, though
it rebuilds if the file is missing and ``gcc`` is present).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "heom_oracle.c")
OUT = os.path.join(HERE, "_build", "libheom_oracle.so")

STAMP = OUT + ".host"

_lib = None


def _host_signature() -> str:
    """The library is compiled with -march=native; a copy built on another CPU
    (the snapshot travels from the build container to the GPU box) is rebuilt."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    sig = _host_signature()
    try:
        with open(STAMP) as f:
            same_host = f.read().strip() == sig
    except OSError:
        same_host = False
    if (not force and same_host and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fcx-limited-range", "-std=c99", "-shared", "-fPIC",
           SRC, "-o", OUT, "-lm"]
    try:
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    except subprocess.CalledProcessError:
        # -march=native can be refused on exotic hosts; retry portable
        cmd.remove("-march=native")
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    with open(STAMP, "w") as f:
        f.write(sig)
    return OUT


def load():
    global _lib
    if _lib is None:
        try:
            path = build()
        except (OSError, subprocess.CalledProcessError):
            if not os.path.exists(OUT):
                raise
            path = OUT
        lib = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        lib.oracle_c_nmax.restype = C.c_longlong
        lib.oracle_c_nmax.argtypes = [C.c_int, C.c_int]
        lib.oracle_c_keys.restype = C.c_int
        lib.oracle_c_keys.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.oracle_c_max_threads.restype = C.c_int
        lib.oracle_c_deom_rk4.restype = C.c_int
        lib.oracle_c_deom_rk4.argtypes = [C.c_int] * 4 + [dp] * 8 + [
            C.POINTER(C.c_int64), dp, C.c_double, C.c_longlong, dp, dp, dp, C.c_int]
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(load().oracle_c_max_threads())


def keys(nind: int, lmax: int) -> np.ndarray:
    lib = load()
    out = np.zeros((int(lib.oracle_c_nmax(nind, lmax)), nind), dtype=np.int32)
    lib.oracle_c_keys(nind, lmax, out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def _c128(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.complex128))
    return a if shape is None else a.reshape(shape)


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def run(system, system_dipole, coupling, coupling_dipole, expn, etal, etar, etaa, mode, lmax,
        rho0, dt, nt, pulse_system=None, pulse_coupling=None, threads: int = 0, ados=None):
    """RK4 trajectory of rho_sys; same arguments as ``DeomOracle`` + ``run``.

    ``pulse_*`` are callables of time or ``None``.  Returns ``(traj, ados)`` with
    ``traj`` of shape ``(nt + 1, N, N)`` and ``ados`` in reference id order.
    """
    lib = load()
    H = _c128(system)
    n = H.shape[0]
    Q = _c128(coupling, (-1, n, n))
    m = Q.shape[0]
    mu = _c128(system_dipole if system_dipole is not None else np.zeros_like(H))
    Qd = _c128(coupling_dipole if coupling_dipole is not None else np.zeros_like(Q), (m, n, n))
    expn, etal, etar, etaa = (_c128(x) for x in (expn, etal, etar, etaa))
    k = expn.shape[0]
    mode = np.ascontiguousarray(np.asarray(mode, dtype=np.int64))
    nmax = int(lib.oracle_c_nmax(k, lmax))
    if ados is None:
        y = np.zeros((nmax, n, n), dtype=np.complex128)
        y[0] = rho0
    else:
        y = _c128(ados, (nmax, n, n)).copy()
    traj = np.zeros((nt + 1, n, n), dtype=np.complex128)

    def table(fn):
        if fn is None:
            return None
        t = np.empty((nt, 3), dtype=np.float64)
        for i in range(nt):
            t[i] = (fn(i * dt), fn(i * dt + dt / 2), fn(i * dt + dt))
        return t
    fs, fc = table(pulse_system), table(pulse_coupling)
    rc = lib.oracle_c_deom_rk4(n, k, m, int(lmax), _dp(H.view(np.float64)), _dp(mu.view(np.float64)),
                               _dp(Q.view(np.float64)), _dp(Qd.view(np.float64)),
                               _dp(expn.view(np.float64)), _dp(etal.view(np.float64)),
                               _dp(etar.view(np.float64)), _dp(etaa.view(np.float64)),
                               mode.ctypes.data_as(C.POINTER(C.c_int64)), _dp(y.view(np.float64)),
                               float(dt), int(nt), _dp(fs), _dp(fc), _dp(traj.view(np.float64)),
                               int(threads))
    if rc != 0:
        raise MemoryError("oracle_c_deom_rk4 failed to allocate")
    return traj, y
