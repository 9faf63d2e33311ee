"""CPU-only checks: C-ABI surface, hierarchy indexing core, host-side helpers."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden
from oracle import deom_oracle as DO


# ---------------------------------------------------------------------------
# C ABI: the library loads and exports everything include/pyqed_heom.h declares
# ---------------------------------------------------------------------------
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pyqed_heom.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pyqed_heom_[a-z0-9_]+)\s*\(", text)))


def test_cabi_exports_every_declared_symbol():
    from pyqed_b200 import _cabi
    lib = _cabi.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
        assert name in _cabi.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_cabi.SIGNATURES) == set(names)
    assert lib.pyqed_heom_version() == 1


def test_hierarchy_size_matches_oracle():
    from pyqed_b200 import _cabi
    lib = _cabi.load()
    for K, L in [(1, 3), (2, 10), (3, 10), (7, 4), (21, 8), (14, 8), (4, 6), (6, 6)]:
        tab = DO.pascal_table(K, L)
        assert lib.pyqed_heom_hierarchy_size(K, L) == int(tab[L + K, L])
    assert lib.pyqed_heom_hierarchy_size(0, 3) == -1
    assert lib.pyqed_heom_hierarchy_size(40, 40) == -1  # >= 2^31 ADOs


def test_no_silent_cpu_fallback():
    """Without a CUDA device the product path must raise, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pyqed_b200 import _cabi
    from pyqed_b200.heom import DEOMSolver, Bath
    with pytest.raises(_cabi.HeomError):
        _cabi.Plan(2, 1, 1, 2)
    g = golden("deom_random3_K1")
    s = DEOMSolver(g["system"], g["system_dipole"],
                   Bath(expn=g["expn"], etal=g["etal"], etar=g["etar"], etaa=g["etaa"], mode=g["mode"]),
                   g["coupling"], g["coupling_dipole"], lmax=int(g["lmax"]))
    with pytest.raises(_cabi.HeomError):
        s.run(g["rho0"].copy(), 0.01, 2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pyqed_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f


# ---------------------------------------------------------------------------
# indexing core shared with the device code (compiled for the host with g++)
# ---------------------------------------------------------------------------
@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("shim") / "libcore_shim.so"
    src = os.path.join(ROOT, "tests", "_shim", "core_shim.cpp")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", str(out), src])
    lib = ctypes.CDLL(str(out))
    lib.shim_rank.restype = ctypes.c_longlong
    lib.shim_rank.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.shim_unrank_all.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p]
    return lib


@pytest.mark.parametrize("K,L", [(1, 6), (2, 10), (3, 10), (7, 4), (21, 2), (21, 3), (4, 6), (6, 6)])
def test_reference_order_unrank_matches_reference_keys(shim, K, L):
    tab = DO.pascal_table(K, L)
    ref = DO.build_keys(K, L, tab)
    keys = np.zeros((len(ref), K), dtype=np.uint8)
    shim.shim_unrank_all(0, len(ref), K, L, keys.ctypes.data)
    assert np.array_equal(keys.astype(np.int64), ref)
    for n in range(0, len(ref), max(1, len(ref) // 97)):
        k = np.ascontiguousarray(keys[n])
        assert shim.shim_rank(0, k.ctypes.data, K, L) == n


def test_reference_keys_equal_golden_keys(shim):
    g = golden("deom_fmo_K21_L3")
    K, L = g["keys"].shape[1], int(g["lmax"])
    keys = np.zeros_like(g["keys"])
    shim.shim_unrank_all(0, len(keys), K, L, keys.ctypes.data)
    assert np.array_equal(keys, g["keys"])


@pytest.mark.parametrize("K,L", [(1, 6), (2, 10), (3, 7), (7, 4), (21, 2), (5, 5)])
def test_lexicographic_order_is_a_sorted_bijection(shim, K, L):
    nmax = int(DO.pascal_table(K, L)[L + K, L])
    keys = np.zeros((nmax, K), dtype=np.uint8)
    shim.shim_unrank_all(1, nmax, K, L, keys.ctypes.data)
    assert keys.sum(axis=1).max() == L
    as_tuples = [tuple(r) for r in keys]
    assert as_tuples == sorted(as_tuples) and len(set(as_tuples)) == nmax
    for n in range(0, nmax, max(1, nmax // 97)):
        k = np.ascontiguousarray(keys[n])
        assert shim.shim_rank(1, k.ctypes.data, K, L) == n


# ---------------------------------------------------------------------------
# host helpers of the drop-in classes
# ---------------------------------------------------------------------------
def test_pulse_sampling_uses_reference_stage_times():
    from pyqed_b200.heom.deom import sample_pulse
    dt, nt = 0.01, 7
    seen = []
    tab = sample_pulse(lambda t: seen.append(t) or np.sin(3 * t) + 2, dt, nt)
    assert tab.shape == (nt, 3)
    expect = []
    for i in range(nt):
        expect += [i * dt, i * dt + dt / 2, i * dt + dt]
    assert seen == expect  # bit-identical arguments
    assert sample_pulse(lambda t: 0, dt, nt) is None and sample_pulse(None, dt, nt) is None


def test_solver_argument_checks():
    from pyqed_b200.heom import DEOMSolver, Bath, HEOMSolver
    b = Bath(expn=[1.0], etal=[1 - 1j], mode=[0])
    assert np.allclose(b.etar, [1 + 1j]) and np.allclose(b.etaa, [abs(1 - 1j)])
    with pytest.raises(ValueError):
        Bath(expn=[1.0, 2.0], etal=[1.0, 1.0], mode=[0])
    s = DEOMSolver(bath=b, coupling=np.eye(2)[None], lmax=2)
    with pytest.raises(ValueError):
        s.check_()  # system missing
    s = DEOMSolver(system=np.eye(2), bath=b, lmax=2)
    with pytest.raises(ValueError):
        s.check_()  # coupling missing
    s = DEOMSolver(system=np.eye(2), bath=Bath(expn=[1.0], etal=[1.0], mode=[1]),
                   coupling=np.eye(2)[None], lmax=2)
    with pytest.raises(ValueError):
        s.check_()  # mode refers to a missing coupling operator
    s = DEOMSolver(system=np.eye(2), bath=b, coupling=np.eye(2)[None], lmax=3)
    s.check_()
    s.init_()
    assert (s.nsys, s.nind, s.nmod, s.nmax) == (2, 1, 1, 4)
    assert np.array_equal(s.comb_list, DO.pascal_table(1, 3))
    with pytest.raises(ValueError):
        HEOMSolver(np.eye(2), [np.eye(2)], [np.eye(2)], verbose=False).run(np.eye(2) / 2, 0.1, 1, 1.0, 1.0, 0.1, 1)


def test_workload_shapes():
    from pyqed_b200 import workloads as W
    from math import comb
    for w, n, k, m in [(W.spin_boson(), 2, 2, 1), (W.fmo(4, 0), 7, 7, 7), (W.fmo(8, 2), 7, 21, 7),
                       (W.polariton(), 32, 4, 2), (W.aggregate_2des(), 7, 6, 3)]:
        assert w["system"].shape == (n, n) and len(w["expn"]) == k and w["coupling"].shape == (m, n, n)
        assert np.allclose(w["system"], w["system"].conj().T)
        assert abs(np.trace(w["rho0"]) - 1) < 1e-15
        assert w["mode"].max() == m - 1
    assert comb(8 + 21, 8) == 4292145
