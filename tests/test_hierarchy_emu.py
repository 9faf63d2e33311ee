"""The hierarchy builder kernels (``pyqed_b200/csrc/heom_hierarchy.cuh``) on the CPU.

The device code that ``pyqed_heom_build_hierarchy`` launches - storage-order
permutation, keys, damping rates, link counts and CSR links - is compiled with
g++ against ``tests/_shim/cuda_emu.h`` and checked against the oracle's index
tables for all three storage orders; the tables it builds (blocked
lexicographic order, as sharded runs use) then drive the emulated kernels 6 and
7, which must still reproduce the oracle's trajectory.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import deom_oracle as DO
from pyqed_b200 import workloads as W
from pyqed_b200.heom.sharded import _lex_rank
from test_sym_kernel_emu import ROOT, C128, link_meta, host_tables


@pytest.fixture(scope="module")
def hier(tmp_path_factory):
    out = tmp_path_factory.mktemp("hier") / "libhier_emu.so"
    subprocess.check_call(["g++", "-O0", "-shared", "-fPIC", "-std=c++17", "-pthread", "-x", "c++", "-o", str(out),
                           os.path.join(ROOT, "tests", "_shim", "hier_emu.cpp")])
    lib = ctypes.CDLL(str(out))
    lib.emu_build_hierarchy.restype = ctypes.c_int
    return lib


def device_tables(lib, o, order, r0_of_mode):
    K, L, nmax = o.nind, o.lmax, o.nmax
    side = K + L + 1
    pascal = np.ascontiguousarray(DO.pascal_table(K, L)[:side, :side], dtype=np.int64)
    tier = o.keys.sum(axis=1)
    nlinks = int((o.keys > 0).sum() + K * (tier < L).sum())
    mode = np.array([int(o.mode[k]) | (r0_of_mode[int(o.mode[k])] << 8) for k in range(K)], dtype=np.int32)
    t = dict(keys=np.zeros((nmax, K), np.uint8), id_of_slot=np.zeros(nmax, np.int32),
             slot_of_id=np.zeros(nmax, np.int32), damp=np.zeros(nmax, C128),
             link_ptr=np.zeros(nmax + 1, np.int32), links=np.zeros((nlinks, 2), np.int32),
             lex2slot=np.zeros(nmax, np.int32), slot2lex=np.zeros(nmax, np.int32))
    expn = np.ascontiguousarray(o.expn, dtype=C128)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.emu_build_hierarchy(ctypes.c_int(K), ctypes.c_int(L), ctypes.c_int(order), p(pascal),
                                 ctypes.c_int(side), ctypes.c_longlong(nmax), p(expn), p(mode), p(t["keys"]),
                                 p(t["id_of_slot"]), p(t["slot_of_id"]), p(t["damp"]), p(t["link_ptr"]),
                                 p(t["links"]), ctypes.c_longlong(nlinks), p(t["lex2slot"]), p(t["slot2lex"]))
    assert rc == 0
    return t


@pytest.mark.parametrize("K,L", [(2, 6), (3, 4), (7, 3), (14, 2)])
@pytest.mark.parametrize("order", [0, 1, 2])
def test_device_built_tables_match_the_oracle(hier, K, L, order):
    rng = np.random.default_rng(K * 10 + L)
    M = min(K, 3)
    w = dict(system=np.eye(3, dtype=C128), system_dipole=None, coupling=np.stack([np.eye(3, dtype=C128)] * M),
             coupling_dipole=None, expn=rng.uniform(0.5, 2, K), etal=np.ones(K), etar=np.ones(K), etaa=np.ones(K),
             mode=np.arange(K) % M, lmax=L)
    o = DO.DeomOracle(w["system"], None, w["coupling"], None, w["expn"], w["etal"], w["etar"], w["etaa"],
                      w["mode"], L)
    r0 = [m % 3 for m in range(M)]
    t = device_tables(hier, o, order, r0)
    nmax = o.nmax
    # the storage order is a bijection, and keys / damping are the oracle's, permuted
    assert sorted(t["id_of_slot"]) == list(range(nmax))
    assert np.array_equal(t["slot_of_id"][t["id_of_slot"]], np.arange(nmax))
    assert np.array_equal(t["keys"].astype(np.int64), o.keys[t["id_of_slot"]])
    assert np.allclose(t["damp"], (o.keys * o.expn[None, :]).sum(axis=1)[t["id_of_slot"]], rtol=0, atol=1e-14)
    lex = _lex_rank(o.keys, L)
    if order == 0:
        assert np.array_equal(t["id_of_slot"], np.arange(nmax))
    elif order == 1:
        assert np.array_equal(t["slot_of_id"], lex)
    else:
        # blocked lexicographic: inside every aligned run of 64 lexicographic ranks the ADOs below
        # the top tier come first, both parts in lexicographic order
        slot = t["slot_of_id"]
        assert np.array_equal(slot // 64, lex // 64)
        tier = o.keys.sum(axis=1)
        for b in range((nmax + 63) // 64):
            ids = np.nonzero(lex // 64 == b)[0]
            ids = ids[np.argsort(lex[ids])]
            expect = np.concatenate([ids[tier[ids] < L], ids[tier[ids] == L]])
            assert np.array_equal(t["id_of_slot"][64 * b:64 * b + len(ids)], expect)
    # links in the reference's summation order: k ascending, n - e_k before n + e_k
    for s in range(nmax):
        n = int(t["id_of_slot"][s])
        expect = []
        for k in range(K):
            m = int(o.mode[k])
            if o.minus[n, k] >= 0:
                expect.append((int(t["slot_of_id"][o.minus[n, k]]), link_meta(0, k, int(o.keys[n, k]), m, r0[m])))
            if o.plus[n, k] >= 0:
                expect.append((int(t["slot_of_id"][o.plus[n, k]]), link_meta(1, k, int(o.keys[n, k]) + 1, m, r0[m])))
        got = [tuple(int(x) for x in r) for r in t["links"][t["link_ptr"][s]:t["link_ptr"][s + 1]]]
        assert got == expect, s


def test_kernels_6_and_7_on_device_built_blocked_order(hier):
    """End to end in storage order 2 (what sharded runs use): tables from the emulated
    builder kernels, rotated visiting order, state and results permuted through
    ``slot_of_id`` as the C ABI does."""
    out = os.path.join(os.path.dirname(hier._name), "libsym_emu.so")
    subprocess.check_call(["g++", "-O0", "-shared", "-fPIC", "-std=c++17", "-pthread", "-x", "c++", "-o", out,
                           "-DHEOM_EMU_FEW_N", os.path.join(ROOT, "tests", "_shim", "sym_emu.cpp")])
    emu = ctypes.CDLL(out)
    emu.emu_sym_run.restype = ctypes.c_int
    emu.emu_packed_run.restype = ctypes.c_int
    w = W.fmo(lmax=3, n_matsubara=0)
    o, t0 = host_tables(w)
    N, K, M, nmax, nt = t0["N"], t0["K"], t0["M"], o.nmax, 2
    r0 = [int(k) >> 8 for k in t0["kmode"]]            # FMO: mode m couples through |m><m|
    t = device_tables(hier, o, 2, {int(o.mode[k]): r0[k] for k in range(K)})
    slot0 = int(t["slot_of_id"][0])
    H = np.ascontiguousarray(o.H0)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _, ref = o.run(w["rho0"], w["dt"], nt)
    for packed in (False, True):
        state = np.zeros((1 if packed else 4, nmax, N, N), dtype=C128)
        state[0, slot0] = w["rho0"]
        traj = np.zeros((nt + 1, N, N), dtype=C128)
        err = ctypes.c_char_p()
        common = [ctypes.c_int(N), ctypes.c_int(K), ctypes.c_int(M), ctypes.c_int(o.lmax), ctypes.c_longlong(nmax),
                  p(H), p(t0["ops"]), p(t0["cbase"]), p(t0["kmode"]), p(t["damp"]), p(t["link_ptr"]), p(t["links"]),
                  ctypes.c_longlong(len(t["links"])), p(state), ctypes.c_double(w["dt"]), ctypes.c_int(nt),
                  ctypes.c_int(1), ctypes.c_int(3), ctypes.c_int(2)]
        if packed:
            rc = emu.emu_packed_run(*common, ctypes.c_longlong(slot0), ctypes.c_int(1), p(traj), ctypes.c_int(1),
                                    ctypes.byref(err))
        else:
            parts = np.array([0, nmax], dtype=np.int64)
            rc = emu.emu_sym_run(*common, p(parts), ctypes.c_int(1), ctypes.c_longlong(slot0), ctypes.c_int(1),
                                 p(traj), ctypes.byref(err))
        assert rc == 0, err.value
        assert np.abs(traj - np.array(ref)).max() < 1e-12
        assert np.abs(state[0][t["slot_of_id"]] - o.ddos).max() < 1e-12
