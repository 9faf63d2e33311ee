"""Kernel 6 (``pyqed_b200/csrc/heom_stage_sym.cu``) on the CPU.

The kernel source is compiled unchanged with g++ against ``tests/_shim/cuda_emu.h``
(one OS thread per CUDA thread, real barriers for the warp/CTA synchronisation;
asynchronous copies land either when issued or only when waited for) and its RK4 trajectory and final ADOs
are compared with the oracle.  This checks the link-record format (``links2``),
the pre-scaled coefficient table, the compile-time stage kinds, the record
strip, tail groups, owned ranges and the visiting-order rotation without a GPU.
Cross-proxy fences are not modelled: that is what the GPU parity tests and
compute-sanitizer are for.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import deom_oracle as DO
from pyqed_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C128 = np.complex128


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu") / "libsym_emu.so"
    src = os.path.join(ROOT, "tests", "_shim", "sym_emu.cpp")
    subprocess.check_call(["g++", "-O0", "-shared", "-fPIC", "-std=c++17", "-pthread", "-x", "c++",
                           "-o", str(out), src])
    lib = ctypes.CDLL(str(out))
    lib.emu_sym_run.restype = ctypes.c_int
    return lib


@pytest.fixture(params=["copies land at issue", "copies land at the wait"])
def emu(request, emu_lib):
    """Both extremes of the legal timing of cp.async / cp.async.bulk (cuda_emu.h)."""
    emu_lib.emu_set_async_late(ctypes.c_int(int(request.param.endswith("wait"))))
    return emu_lib


def link_meta(direction, k, neff, mode, r0):
    """``heom::link_meta`` (pyqed_b200/csrc/heom_core.cuh)."""
    return neff | (direction << 8) | (k << 9) | ((r0 & 0xf) << 16) | (mode << 24)


def host_tables(w, single_support=True):
    """The tables ``pyqed_heom_build_hierarchy`` builds on the device, in the
    reference's id order (storage order 0), from the oracle's index tables."""
    o = DO.DeomOracle(w["system"], w["system_dipole"], w["coupling"], w["coupling_dipole"], w["expn"],
                      w["etal"], w["etar"], w["etaa"], w["mode"], w["lmax"])
    N, K, M = o.nsys, o.nind, o.Q0.shape[0]
    r0 = []
    # diagonal-Q support tables: [M][N+1] (count, rows) then [M][N] membership
    supp = np.zeros(M * (2 * N + 1), dtype=np.uint8)
    for m in range(M):
        nz = np.nonzero(np.diag(o.Q0[m]))[0]
        assert np.count_nonzero(o.Q0[m]) == len(nz) >= 1, "diagonal coupling operators only"
        if single_support:
            assert len(nz) == 1, "kernels 6 / 7 need one-entry diagonal Q_m"
        r0.append(int(nz[0]))
        supp[m * (N + 1)] = len(nz)
        supp[m * (N + 1) + 1:m * (N + 1) + 1 + len(nz)] = nz
        supp[M * (N + 1) + m * N + nz] = 1
    kmode = np.array([int(o.mode[k]) | (r0[int(o.mode[k])] << 8) for k in range(K)], dtype=np.int32)
    sa = np.sqrt(o.etaa)
    cbase = np.stack([-(1j / sa) * o.etal, (1j / sa) * o.etar, -1j * sa, 1j * sa], axis=1).astype(C128)
    damp = (o.keys * o.expn[None, :]).sum(axis=1).astype(C128)
    ptr = [0]
    recs = []
    for n in range(o.nmax):
        for k in range(K):
            m = int(o.mode[k])
            if o.minus[n, k] >= 0:
                recs.append((int(o.minus[n, k]), link_meta(0, k, int(o.keys[n, k]), m, r0[m])))
            if o.plus[n, k] >= 0:
                recs.append((int(o.plus[n, k]), link_meta(1, k, int(o.keys[n, k]) + 1, m, r0[m])))
        ptr.append(len(recs))
    links = np.array(recs, dtype=np.int32).reshape(-1, 2)
    ops = np.concatenate([o.H0[None], o.Q0]).astype(C128)
    return o, dict(N=N, K=K, M=M, kmode=kmode, supp=supp, cbase=np.ascontiguousarray(cbase), damp=damp,
                   link_ptr=np.array(ptr, dtype=np.int32), links=np.ascontiguousarray(links), ops=ops)


def run_emu(emu, w, nt, sm_count=3, warps=2, parts=None, scramble=0):
    o, t = host_tables(w)
    N, NN = t["N"], t["N"] ** 2
    state = np.zeros((4, o.nmax, N, N), dtype=C128)
    state[0, 0] = w["rho0"]
    # the kernel must never depend on what the stage buffers held before
    state[1:] = np.nan
    traj = np.zeros((nt + 1, N, N), dtype=C128)
    parts = np.array(parts if parts is not None else [0, o.nmax], dtype=np.int64)
    H = np.ascontiguousarray(o.H0)
    err = ctypes.c_char_p()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = emu.emu_sym_run(
        ctypes.c_int(N), ctypes.c_int(t["K"]), ctypes.c_int(t["M"]), ctypes.c_int(o.lmax),
        ctypes.c_longlong(o.nmax), p(H), p(t["ops"]), p(t["cbase"]), p(t["kmode"]), p(t["damp"]),
        p(t["link_ptr"]), p(t["links"]), ctypes.c_longlong(len(t["links"])), p(state),
        ctypes.c_double(w["dt"]), ctypes.c_int(nt), ctypes.c_int(int(np.all(H.imag == 0))),
        ctypes.c_int(sm_count), ctypes.c_int(warps), p(parts), ctypes.c_int(len(parts) // 2),
        ctypes.c_longlong(0), ctypes.c_int(scramble), p(traj), ctypes.byref(err))
    assert rc == 0, err.value
    _, ref = o.run(w["rho0"], w["dt"], nt)
    return state[0], traj, o.ddos, np.array(ref)


def run_emu_packed(emu, w, nt, sm_count=3, warps=2, scramble=0, prefetch=0):
    """Kernel 7: ``heom_packed_propagate`` (pack, nt steps on upper triangles, unpack)."""
    o, t = host_tables(w)
    N = t["N"]
    y = np.zeros((o.nmax, N, N), dtype=C128)
    y[0] = w["rho0"]
    traj = np.zeros((nt + 1, N, N), dtype=C128)
    H = np.ascontiguousarray(o.H0)
    err = ctypes.c_char_p()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    emu.emu_packed_run.restype = ctypes.c_int
    rc = emu.emu_packed_run(
        ctypes.c_int(N), ctypes.c_int(t["K"]), ctypes.c_int(t["M"]), ctypes.c_int(o.lmax),
        ctypes.c_longlong(o.nmax), p(H), p(t["ops"]), p(t["cbase"]), p(t["kmode"]), p(t["damp"]),
        p(t["link_ptr"]), p(t["links"]), ctypes.c_longlong(len(t["links"])), p(y),
        ctypes.c_double(w["dt"]), ctypes.c_int(nt), ctypes.c_int(int(np.all(H.imag == 0))),
        ctypes.c_int(sm_count), ctypes.c_int(warps), ctypes.c_longlong(0), ctypes.c_int(scramble),
        p(traj), ctypes.c_int(prefetch), ctypes.byref(err))
    assert rc == 0, err.value
    _, ref = o.run(w["rho0"], w["dt"], nt)
    return y, traj, o.ddos, np.array(ref)


def check(emu, w, nt, **kw):
    y, traj, ref_ados, ref_traj = run_emu(emu, w, nt, **kw)
    scale = max(1.0, float(np.abs(ref_ados).max()))
    assert np.isfinite(y).all()
    assert np.abs(traj - ref_traj).max() < 1e-12
    assert np.abs(y - ref_ados).max() < 1e-12 * scale
    # kernel 6 keeps every ADO Hermitian bit for bit
    assert np.array_equal(y, y.conj().transpose(0, 2, 1))
    if "parts" not in kw:   # kernel 7 (packed storage) owns the whole hierarchy
        kw.pop("parts", None)
        y7, traj7, _, _ = run_emu_packed(emu, w, nt, **kw)
        assert np.abs(traj7 - ref_traj).max() < 1e-12
        assert np.abs(y7 - ref_ados).max() < 1e-12 * scale
        assert np.array_equal(y7, y7.conj().transpose(0, 2, 1))
        # same arithmetic on the same values as kernel 6: identical bits
        assert np.array_equal(y7, y) and np.array_equal(traj7, traj)
        # ... and with the streamed tiles double-buffered and fetched one group ahead
        y7p, traj7p, _, _ = run_emu_packed(emu, w, nt, prefetch=1, **kw)
        assert np.array_equal(y7p, y) and np.array_equal(traj7p, traj)


def projector_problem(n, nind_per_mode, lmax, seed, complex_h):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n)) + (1j * rng.normal(size=(n, n)) if complex_h else 0)
    H = ((A + A.conj().T) / 2).astype(C128)
    M = n
    Q = np.zeros((M, n, n), C128)
    for m in range(M):
        Q[m, (m * 2 + 1) % n, (m * 2 + 1) % n] = 0.7 + 0.1 * m     # one diagonal entry, rows not in order
    K = M * nind_per_mode
    mode = np.repeat(np.arange(M), nind_per_mode)
    expn = (0.5 + rng.random(K)).astype(C128)
    etal = (rng.normal(size=K) * 0.3 + 1j * rng.normal(size=K) * 0.2).astype(C128)
    etar = etal.conj()
    etaa = np.abs(etal).astype(C128)
    psi = rng.normal(size=n) + 1j * rng.normal(size=n)
    rho0 = np.outer(psi, psi.conj())
    rho0 = (rho0 + rho0.conj().T) / 2 / np.trace(rho0).real   # exactly Hermitian, as kernel 6 requires
    return dict(system=H, system_dipole=np.zeros((n, n), C128), coupling=Q,
                coupling_dipole=np.zeros_like(Q), expn=expn, etal=etal, etar=etar, etaa=etaa,
                mode=mode, lmax=lmax, rho0=rho0.astype(C128), dt=0.01, nt=4)


def test_fmo_n7_real_h(emu):
    """The headline shape (N=7, projector couplings, real H) at depth 3: groups of
    4 ADOs, 120 ADOs = 30 full groups; the odd-N bulk-tile path."""
    check(emu, W.fmo(lmax=3, n_matsubara=0), nt=3)


def test_fmo_matsubara_links_beyond_one_chunk(emu):
    """K=21: up to 23 links per ADO at depth 2, i.e. four chunks of 7 from the strip."""
    check(emu, W.fmo(lmax=2, n_matsubara=2), nt=2, sm_count=2, warps=3)


@pytest.mark.parametrize("n,complex_h", [(2, False), (3, True), (4, True), (5, False), (6, True), (8, False)])
def test_every_system_size(emu, n, complex_h):
    """Even N (padded tiles, cp.async tile path), idle lanes (N=3,5,6,7), complex H."""
    check(emu, projector_problem(n, 1, 3, seed=n, complex_h=complex_h), nt=2)


def test_records_beyond_the_strip(emu):
    """N=2 with K=12: 12 + 3 links per ADO > 4 chunks of 2 -> on-demand record loads."""
    w = projector_problem(2, 6, 3, seed=11, complex_h=False)
    check(emu, w, nt=2)


def test_owned_ranges_and_rotation(emu):
    """Two owned ranges (a sharded run's ranks) with a ragged boundary, the
    visiting-order rotation of storage order 2, one warp per CTA."""
    w = W.fmo(lmax=3, n_matsubara=0)
    check(emu, w, nt=2, parts=[0, 57, 57, 120], scramble=1, sm_count=2, warps=1)
    check(emu, w, nt=2, scramble=1, sm_count=2, warps=1)   # whole hierarchy: kernel 7 too


def test_dynamic_group_schedule(emu):
    """Groups handed out by the global counter (SymArgs::sched) instead of the static stride:
    several launches share one never-reset counter, every warp draws exactly one index past the
    end (the host's running total must equal the device counter), full and packed storage,
    one and two owned ranges."""
    w = W.fmo(lmax=3, n_matsubara=0)
    emu.emu_set_dynsched(ctypes.c_int(1))
    emu.emu_sched_counter.restype = ctypes.c_longlong
    try:
        before = emu.emu_sched_counter()
        check(emu, w, nt=2, sm_count=2, warps=2)
        check(emu, w, nt=2, parts=[0, 57, 57, 120], sm_count=1, warps=3)
        after = emu.emu_sched_counter()
        assert after > before >= 0     # the dynamic path ran, and host total == device counter
    finally:
        emu.emu_set_dynsched(ctypes.c_int(0))


def test_more_warps_than_groups(emu):
    w = projector_problem(3, 1, 2, seed=5, complex_h=True)   # 10 ADOs = 1 group of 10
    check(emu, w, nt=3, sm_count=4, warps=4)


def _pack(full):
    n = full.shape[-1]
    iu = np.triu_indices(n)
    return np.ascontiguousarray(full[..., iu[0], iu[1]])


def _unpack(tri, n):
    iu = np.triu_indices(n)
    full = np.zeros(tri.shape[:-1] + (n, n), dtype=C128)
    full[..., iu[0], iu[1]] = tri
    low = np.conj(np.swapaxes(full, -1, -2))
    idx = np.tril_indices(n, -1)
    full[..., idx[0], idx[1]] = low[..., idx[0], idx[1]]
    return full


@pytest.mark.parametrize("shape", ["fmo N=7", "even N=4"])
@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("fused", [False, True])
def test_sharded_ranks_with_local_arrays(emu, fused, packed, shape):
    """The rank-local layout of ``csrc/heom_shard.cu`` with two ranks: a rank's arrays hold only
    its own ADOs (full matrices, or upper triangles = kernel 7) followed by a pool of halo rows,
    its link table points at local slots and pool rows.  Everything starts as NaN, so a read of
    anything that was not owned or exchanged would poison the result.  ``fused``: no exchange by
    the test at all - the PUSH instantiation's epilogue stores the rows the other rank reads
    into that rank's pool (bulk shared->global stores, tables as ``ShardedDEOM`` builds them)."""
    w = W.fmo(lmax=3, n_matsubara=0) if "fmo" in shape else projector_problem(4, 2, 3, seed=31, complex_h=True)
    o, t = host_tables(w)
    N, nmax, dt, nt = t["N"], o.nmax, w["dt"], 2
    EL = N * (N + 1) // 2 if packed else N * N
    bounds = [0, 53, nmax]
    ptr, links = t["link_ptr"], t["links"]
    err = ctypes.c_char_p()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    links2 = np.zeros_like(links)
    assert emu.emu_convert_links(p(links), p(links2), ctypes.c_longlong(len(links)), ctypes.c_int(N),
                                 ctypes.c_int(o.lmax), ctypes.c_int(int(packed)), ctypes.byref(err)) == 0
    need = []   # per rank: sorted (slot, row) of the foreign rows its links read
    for r in range(2):
        lo, hi = bounds[r], bounds[r + 1]
        rec = links[ptr[lo]:ptr[hi]]
        foreign = (rec[:, 0] < lo) | (rec[:, 0] >= hi)
        nd = sorted({(int(s_), int((m_ >> 16) & 0xf)) for s_, m_ in rec[foreign]})
        assert nd
        need.append(nd)
    n_own = [bounds[1], nmax - bounds[1]]
    n_own_max, pool_max = max(n_own), max(len(x) for x in need)
    PS = (N + 1) & ~1                       # sym_pool_stride: whole sectors per pool row
    pool_off = (n_own_max * EL + 7) & ~7    # sym_pool_offset
    arr_elems = pool_off + pool_max * PS
    # the rank-local link tables (shard_localize_links_kernel restated): x = element offset of the row's
    # storage (local ADO: as the converter resolves it for the local slot; foreign: its pool row),
    # table row (bits 28-31 of y) = r0, or N for pool rows
    local_links = []
    for r in range(2):
        lo, hi = bounds[r], bounds[r + 1]
        pos = {item: i for i, item in enumerate(need[r])}
        loc = links2.copy()
        for l in range(ptr[lo], ptr[hi]):
            nb, r0 = int(links[l, 0]), int((links[l, 1] >> 16) & 0xf)
            y_ = int(links2[l, 1]) & 0x0fffffff
            if lo <= nb < hi:
                x_ = (nb - lo) * EL if packed else ((nb - lo) * N + r0) * N
                tr = r0
            else:
                x_ = pool_off + pos[(nb, r0)] * PS
                tr = N
            yy = y_ | (tr << 28)
            loc[l] = (x_, yy - (1 << 32) if yy >= (1 << 31) else yy)
        local_links.append(loc)
    state = np.full((2, 4, arr_elems), np.nan, dtype=C128)   # rank, (Y, SA, SB, ACC)
    y0 = np.zeros((nmax, N, N), C128)
    y0[0] = w["rho0"]
    for r in range(2):
        own = y0[bounds[r]:bounds[r + 1]]
        state[r, 0, :n_own[r] * EL] = (_pack(own) if packed else own).reshape(-1)

    def own_ados(r, arr):
        flat = state[r, arr, :n_own[r] * EL]
        return _unpack(flat.reshape(n_own[r], EL), N) if packed else flat.reshape(n_own[r], N, N)

    def exchange(arr):
        full = [own_ados(r, arr) for r in range(2)]
        for r in range(2):
            for i, (slot, row) in enumerate(need[r]):
                src = full[1 - r][slot - bounds[1 - r], row].copy()
                if packed:   # as a gather through the triangle delivers the row: (min, max), not conjugated
                    src[:row] = np.conj(src[:row])
                state[r, arr, pool_off + i * PS: pool_off + i * PS + N] = src
    exchange(0)
    # push tables of rank r: CSR over its owned slots, entry = (row index in the peer's pool,
    # staging slot << 8 | peer << 4 | row); the distinct (ADO, row) pairs of a group of 32/N
    # consecutive ADOs are numbered in order, 255 beyond the kernel's staging area (12 slots)
    push_ptr, push_ent = [], []
    apw = 32 // N
    for r in range(2):
        lo, hi = bounds[r], bounds[r + 1]
        per_slot = [[] for _ in range(hi - lo)]
        for i, (slot, row) in enumerate(need[1 - r]):          # what the other rank reads from this one
            per_slot[slot - lo].append((row, i))
        ents = []
        for g0 in range(0, hi - lo, apw):
            ids = {}
            for sl in range(g0, min(g0 + apw, hi - lo)):
                per_slot[sl].sort()
                for row, i in per_slot[sl]:
                    sid = ids.setdefault((sl, row), len(ids))
                    ents.append((i, (min(sid, 255) if sid < 12 else 255) << 8 | ((1 - r) << 4) | row))
        push_ptr.append(np.concatenate([[0], np.cumsum([len(x) for x in per_slot])]).astype(np.int32))
        push_ent.append(np.array(ents or [(0, 0)], dtype=np.int32))
    peers = np.array([state[0].ctypes.data, state[1].ctypes.data], dtype=np.uint64)
    H = np.ascontiguousarray(o.H0)
    emu.emu_sym_stage.restype = ctypes.c_int
    plan = [(0, 1, 0, dt / 2, 0.0), (1, 2, 1, dt / 2, 0.0), (2, 3, 1, dt, 0.0), (3, 0, 2, 2.0 / dt, dt / 6)]
    for _ in range(nt):
        for yin, out, kind, a, wgt in plan:
            for r in range(2):
                st = state[r]
                lo, hi = bounds[r], bounds[r + 1]
                damp_r = np.ascontiguousarray(t["damp"][lo:hi])
                ptr_r = np.ascontiguousarray(ptr[lo:hi + 1])
                rc = emu.emu_sym_stage(
                    ctypes.c_int(N), ctypes.c_int(t["K"]), ctypes.c_int(t["M"]), ctypes.c_int(o.lmax), p(H),
                    p(t["ops"]), p(t["cbase"]), p(t["kmode"]), p(damp_r), p(ptr_r), p(local_links[r]),
                    p(st[yin]), p(st[0]), p(st[1]), p(st[2]), p(st[out]),
                    ctypes.c_double(a), ctypes.c_double(wgt), ctypes.c_int(kind),
                    ctypes.c_int(int(np.all(H.imag == 0))), ctypes.c_int(2), ctypes.c_int(2),
                    ctypes.c_longlong(n_own[r]), ctypes.c_int(int(packed)), ctypes.c_longlong(pool_off),
                    p(push_ptr[r]) if fused else None, p(push_ent[r]) if fused else None,
                    p(peers) if fused else None, ctypes.c_longlong(out * arr_elems + pool_off), ctypes.byref(err))
                assert rc == 0, err.value
            if not fused:
                exchange(out)
    o.run(w["rho0"], dt, nt)
    got = np.concatenate([own_ados(0, 0), own_ados(1, 0)])
    assert np.isfinite(got).all()
    assert np.abs(got - o.ddos).max() < 1e-12
    # pool rows nobody asked for were never written
    for r in range(2):
        if len(need[r]) < pool_max:
            assert np.isnan(state[r, 0, pool_off + len(need[r]) * PS:]).all()
