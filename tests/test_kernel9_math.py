"""Kernel 9's arithmetic (pyqed_b200/csrc/heom_dataflow_tma.cuh), restated in NumPy and held against
the oracle's right-hand side (generate_dot_element, pyqed/heom/deom.py:641-664) on the CPU:

* a Hermitian problem (Hermitian H, Q_m; real exponents, eta_r = conj eta_l, eta_a > 0) keeps every ADO
  Hermitian and makes the link coefficients obey alphaR = conj(alphaL), so
  d rho/dt = W + W^dagger with W = (-iH - gamma/2) rho + sum_m Q_m S_m, S_m = sum_{links of m} alphaL rho';
* the state of an ADO is one complex number per UNIT - the pair (i,j),(j,i) with i < j holds rho_ij, a
  diagonal unit holds two real diagonal entries - and S_m for both elements of a unit comes from four real
  sums over the links (A1..A4), identically for both kinds of unit;
* operators are applied through zero-padded, entry-major sparse rows.

The GPU tests (tests/test_gpu_parity.py, kernel = 9) check the kernel itself; this file pins the algebra
it relies on, including odd N (a diagonal unit with one element) and a non-diagonal coupling operator."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle.deom_oracle import DeomOracle  # noqa: E402


def _units(n):
    """(ia, ja, ib, jb, isdiag, hasb) per unit, in the kernel's order."""
    out = []
    for i in range(n):
        for j in range(i + 1, n):
            out.append((i, j, j, i, False, True))
    for d in range(0, n, 2):
        hasb = d + 1 < n
        out.append((d, d, d + 1 if hasb else d, d + 1 if hasb else d, True, hasb))
    return out


def _pack(rho, units):
    v = np.zeros(len(units), complex)
    for u, (ia, ja, ib, jb, isdiag, hasb) in enumerate(units):
        v[u] = complex(rho[ia, ia].real, rho[ib, ib].real if hasb else 0.0) if isdiag else rho[ia, ja]
    return v


def _padded_rows(op):
    """Entry-major padded sparse rows: val[k][row], col[k][row]; padding = (0, row)."""
    n = op.shape[0]
    rows = [np.nonzero(op[i])[0] for i in range(n)]
    longest = max((len(r) for r in rows), default=0)
    val = np.zeros((longest, n), complex)
    col = np.tile(np.arange(n), (longest, 1))
    for i, r in enumerate(rows):
        val[:len(r), i] = op[i, r]
        col[:len(r), i] = r
    return val, col


def _apply_rows(val, col, src, i, j):
    return sum(val[k, i] * src[col[k, i], j] for k in range(val.shape[0]))


def kernel9_rhs(o, rho):
    """d rho/dt of every ADO the way kernel 9 forms it; returns full matrices."""
    n, units = o.nsys, _units(o.nsys)
    H, Q = o.operators_at(0.0)
    a_val, a_col = _padded_rows(-1j * H)
    q_rows = [_padded_rows(q) for q in Q]
    q_diag = [np.count_nonzero(q - np.diag(np.diag(q))) == 0 for q in Q]
    packed = np.array([_pack(r, units) for r in rho])       # what the neighbours publish
    out = np.zeros_like(rho)
    for ado in range(o.nmax):
        key = o.keys[ado]
        gamma = float(np.sum(key * o.expn).real)
        # links: (neighbour, mode, alphaL)
        links = []
        for k in range(o.nind):
            sa = np.sqrt(o.etaa[k])
            if key[k] > 0:
                links.append((o.minus[ado, k], o.mode[k], -1j * np.sqrt(key[k]) / sa * o.etal[k]))
            if key.sum() < o.lmax:
                links.append((o.plus[ado, k], o.mode[k], -1j * np.sqrt(key[k] + 1) * sa))
        full = rho[ado]
        k_units = np.zeros(len(units), complex)
        # S_m of both elements of every unit from the four real sums
        S = {}
        for m in range(len(Q)):
            A = np.zeros((4, len(units)))
            for nb, mode, cl in links:
                if mode != m:
                    continue
                x = packed[nb]
                A[0] += cl.real * x.real
                A[1] += cl.imag * x.imag
                A[2] += cl.real * x.imag
                A[3] += cl.imag * x.real
            S[m] = A
        # full S_m matrices (the kernel's shared-memory tile) for the non-diagonal modes
        S_full = {}
        for m in range(len(Q)):
            if q_diag[m]:
                continue
            t = np.zeros((n, n), complex)
            for u, (ia, ja, ib, jb, isdiag, hasb) in enumerate(units):
                A1, A2, A3, A4 = S[m][:, u]
                sa_, sb_ = ((A1 + 1j * A4, A3 + 1j * A2) if isdiag else (A1 - A2 + 1j * (A3 + A4), A1 + A2 + 1j * (A4 - A3)))
                t[ia, ja] = sa_
                if hasb:
                    t[ib, jb] = sb_
            S_full[m] = t
        for u, (ia, ja, ib, jb, isdiag, hasb) in enumerate(units):
            wa = -0.5 * gamma * full[ia, ja] + _apply_rows(a_val, a_col, full, ia, ja)
            wb = -0.5 * gamma * full[ib, jb] + _apply_rows(a_val, a_col, full, ib, jb)
            for m in range(len(Q)):
                A1, A2, A3, A4 = S[m][:, u]
                sa_, sb_ = ((A1 + 1j * A4, A3 + 1j * A2) if isdiag else (A1 - A2 + 1j * (A3 + A4), A1 + A2 + 1j * (A4 - A3)))
                val, col = q_rows[m]
                if q_diag[m]:
                    if val.shape[0]:
                        wa += val[0, ia] * sa_
                        wb += val[0, ib] * sb_
                else:
                    wa += _apply_rows(val, col, S_full[m], ia, ja)
                    wb += _apply_rows(val, col, S_full[m], ib, jb)
            k_units[u] = complex(2 * wa.real, 2 * wb.real) if isdiag else wa + np.conj(wb)
        for u, (ia, ja, ib, jb, isdiag, hasb) in enumerate(units):
            if isdiag:
                out[ado, ia, ia] = k_units[u].real
                if hasb:
                    out[ado, ib, ib] = k_units[u].imag
            else:
                out[ado, ia, ja] = k_units[u]
                out[ado, ib, jb] = np.conj(k_units[u])
    return out


def _hermitian_problem(n, nmod, nind, lmax, seed, offdiag_q):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    H = a + a.conj().T
    H[np.abs(H) < 1.2] = 0                                   # sparse rows of different lengths
    H = np.triu(H) + np.triu(H, 1).conj().T
    Q = np.zeros((nmod, n, n), complex)
    for m in range(nmod):
        Q[m] = np.diag(rng.normal(size=n))
        if offdiag_q and m == nmod - 1:                      # a bidiagonal Hermitian coupling operator
            off = rng.normal(size=n - 1) + 1j * rng.normal(size=n - 1)
            Q[m] = np.diag(rng.normal(size=n)) + np.diag(off, 1) + np.diag(off.conj(), -1)
    mode = np.arange(nind) % nmod
    expn = rng.uniform(0.5, 2.0, nind).astype(complex)
    etal = rng.normal(size=nind) + 1j * rng.normal(size=nind)
    etar = etal.conj()
    etaa = np.abs(etal)
    o = DeomOracle(H, None, Q, None, expn, etal, etar, etaa, mode, lmax)
    rho = np.zeros((o.nmax, n, n), complex)
    for k in range(o.nmax):                                  # every ADO Hermitian, as the dynamics keeps them
        b = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        rho[k] = b + b.conj().T
    return o, rho


@pytest.mark.parametrize("n,nmod,nind,lmax,offdiag", [(4, 2, 4, 3, True), (5, 2, 3, 2, True), (7, 3, 3, 2, False),
                                                      (2, 1, 2, 4, False), (3, 1, 2, 3, True)])
def test_kernel9_form_equals_the_reference_rhs(n, nmod, nind, lmax, offdiag):
    o, rho = _hermitian_problem(n, nmod, nind, lmax, seed=10 * n + lmax, offdiag_q=offdiag)
    ref = o.rhs_batched(rho, 0.0)
    got = kernel9_rhs(o, rho)
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(got - ref)) < 1e-13 * scale
    # the packed form is Hermitian by construction, bit for bit
    assert np.array_equal(got, np.conj(np.swapaxes(got, 1, 2)))


def test_unit_count_fits_one_cta_up_to_n32():
    for n in range(2, 33):
        units = _units(n)
        assert len(units) == n * (n - 1) // 2 + (n + 1) // 2 <= 512
        covered = {(ia, ja) for ia, ja, *_ in units} | {(ib, jb) for _, _, ib, jb, _, hasb in units if hasb}
        assert covered == {(i, j) for i in range(n) for j in range(n)}
