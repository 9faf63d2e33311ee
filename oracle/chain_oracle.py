"""NumPy restatement of the reference's single-exponential chain HEOM.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  Parity status: pinned
against ``tests/golden/chain_*.npz`` (outputs of the unmodified reference).

* ``heom_chain_rk4``   - ``_heom`` in ``pyqed/HEOM/heom.py:275-347`` driven by
  ``rk4`` in ``pyqed/phys.py:1051-1064``; high-temperature Drude coefficient
  ``D0 = lambda (2T - i gamma)`` (``heom.py:312``).
* ``heom_chain_euler`` - ``_heom`` in ``pyqed/oqs.py:1808-1875``: explicit Euler
  with the in-place (Gauss-Seidel-like) update order of ``oqs.py:1860-1867`` and
  ``D0 = lambda gamma (coth(gamma / 2T) - i)`` (``oqs.py:1844``).
* ``chain_as_deom``    - the table fill that maps the chain onto the DEOM form
  (SURVEY.md section 8a, KAT-3): K=1, etal=D0, etar=conj(D0), etaa=|D0|,
  lmax=nado-2; rho_sys then agrees to rounding.

ADO layout is the reference's ``[N, N, nado]`` (ADO index fastest); only
``c_ops[0]`` is used (``heom.py:317``); ADO ``nado-1`` has zero right-hand side.
"""
from __future__ import annotations

import numpy as np


def _cm(a, b):
    return a @ b - b @ a


def _acm(a, b):
    return a @ b + b @ a


def d0_high_temperature(temperature, cutoff, reorganization):
    """``pyqed/HEOM/heom.py:312``."""
    return reorganization * (2.0 * temperature - 1j * cutoff)


def d0_coth(temperature, cutoff, reorganization):
    """``pyqed/oqs.py:1844``."""
    return reorganization * cutoff * (1.0 / np.tanh(cutoff / (2.0 * temperature)) - 1j)


def _chain_rhs(ado, H, S, gamma, D0):
    """Inner ``L`` of ``pyqed/HEOM/heom.py:319-330``."""
    nado = ado.shape[2]
    out = np.zeros_like(ado)
    out[:, :, 0] = -1j * _cm(H, ado[:, :, 0]) - _cm(S, ado[:, :, 1])
    for n in range(1, nado - 1):
        out[:, :, n] = (-1j * _cm(H, ado[:, :, n]) - _cm(S, ado[:, :, n + 1])
                        - n * gamma * ado[:, :, n]
                        + n * (D0.real * _cm(S, ado[:, :, n - 1])
                               + 1j * D0.imag * _acm(S, ado[:, :, n - 1])))
    return out


def _expect(rho, op):
    """``obs`` of ``pyqed/superoperator.py:313-314``: ``vdot(dag(a).ravel(), rho)``
    which equals ``Tr(a rho)``."""
    return np.vdot(op.conj().T.ravel(), rho.ravel())


def heom_chain_rk4(H, rho0, c_ops, e_ops, temperature, cutoff, reorganization,
                   nado, dt, nt, return_ados=False):
    H = np.asarray(H, dtype=np.complex128)
    S = np.asarray(c_ops[0], dtype=np.complex128)
    e_ops = [np.asarray(e, dtype=np.complex128) for e in e_ops]
    n = H.shape[0]
    ado = np.zeros((n, n, nado), dtype=np.complex128)
    ado[:, :, 0] = rho0
    D0 = d0_high_temperature(temperature, cutoff, reorganization)
    obs = np.zeros((len(e_ops), nt), dtype=np.complex128)
    half = dt / 2.0
    for step in range(nt):
        k1 = _chain_rhs(ado, H, S, cutoff, D0)
        k2 = _chain_rhs(ado + k1 * half, H, S, cutoff, D0)
        k3 = _chain_rhs(ado + k2 * half, H, S, cutoff, D0)
        k4 = _chain_rhs(ado + k3 * dt, H, S, cutoff, D0)
        ado = ado + (k1 + 2 * k2 + 2 * k3 + k4) / 6.0 * dt
        obs[:, step] = [_expect(ado[:, :, 0], e) for e in e_ops]
    return (obs, ado) if return_ados else obs


def heom_chain_euler(H, rho0, c_ops, e_ops, temperature, cutoff, reorganization,
                     nado, dt, nt, return_ados=False):
    H = np.asarray(H, dtype=np.complex128)
    S = np.asarray(c_ops[0], dtype=np.complex128)
    e_ops = [np.asarray(e, dtype=np.complex128) for e in e_ops]
    n = H.shape[0]
    ado = np.zeros((n, n, nado), dtype=np.complex128)
    ado[:, :, 0] = rho0
    D0 = d0_coth(temperature, cutoff, reorganization)
    obs = np.zeros((len(e_ops), nt), dtype=np.complex128)
    for step in range(nt):
        # sequential in-place sweep: ADO n sees the already-updated ADO n-1
        ado[:, :, 0] = (ado[:, :, 0] - 1j * _cm(H, ado[:, :, 0]) * dt
                        - _cm(S, ado[:, :, 1]) * dt)
        for m in range(1, nado - 1):
            ado[:, :, m] = ado[:, :, m] + (
                -1j * _cm(H, ado[:, :, m]) * dt
                + (-_cm(S, ado[:, :, m + 1]) - m * cutoff * ado[:, :, m]
                   + m * (D0.real * _cm(S, ado[:, :, m - 1])
                          + 1j * D0.imag * _acm(S, ado[:, :, m - 1]))) * dt)
        obs[:, step] = [_expect(ado[:, :, 0], e) for e in e_ops]
    return (obs, ado) if return_ados else obs


def chain_as_deom(D0, cutoff, nado):
    """Bath arrays + depth that make the DEOM form reproduce the chain's
    rho_sys: ``(expn, etal, etar, etaa, mode, lmax)``."""
    c128 = np.complex128
    return (np.array([cutoff], dtype=c128), np.array([D0], dtype=c128),
            np.array([np.conj(D0)], dtype=c128), np.array([abs(D0)], dtype=c128),
            np.array([0], dtype=np.int64), nado - 2)
