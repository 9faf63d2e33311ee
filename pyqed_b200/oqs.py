"""``pyqed_b200.oqs.HEOMSolver`` - drop-in for the *Euler* chain solver of
``pyqed/oqs.py:1332-1371`` (the class ``examples/heom.py:15`` imports).

Same equations as ``pyqed_b200.heom.HEOMSolver`` but advanced like the
reference's ``_heom`` in ``oqs.py:1808-1875``: explicit Euler with an in-place
sweep in which ADO n sees the already-advanced ADO n-1, and the amplitude
``D0 = lambda gamma (coth(gamma / 2T) - i)`` (``oqs.py:1844``).  The sweep is a
sequential dependence chain, so it runs as one small CUDA kernel
(``chain_euler_kernel``), one CTA per trajectory.
"""
from __future__ import annotations

import numpy as np

from ._cabi import chain_euler

C128 = np.complex128
AU2K = 315775.13  # pyqed/units.py:6


def _unit_matrix_ados(n, nado):
    """ADO 0 of trajectory c = unit matrix E_c (row-major c = i*N + j): the
    columns of ``u[0] = eye(N^2)`` in ``_heom_propagator``."""
    a = np.zeros((n * n, nado, n, n), dtype=C128)
    for c in range(n * n):
        a[c, 0].flat[c] = 1.0
    return a


def _columns_to_superoperators(ados):
    """[N^2 trajectories, nado, N, N] -> u[nado, N^2, N^2] with
    ``u[n][:, c] = vec(ADO n of trajectory c)``."""
    nn, nado, n, _ = ados.shape
    return np.ascontiguousarray(ados.reshape(nn, nado, nn).transpose(1, 2, 0))


class HEOMSolver:
    def __init__(self, H=None, c_ops=None, e_ops=None, device=0, verbose=True):
        self.c_ops = c_ops
        self.e_ops = e_ops
        self.H = H
        self.device = device
        self.verbose = verbose

    def set_c_ops(self, c_ops):
        self.c_ops = c_ops

    def set_e_ops(self, e_ops):
        self.e_ops = e_ops

    def setH(self, H):
        self.H = H

    def configure(self, c_ops, e_ops):
        self.c_ops = c_ops
        self.e_ops = e_ops

    def _say(self, *lines):
        if self.verbose:
            for line in lines:
                print(line)

    def run(self, rho0, dt, nt, temperature, cutoff, reorganization, nado):
        """``observables[len(e_ops), nt]`` (value after every step), Euler."""
        if nado < 2:
            raise ValueError('nado must be >= 2 (the last ADO is the terminator).')
        gamma, T = cutoff, temperature
        D0 = reorganization * gamma * (1.0 / np.tanh(gamma / (2.0 * T)) - 1j)
        self._say('Temperature of the environment = {}'.format(T),
                  'Cutoff gamma/(kT) = {}'.format(gamma / T),
                  *(['WARNING: High-Temperature Approximation may fail.'] if gamma / T > 0.8 else []),
                  'Reorganization energy = {}'.format(reorganization),
                  'Amplitude of the fluctuations = {}'.format(D0))
        H = np.asarray(self.H, dtype=C128)
        n = H.shape[0]
        ado = np.zeros((1, nado, n, n), dtype=C128)
        ado[0, 0] = rho0
        _, obs = chain_euler(H, self.c_ops[0], ado, gamma, D0, dt, nt, e_ops=list(self.e_ops),
                             device=self.device)
        return obs[0]

    def propagator(self, dt, nt, temperature, cutoff, reorganization, nado):
        """``u[nado, N^2, N^2]`` of ``oqs.py:1877-1941`` (temperature in kelvin;
        its loop advances ``u[0]`` twice per step, reproduced here)."""
        return liouville_propagator(self.H, self.c_ops[0], dt, nt, temperature, cutoff, reorganization,
                                    nado, double_update0=True, device=self.device,
                                    say=self._say)


def liouville_propagator(H, S, dt, nt, temperature, cutoff, reorganization, nado,
                         double_update0=False, device=0, say=None):
    """Shared body of the two ``_heom_propagator`` variants: amplitude
    ``a = pi lambda T`` with ``T = temperature / au2k``, ``b = 0``."""
    if nado < 2:
        raise ValueError('nado must be >= 2 (the last ADO is the terminator).')
    T = temperature / AU2K
    a = np.pi * reorganization * T
    if say:
        say('Temperature of the environment = {}'.format(T),
            'High-Temperature check gamma/(kT) = {}'.format(cutoff / T),
            *(['WARNING: High-Temperature Approximation may fail.'] if cutoff / T > 0.8 else []),
            'Reorganization energy = {}'.format(reorganization),
            'Amplitude of the fluctuations = {}'.format(a))
    H = np.asarray(H, dtype=C128)
    n = H.shape[0]
    ados, _ = chain_euler(H, S, _unit_matrix_ados(n, nado), cutoff, a + 0j, dt, nt,
                          double_update0=double_update0, device=device)
    return _columns_to_superoperators(ados)
