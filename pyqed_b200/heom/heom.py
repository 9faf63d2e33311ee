"""``HEOMSolver`` - drop-in for the single-exponential chain HEOM of
``pyqed/HEOM/heom.py::HEOMSolver`` (:161-205) and its RK4 driver ``_heom``
(:275-347).

The chain is the DEOM hierarchy with one dissipaton: with
``D0 = lambda (2T - i gamma)`` (``heom.py:312``), ``K = 1, etal = D0, etar = conj(D0),
etaa = |D0|, lmax = nado - 2`` reproduces rho_sys(t) of the chain (the ADOs
themselves differ by the fixed rescaling i^n sqrt(n! |D0|^n)), so the same
device tables and stage kernel serve both solvers.  ``run`` returns
``complex[len(e_ops), nt]`` - the value after each step, no t = 0 column - like
the reference.
"""
from __future__ import annotations

import numpy as np

from .bath import Bath
from .deom import DEOMSolver

C128 = np.complex128


class HEOMSolver:
    def __init__(self, H=None, c_ops=None, e_ops=None, device=0, verbose=True):
        self.c_ops = c_ops
        self.e_ops = e_ops
        self.H = H
        self.device = device
        self.verbose = verbose
        self._deom = None

    def set_c_ops(self, c_ops):
        self.c_ops = c_ops

    def set_e_ops(self, e_ops):
        self.e_ops = e_ops

    def setH(self, H):
        self.H = H

    def configure(self, c_ops, e_ops):
        self.c_ops = c_ops
        self.e_ops = e_ops

    @staticmethod
    def amplitude(temperature, cutoff, reorganization):
        """``D0`` of ``heom.py:312`` (temperature already in energy units)."""
        return reorganization * (2.0 * temperature - 1j * cutoff)

    def propagator(self, dt, nt, temperature, cutoff, reorganization, nado):
        """Liouville-space Euler propagator ``u[nado, N^2, N^2]`` of
        ``pyqed/HEOM/heom.py:349-413`` (temperature in kelvin)."""
        from ..oqs import liouville_propagator

        def say(*lines):
            if self.verbose:
                for line in lines:
                    print(line)
        return liouville_propagator(self.H, self.c_ops[0], dt, nt, temperature, cutoff, reorganization,
                                    nado, double_update0=False, device=self.device, say=say)

    def run(self, rho0, dt, nt, temperature, cutoff, reorganization, nado):
        if self.H is None or self.c_ops is None or self.e_ops is None:
            raise ValueError('H, c_ops and e_ops must be set.')
        if nado < 2:
            raise ValueError('nado must be >= 2 (the last ADO is the terminator).')
        gamma, T = cutoff, temperature
        D0 = self.amplitude(T, gamma, reorganization)
        if self.verbose:  # same information the reference prints (heom.py:298-314)
            print('Temperature of the environment = {}'.format(T))
            print('Cutoff gamma/(kT) = {}'.format(gamma / T))
            if gamma / T > 0.8:
                print('WARNING: High-Temperature Approximation may fail.')
            print('Reorganization energy = {}'.format(reorganization))
            print('Amplitude of the fluctuations = {}'.format(D0))
        H = np.asarray(self.H, dtype=C128)
        S = np.asarray(self.c_ops[0], dtype=C128)  # only c_ops[0] is used (heom.py:317)
        bath = Bath(expn=[gamma], etal=[D0], etar=[np.conj(D0)], etaa=[abs(D0)], mode=[0])
        solver = DEOMSolver(system=H, system_dipole=None, bath=bath, coupling=S[None],
                            coupling_dipole=None, lmax=nado - 2, device=self.device,
                            alias_rho0=False)
        self._deom = solver
        solver.check_()
        solver.init_()
        e_ops = np.stack([np.asarray(e, dtype=C128) for e in self.e_ops])
        plan = solver._ensure_plan(1)
        import torch
        n = H.shape[0]
        plan.set_state(np.asarray(rho0, dtype=C128).reshape(1, n, n))
        traj = torch.empty((1, nt + 1, n, n), dtype=torch.complex128,
                           device=torch.device("cuda", self.device))
        plan.propagate(dt, nt, None, None, traj, method=0)
        obs = plan.expectation(traj, e_ops)[0].cpu().numpy()
        return np.ascontiguousarray(obs[:, 1:])
