"""Minimal ``Mol`` holder with the ``deom`` factory of ``pyqed/mol.py:755-763``
and the ``Result`` container of ``pyqed/mol.py:98-165``.

Only what ``examples/deom.py:41,67`` uses on the HEOM path: ``Mol(H, edip)``
keeps the system Hamiltonian and the transition dipole, ``Mol.deom(...)``
returns a ``DEOMSolver`` wired to them.  (The reference factory imports
``pyqed.HEOM.deom``, which does not exist on a case-sensitive file system, and
passes ``mode`` positionally into the solver's ``lmax`` slot; here ``mode`` is
accepted for signature compatibility and, like there, ends up as ``lmax`` -
callers then use ``set_hierarchy`` exactly as ``examples/deom.py:69`` does.)
"""
from __future__ import annotations

import numpy as np

from .heom.deom import DEOMSolver


class Result:
    """Container the reference's solvers return their output in (``mol.py:98-165``):
    ``times`` (``t0 + i dt nout``), ``observables`` (``[time, operator]``),
    ``rholist`` / ``psilist``, ``rho0`` / ``psi0``; ``dump`` / ``save`` pickle it.
    (The reference's ``analyze`` needs proplot, which neither side has here;
    it raises the same ``ValueError`` when there is nothing to plot.)"""

    def __init__(self, description=None, psi0=None, rho0=None, dt=None, Nt=None, times=None,
                 t0=0, nout=1):
        self.description = description
        self.dt = dt
        self.timesteps = self.nt = Nt
        self.observables = None
        self.rholist = None
        self.psilist = []
        self.psi = None
        self.rho = None
        self.rho0 = rho0
        self.psi0 = psi0
        self.nout = nout
        self.times = t0 + np.arange(Nt // nout + 1) * dt * nout if times is None else np.asarray(times)

    def expect(self):
        return self.observables

    def analyze(self):
        if self.observables is None or self.observables.shape[-1] == 0:
            raise ValueError('There are no observables to analyze.')
        raise NotImplementedError("plotting is outside the HEOM path")

    def dump(self, fname):
        import pickle
        solver = self.__dict__.pop("solver", None)   # device handles do not pickle
        try:
            with open(fname, 'wb') as f:
                pickle.dump(self, f)
        finally:
            if solver is not None:
                self.solver = solver

    def save(self, fname):
        self.dump(fname)


class Mol:
    def __init__(self, H, edip=None, lowering=None, edip_rms=None, gamma=None):
        self.H = self.h = np.asarray(H, dtype=np.complex128)
        self.edip = self.dip = (np.zeros_like(self.H) if edip is None
                                else np.asarray(edip, dtype=np.complex128))
        self.lowering = self.raising = None
        if lowering is not None:
            self.lowering = np.asarray(lowering, dtype=np.complex128)
            self.raising = self.lowering.conj().T
        self.dim = self.nstates = self.size = self.H.shape[0]
        self.idm = np.eye(self.dim, dtype=np.complex128)
        self.gamma = gamma

    def getH(self):
        return self.H

    def deom(self, bath, coupling, coupling_dipole=None, pulse_system_func=None,
             pulse_coupling_func=None, mode=None, **solver_kwargs):
        """hierarchical equations of motion (``mol.py:755``); ``lmax=...`` may be
        given by keyword and then wins over the positional ``mode`` quirk"""
        lmax = solver_kwargs.pop("lmax", mode)
        return DEOMSolver(self.H, self.edip, bath, coupling, coupling_dipole, pulse_system_func,
                          pulse_coupling_func, lmax, **solver_kwargs)
