"""Minimal ``Mol`` holder with the ``deom`` factory of ``pyqed/mol.py:755-763``.

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


class Mol:
    def __init__(self, H, edip=None):
        self.H = np.asarray(H, dtype=np.complex128)
        self.edip = (np.zeros_like(self.H) if edip is None
                     else np.asarray(edip, dtype=np.complex128))
        self.dim = self.H.shape[0]

    def getH(self):
        return self.H

    def deom(self, bath, coupling, coupling_dipole=None, pulse_system_func=None,
             pulse_coupling_func=None, mode=None, **solver_kwargs):
        """hierarchical equations of motion (``mol.py:755``)"""
        return DEOMSolver(self.H, self.edip, bath, coupling, coupling_dipole, pulse_system_func,
                          pulse_coupling_func, mode, **solver_kwargs)
