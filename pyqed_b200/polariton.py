"""Cavity-molecule model builders that feed the HEOM path.

Mirrors the part of ``pyqed/polariton/cavity.py`` that BASELINE config 4 needs -
``Cavity`` (``cavity.py:404-533``) and ``Polariton.getH`` (``cavity.py:577-678``,
length gauge) - so that a polariton model can be written with the reference's
own vocabulary and handed to ``DEOMSolver`` without touching the reference:

    mol = Mol(0.5 * w0 * sz, edip=sx, lowering=sm)
    pol = Polariton(mol, Cavity(wc, 16), g=0.1)
    H = pol.getH(RWA=False)
    solver = pol.deom(bath, [pol.promote_op(sz, 'mol'), pol.promote_op(cav.quadrature(), 'cav')])

Differences from the reference, on purpose: operators are dense ``complex128``
arrays (the reference returns scipy sparse matrices; the solver needs dense
N x N blocks anyway), and only the length/dipole gauge is provided.

``Env`` / ``env_bath`` restate the high-temperature two-bath environment of
``pyqed/polariton/exact.py:41-50, 660-671`` (``theta``) as a ``Bath`` with one
exponent per collapse operator, and ``heom`` is the consistent RK4 version of
the (dead, hard-coded 5 x 5 Euler) loop at ``exact.py:674-794``: same model,
same observables, propagated by the DEOM kernel.  The reference loop itself is
not reproduced - it is never called there, freezes the edge ADOs and mixes
``gamma`` and tier prefactors (``exact.py:738-756``).
"""
from __future__ import annotations

import numpy as np

from .mol import Mol, Result

C128 = np.complex128

__all__ = ["ham_ho", "Cavity", "Polariton", "Env", "env_bath", "heom"]


def ham_ho(freq, n, ZPE=False):
    """Harmonic-oscillator Hamiltonian in the number basis (``cavity.py:286-301``).
    Like the reference, ``ZPE=True`` takes ``arange(n + 0.5) * freq`` - which is
    the same ladder without a zero-point shift - so the flag changes nothing."""
    return np.diagflat(np.arange(n) * float(freq)).astype(C128)


class Cavity:
    """Single cavity mode truncated to ``n_cav`` Fock states (``cavity.py:404-533``)."""

    def __init__(self, freq, n_cav=None, x=None, decay=None, g=None):
        if n_cav is None or int(n_cav) < 1:
            raise ValueError("n_cav must be a positive integer")
        self.freq = self.omega = self.omegac = freq
        self.resonance = freq
        self.ncav = self.n_cav = self.n = self.dim = int(n_cav)
        self.idm = np.eye(self.n_cav, dtype=C128)
        self.H = self.getH()
        if x is not None:
            self.x = x
            self.nx = len(x)
        self.decay = decay
        self.g = g

    def getH(self, zpe=False):
        return ham_ho(self.freq, self.n_cav)

    def create(self):
        return np.diag(np.sqrt(np.arange(1, self.n_cav)), -1).astype(C128)

    def annihilate(self):
        return np.diag(np.sqrt(np.arange(1, self.n_cav)), 1).astype(C128)

    def get_number_operator(self):
        return np.diag(np.arange(self.n_cav)).astype(C128)

    def num(self):
        return self.get_number_operator()

    def quadrature(self):
        """a + a^dagger, the displacement the second bath of config 4 couples to."""
        a = self.annihilate()
        return a + a.conj().T

    def vacuum(self):
        v = np.zeros(self.n_cav, dtype=C128)
        v[0] = 1.0
        return v

    def get_dm(self):
        v = self.vacuum()
        return np.outer(v, v.conj())

    vacuum_dm = get_dm


class Polariton(Mol):
    """Molecule x cavity (``cavity.py:577-678``); basis ordering |mol> (x) |n>."""

    def __init__(self, mol, cav, g=None, gauge="length"):
        if not isinstance(mol, Mol):
            raise TypeError("mol must be a Mol")
        if gauge not in ("length", "dipole", "dip"):
            raise NotImplementedError("only the length (dipole) gauge is provided")
        self.mol, self.cav = mol, cav
        self.dims = [mol.dim, cav.n_cav]
        self.gauge = gauge
        self._g = g
        dim = mol.dim * cav.n_cav
        super().__init__(np.zeros((dim, dim), dtype=C128),
                         edip=np.kron(mol.edip, cav.idm))
        self.H = None

    @property
    def g(self):
        return self._g

    @g.setter
    def g(self, value):
        self._g = value

    def getH(self, RWA=False):
        """``H_mol (x) 1 + 1 (x) H_cav + H_int`` with, in the length gauge,
        ``H_int = g (sigma^+ a + sigma^- a^dagger)`` under the RWA and
        ``i g mu (a - a^dagger) + g^2 / omega_c mu^2`` (dipole self-energy)
        otherwise (``cavity.py:650-662``)."""
        if self._g is None:
            raise ValueError("set the coupling strength g first")
        mol, cav, g = self.mol, self.cav, self._g
        a = cav.annihilate()
        ad = a.conj().T
        if RWA:
            if getattr(mol, "lowering", None) is None:
                raise ValueError("the RWA coupling needs Mol(..., lowering=...)")
            hint = g * (np.kron(mol.raising, a) + np.kron(mol.lowering, ad))
        else:
            hint = 1j * g * np.kron(mol.edip, a - ad) + g ** 2 / cav.omegac * np.kron(mol.edip @ mol.edip, cav.idm)
        self.H = np.kron(mol.getH(), cav.idm) + np.kron(mol.idm, cav.getH()) + hint
        return self.H

    get_ham = getH

    def setH(self, h):
        self.H = np.asarray(h, dtype=C128)

    def promote_op(self, a, kind="mol"):
        """Operator of one subsystem on the composite space (``cavity.py:804-827``)."""
        a = np.asarray(a, dtype=C128)
        if kind in ("mol", "m"):
            return np.kron(a, self.cav.idm)
        if kind in ("cav", "c"):
            return np.kron(self.mol.idm, a)
        raise ValueError("kind must be 'mol' or 'cav'")

    def get_dm(self, mol_dm=None):
        """Product state: molecular density matrix (default: first basis state)
        times the cavity vacuum."""
        if mol_dm is None:
            mol_dm = np.zeros((self.mol.dim, self.mol.dim), dtype=C128)
            mol_dm[0, 0] = 1.0
        return np.kron(np.asarray(mol_dm, dtype=C128), self.cav.get_dm())

    def deom(self, bath, coupling, coupling_dipole=None, pulse_system_func=None,
             pulse_coupling_func=None, mode=None, **solver_kwargs):
        if self.H is None:
            self.getH()
        return super().deom(bath, coupling, coupling_dipole, pulse_system_func,
                            pulse_coupling_func, mode, **solver_kwargs)


class Env:
    """High-temperature Drude environment, one bath per collapse operator
    (``exact.py:41-50``): ``temperature`` (kT), ``cutoff`` = [gamma_i], ``reorg`` = [lambda_i]."""

    def __init__(self, temperature, cutoff, reorg):
        self.temperature = temperature
        self.gamma = cutoff
        self.reorg = reorg
        self.c_ops = None

    def set_c_ops(self, c_ops):
        self.c_ops = c_ops


def env_bath(env):
    """``Bath`` equivalent of ``theta`` (``exact.py:660-671``): bath i contributes
    one exponent ``gamma_i`` with ``eta_i = lambda_i (2 kT - i gamma_i)``, since
    ``theta_i(rho) = i (eta_i c rho - conj(eta_i) rho c)``."""
    from .heom.bath import Bath
    gam = np.atleast_1d(np.asarray(env.gamma, dtype=float))
    lam = np.atleast_1d(np.asarray(env.reorg, dtype=float))
    if gam.shape != lam.shape:
        raise ValueError("cutoff and reorg must have one entry per collapse operator")
    eta = lam * (2.0 * float(env.temperature) - 1j * gam)
    return Bath(expn=gam.astype(C128), etal=eta, etar=eta.conj(), etaa=np.abs(eta).astype(C128),
                mode=np.arange(len(gam)))


def heom(env, hs, rho, obs_ops, Nt, dt, lmax=4, **solver_kwargs):
    """Two-bath (or any-number-of-bath) high-temperature HEOM of ``exact.py:674``,
    propagated with RK4 by the DEOM kernel.  Returns a ``Result`` whose
    ``observables[i, j] = Tr(obs_ops[j] rho(t_i))``, ``t_i = i dt``, ``i = 0..Nt``,
    and whose ``rho`` is the final reduced density matrix (what the reference
    function returns)."""
    from .heom.deom import DEOMSolver
    if env.c_ops is None:
        raise ValueError("env.set_c_ops(...) first")
    hs = np.asarray(hs, dtype=C128)
    c_ops = [np.asarray(c, dtype=C128) for c in env.c_ops]
    solver = DEOMSolver(hs, np.zeros_like(hs), env_bath(env), c_ops, lmax=lmax, **solver_kwargs)
    rho0 = np.array(rho, dtype=C128)
    _, traj = solver.run(rho0.copy(), dt, Nt)
    traj = np.asarray(traj)
    res = Result(description="HEOM (high-temperature Drude baths, RK4)", rho0=rho0, dt=dt, Nt=Nt)
    res.rholist = traj
    res.rho = traj[-1]
    ops = np.stack([np.asarray(o, dtype=C128) for o in obs_ops]) if len(obs_ops) else np.zeros((0,) + hs.shape)
    res.observables = np.einsum("oij,tji->to", ops, traj)
    res.solver = solver
    return res
