"""NumPy restatement of the reference's DEOM (multi-exponential HEOM) RK4 path.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  Parity status: pinned
against ``tests/golden/*.npz`` (outputs of the unmodified reference).

Every function cites the reference lines it restates (paths relative to
``/root/reference``).  Two evaluators of the same right-hand side are provided:

* ``rhs_loop``  - one ADO at a time, same operation order as the reference's
  ``generate_dot_element`` (``pyqed/heom/deom.py:641-664``); this is the form
  timed as the CPU baseline because it has the reference's cost structure.
* ``rhs_batched`` - all ADOs at once with batched matmuls; same equations,
  same k-ordering of the link sums, used for mid-size parity checks where the
  loop form would take minutes.
"""
from __future__ import annotations

from math import comb

import numpy as np

__all__ = [
    "pascal_table", "ado_id", "build_keys", "build_neighbours", "DeomOracle",
]


# --------------------------------------------------------------------------
# hierarchy indexing
# --------------------------------------------------------------------------
def pascal_table(nind: int, lmax: int) -> np.ndarray:
    """``comb_list[a, b] = C(a, b)`` on a square of side ``nind+lmax+1``.

    Restates ``DEOMSolver.init_`` (``pyqed/heom/deom.py:1048-1059``).
    """
    side = nind + lmax + 1
    tab = np.zeros((side, side), dtype=np.int64)
    for a in range(side):
        for b in range(a + 1):
            tab[a, b] = comb(a, b)
    return tab


def ado_id(key, tab) -> int:
    """Flat id of a multi-index: ``sum_i C(s_i + i, i + 1)``, ``s_i`` the
    running sum of ``key`` (``gen_hash_value``, ``pyqed/heom/deom.py:555-565``)."""
    run = 0
    out = 0
    for i, n in enumerate(key):
        run += int(n)
        out += int(tab[run + i, i + 1])
    return out


def _ids_of(keys: np.ndarray, tab: np.ndarray) -> np.ndarray:
    """Vectorised ``ado_id`` over rows of ``keys``."""
    run = np.cumsum(keys, axis=1)
    cols = np.arange(keys.shape[1])
    return tab[run + cols[None, :], cols[None, :] + 1].sum(axis=1)


def build_keys(nind: int, lmax: int, tab: np.ndarray | None = None) -> np.ndarray:
    """``keys[id] = multi-index`` for every ``|n| <= lmax``.

    The reference fills this table by a breadth-first walk over tiers
    (``gen_keys`` / ``gen_keys_element``, ``pyqed/heom/deom.py:608-638``); since
    the id is a bijection onto ``[0, C(lmax+nind, lmax))`` the table is simply the
    inverse of ``ado_id``, which is how it is produced here.
    """
    if tab is None:
        tab = pascal_table(nind, lmax)
    cur = np.zeros((1, 0), dtype=np.int64)
    for _ in range(nind):
        used = cur.sum(axis=1)
        parts = []
        for v in range(lmax + 1):
            keep = cur[used + v <= lmax]
            parts.append(np.concatenate(
                [keep, np.full((keep.shape[0], 1), v, dtype=np.int64)], axis=1))
        cur = np.concatenate(parts, axis=0)
    nmax = int(tab[lmax + nind, lmax])
    assert cur.shape[0] == nmax
    keys = np.zeros((nmax, nind), dtype=np.int64)
    keys[_ids_of(cur, tab)] = cur
    return keys


def build_neighbours(keys: np.ndarray, lmax: int, tab: np.ndarray):
    """ids of ``n - e_k`` and ``n + e_k`` (``-1`` where the reference skips the
    term: ``n_k == 0`` resp. ``|n| == lmax``; ``hash_minus`` / ``hash_plus``,
    ``pyqed/heom/deom.py:588-605`` as used at ``:653-664``)."""
    nmax, nind = keys.shape
    minus = np.full((nmax, nind), -1, dtype=np.int64)
    plus = np.full((nmax, nind), -1, dtype=np.int64)
    tier = keys.sum(axis=1)
    for k in range(nind):
        sel = keys[:, k] > 0
        kk = keys[sel].copy()
        kk[:, k] -= 1
        minus[sel, k] = _ids_of(kk, tab)
        sel = tier < lmax
        kk = keys[sel].copy()
        kk[:, k] += 1
        plus[sel, k] = _ids_of(kk, tab)
    return minus, plus


# --------------------------------------------------------------------------
# propagation
# --------------------------------------------------------------------------
class DeomOracle:
    """State holder mirroring what ``DEOMSolver.run`` keeps
    (``pyqed/heom/deom.py:1072-1114``), with arrays instead of lists.

    Parameters follow the reference constructor (``deom.py:958``): ``system``
    N x N, ``system_dipole`` N x N, ``coupling`` M x N x N, ``coupling_dipole``
    broadcastable to M x N x N, the five bath arrays, and two callables
    ``t -> float``.
    """

    def __init__(self, system, system_dipole, coupling, coupling_dipole,
                 expn, etal, etar, etaa, mode, lmax,
                 pulse_system_func=None, pulse_coupling_func=None):
        c128 = np.complex128
        self.H0 = np.array(system, dtype=c128)
        n = self.H0.shape[0]
        self.mu = (np.zeros((n, n), c128) if system_dipole is None
                   else np.array(system_dipole, dtype=c128))
        self.Q0 = np.array(coupling, dtype=c128)
        nmod = self.Q0.shape[0]
        if coupling_dipole is None:
            self.Qd = np.zeros_like(self.Q0)
        else:
            qd = np.array(coupling_dipole, dtype=c128)
            # deom.py:685-686 indexes coupling_dip[i] and lets numpy broadcast
            self.Qd = np.stack([np.broadcast_to(qd[i], (n, n)) for i in range(nmod)])
        self.expn = np.array(expn, dtype=c128)
        self.etal = np.array(etal, dtype=c128)
        self.etar = np.array(etar, dtype=c128)
        self.etaa = np.array(etaa, dtype=c128)
        self.mode = np.array(mode, dtype=np.int64)
        self.lmax = int(lmax)
        self.f = pulse_system_func or (lambda t: 0.0)
        self.g = pulse_coupling_func or (lambda t: 0.0)
        self.nsys = n
        self.nind = len(self.expn)
        self.tab = pascal_table(self.nind, self.lmax)
        self.nmax = int(self.tab[self.lmax + self.nind, self.lmax])
        self.keys = build_keys(self.nind, self.lmax, self.tab)
        self.minus, self.plus = build_neighbours(self.keys, self.lmax, self.tab)
        self.ddos = None

    # -- operators at time t (generate_time, deom.py:676-687) ---------------
    def operators_at(self, t):
        return self.H0 + self.mu * self.f(t), self.Q0 + self.Qd * self.g(t)

    # -- right-hand side, reference cost structure (deom.py:641-673) --------
    def rhs_loop(self, rho, t):
        H, Q = self.operators_at(t)
        out = np.empty_like(rho)
        for n in range(self.nmax):
            key = self.keys[n]
            tier = int(key.sum())
            acc = -np.sum(key * self.expn) * rho[n]
            acc = acc - 1j * (H @ rho[n] - rho[n] @ H)
            for k in range(self.nind):
                nk = key[k]
                q = Q[self.mode[k]]
                if nk > 0:
                    src = rho[self.minus[n, k]]
                    acc = acc - 1j * np.sqrt(nk) / np.sqrt(self.etaa[k]) * (
                        self.etal[k] * q @ src - self.etar[k] * src @ q)
                if tier < self.lmax:
                    src = rho[self.plus[n, k]]
                    acc = acc - 1j * np.sqrt(nk + 1) * np.sqrt(self.etaa[k]) * (
                        q @ src - src @ q)
            out[n] = acc
        return out

    # -- same equations, all ADOs per numpy call ---------------------------
    def rhs_batched(self, rho, t):
        H, Q = self.operators_at(t)
        damp = (self.keys * self.expn[None, :]).sum(axis=1)
        out = -damp[:, None, None] * rho
        out -= 1j * (H @ rho - rho @ H)
        for k in range(self.nind):
            q = Q[self.mode[k]]
            sel = np.nonzero(self.minus[:, k] >= 0)[0]
            if sel.size:
                src = rho[self.minus[sel, k]]
                c = (1j * np.sqrt(self.keys[sel, k]) / np.sqrt(self.etaa[k]))[:, None, None]
                out[sel] -= c * (self.etal[k] * (q @ src) - self.etar[k] * (src @ q))
            sel = np.nonzero(self.plus[:, k] >= 0)[0]
            if sel.size:
                src = rho[self.plus[sel, k]]
                c = (1j * np.sqrt(self.keys[sel, k] + 1) * np.sqrt(self.etaa[k]))[:, None, None]
                out[sel] -= c * (q @ src - src @ q)
        return out

    # -- classical RK4, stage times t, t+dt/2, t+dt/2, t+dt (deom.py:725-766)
    def rk4_step(self, rho, dt, t, rhs):
        k = rhs(rho, t)
        acc = k
        k = rhs(rho + k * dt / 2, t + dt / 2)
        acc = acc + k * 2
        k = rhs(rho + k * dt / 2, t + dt / 2)
        acc = acc + k * 2
        k = rhs(rho + k * dt, t + dt)
        acc = acc + k
        return rho + acc * dt / 6

    # -- DEOMSolver.run (deom.py:1072-1114) ----------------------------------
    def run(self, rho0, dt, nt, p1=None, batched=True, e_ops=None):
        """Returns ``(t_save, ddos_save)`` like the reference: ``ddos_save`` is a
        list of nt+1 system density matrices, or complex[nt+1] of
        ``trace(p1 @ rho_sys)`` when ``p1`` is given.  All ADOs stay in
        ``self.ddos`` afterwards."""
        rhs = self.rhs_batched if batched else self.rhs_loop
        rho = np.zeros((self.nmax, self.nsys, self.nsys), dtype=np.complex128)
        rho[0] = np.array(rho0, dtype=np.complex128)
        t_save = np.zeros(nt + 1)
        if p1 is None:
            save = [rho[0].copy()]
        else:
            p1 = np.array(p1, dtype=np.complex128)
            save = np.zeros(nt + 1, dtype=np.complex128)
            save[0] = np.trace(p1 @ rho[0])
        for i in range(nt):
            rho = self.rk4_step(rho, dt, i * dt, rhs)
            t_save[i + 1] = (i + 1) * dt
            if p1 is None:
                save.append(rho[0].copy())
            else:
                save[i + 1] = np.trace(p1 @ rho[0])
        self.ddos = rho
        return t_save, save
