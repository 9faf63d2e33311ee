"""Bath correlation-function exponents for the DEOM/HEOM solvers (host side).

Mirrors the reference's ``Bath`` container and spectrum decompositions
(``pyqed/heom/deom.py:84-307, 545-552, 895-942``): a bath is five arrays
``expn, etal, etar, etaa`` (complex128[K]) and ``mode`` (int[K]) such that
``C(t) = sum_k etal_k exp(-expn_k t)``.  This is host-only, O(ms) work done
once per run; the device path only ever sees the five arrays.

Implemented here from the published formulas rather than by symbolic algebra:

* ``bose_poles``      - Matsubara poles (``deom.py:84-102``) and the [N-1/N] Pade
  spectrum decomposition of the Bose function (Hu, Xu, Yan, JCP 133, 101106
  (2010); ``deom.py:104-163`` with ``pade=1``).
* ``rational_exponents`` - exponents of any spectral density given as a ratio
  of polynomials in omega (what ``decompose_spectrum_pade`` does via sympy,
  ``deom.py:226-307``).
* ``drude_exponents`` - the Drude-Lorentz special case in closed form.
* ``single_oscillator`` - ``deom.py:545-552``.
"""
from __future__ import annotations

import numpy as np

C128 = np.complex128


def bose_poles(n: int, pade: int = 1):
    """Poles ``p_j`` and residues ``r_j`` of
    ``1/(1-exp(-x)) ~ 1/x + 1/2 + sum_j 2 r_j x / (x^2 + p_j^2)``.

    ``pade=0``: Matsubara, ``p_j = 2 pi j``, ``r_j = 1``.  ``pade=1``: [N-1/N] PSD.
    """
    if n < 0 or pade not in (0, 1):
        raise ValueError("N or BoseFermi or pade has wrong value!")
    if n == 0:
        return np.zeros(0), np.zeros(0)
    if pade == 0:
        return 2.0 * np.pi * np.arange(1, n + 1), np.ones(n)

    def spectrum(size, first):
        b = first + 2.0 * np.arange(size)
        off = 1.0 / np.sqrt(b[:-1] * b[1:])
        mat = np.diag(off, 1) + np.diag(off, -1)
        ev = np.sort(np.linalg.eigvalsh(mat))[::-1]
        return ev

    xi = 2.0 / spectrum(2 * n, 3.0)[:n]
    zeta = 2.0 / spectrum(2 * n - 1, 5.0)[:n - 1]
    xi2, zeta2 = xi * xi, zeta * zeta
    resi = np.zeros(n)
    for j in range(n):
        val = 0.5 * n * (2.0 * n + 3.0)
        for k in range(n - 1):
            val *= (zeta2[k] - xi2[j])
        for k in range(n):
            if k != j:
                val /= (xi2[k] - xi2[j])
        resi[j] = val
    return xi, resi


def _bose_approx(x, pole, resi):
    """``function_bose`` of ``deom.py:66-71``."""
    return 1.0 / x + 0.5 + sum(2.0 * r * x / (x * x + p * p) for p, r in zip(pole, resi))


def rational_exponents(numer, denom, beta, npsd, pade=1):
    """Exponents for ``J(w) = numer(w) / denom(w)`` (polynomial coefficient
    arrays, highest power first, as ``numpy.polyval`` takes them).

    Follows the construction of ``decompose_spectrum_pade``
    (``pyqed/heom/deom.py:226-307``): every pole ``z`` of ``J`` in the lower half
    plane gives ``expn = i z`` with ``etal = -2i Res-like(J, z) f_Bose(z beta)``;
    conjugate pole pairs come first (sorted by decreasing ``|Im expn|``), then
    the purely damped ones, then one term per Bose pole.  ``etar`` of a conjugate
    pair is the conjugate of the partner's ``etal``; ``etaa = sqrt(|etal| |etar|)``.
    """
    numer = np.atleast_1d(np.asarray(numer, dtype=C128))
    denom = np.atleast_1d(np.asarray(denom, dtype=C128))
    lead = denom[0]
    poles = np.roots(denom)
    temp = 1.0 / beta
    pole_b, resi_b = bose_poles(npsd, pade)

    # the reference takes the poles in sympy.nroots order (real part, then
    # imaginary part, ascending) and then sorts by |Im expn| descending with a
    # reversed argsort, which fixes the order inside a conjugate pair
    scale = max(1.0, float(np.max(np.abs(poles)))) if len(poles) else 1.0
    poles = np.array(sorted(poles, key=lambda z: (round(z.real / scale, 10), z.imag)))
    lower = [z for z in poles if z.imag < 0]
    expn_sys = np.array([1j * z for z in lower], dtype=C128)
    order = np.argsort(np.abs(expn_sys.imag), kind="stable")[::-1]
    expn_sys = expn_sys[order]
    paired = [e for e in expn_sys if abs(e.imag) > 1e-12 * max(1.0, abs(e))]
    single = [e for e in expn_sys if not abs(e.imag) > 1e-12 * max(1.0, abs(e))]
    single = [complex(e.real, 0.0) for e in single]

    def coeff(e):
        z0 = -1j * e
        rest = np.prod([z0 - p for p in poles if abs(p - z0) > 1e-12 * max(1.0, abs(z0))])
        return -2j * np.polyval(numer, z0) / (lead * rest) * _bose_approx(z0 / temp, pole_b, resi_b)

    expn, etal, etar, etaa = [], [], [], []
    for i in range(0, len(paired) - 1, 2):
        a, b = coeff(paired[i]), coeff(paired[i + 1])
        expn += [paired[i], paired[i + 1]]
        etal += [a, b]
        etar += [np.conj(b), np.conj(a)]
        etaa += [np.sqrt(abs(a) * abs(np.conj(b))), np.sqrt(abs(b) * abs(np.conj(a)))]
    for e in single:
        a = coeff(e)
        expn.append(e)
        etal.append(a)
        etar.append(np.conj(a))
        etaa.append(abs(a))
    for p, r in zip(pole_b, resi_b):
        z = -1j * p * temp
        a = -2j * r * temp * np.polyval(numer, z) / np.polyval(denom, z)
        expn.append(p * temp)
        etal.append(a)
        etar.append(np.conj(a))
        etaa.append(abs(a))
    return (np.array(expn, C128), np.array(etal, C128), np.array(etar, C128),
            np.array(etaa, C128))


def drude_exponents(lam, gam, beta, npsd, pade=1):
    """Drude-Lorentz ``J(w) = 2 lam gam w / (gam^2 + w^2)``: one damped pole at
    ``gam`` plus ``npsd`` Bose-function terms.  Returns
    ``(expn, etal, etar, etaa)``, each complex128[1 + npsd]."""
    return rational_exponents([2.0 * lam * gam, 0.0], [1.0, 0.0, gam * gam], beta, npsd, pade)


def single_oscillator(omega, beta):
    """Undamped mode of frequency ``omega`` (``deom.py:545-552``); returns
    ``(etal, etar, etaa, expn)`` in the reference's order."""
    etal = np.array([1 / (2 * (1 - np.exp(-beta * omega))),
                     -1 / (2 * (1 - np.exp(beta * omega)))], dtype=C128)
    etar = np.array([-1 / (2 * (1 - np.exp(beta * omega))),
                     1 / (2 * (1 - np.exp(-beta * omega)))], dtype=C128)
    etaa = np.sqrt(np.abs(etal + etar)).astype(C128)
    expn = np.array([1j * omega, -1j * omega], dtype=C128)
    return etal, etar, etaa, expn


class Bath:
    """Container the solver reads ``expn, etal, etar, etaa, mode`` from.

    Two ways to build it:

    * ``Bath(expn=..., etal=..., etar=..., etaa=..., mode=...)`` - arrays
      directly (the reference solver is duck-typed on exactly these five
      attributes, ``deom.py:1040-1041, 1070``).
    * ``Bath(spectrum_sp, w_sp, beta, npsd, mode, function)`` - the reference's
      signature (``deom.py:900``); see ``pyqed_b200.heom.spectrum`` for the
      decomposition functions that accept symbolic spectral densities.
    """

    def __init__(self, spectrum_sp=None, w_sp=None, beta=None, npsd=None, mode=None,
                 function=None, *, expn=None, etal=None, etar=None, etaa=None):
        self.bath, self.w_sp, self.beta, self.npsd = spectrum_sp, w_sp, beta, npsd
        if expn is not None:
            self.expn = np.array(expn, dtype=C128)
            self.etal = np.array(etal, dtype=C128)
            self.etar = np.array(np.conj(self.etal) if etar is None else etar, dtype=C128)
            self.etaa = np.array(np.abs(self.etal) if etaa is None else etaa, dtype=C128)
            self.mode = (np.zeros(len(self.expn), np.int64) if mode is None
                         else np.array(mode, dtype=np.int64))
            if len(self.mode) != len(self.expn):
                raise ValueError("the length of mode is not equal to the number of dissipatons!")
            return
        from .spectrum import decompose_spectrum_pade
        if isinstance(spectrum_sp, list) and isinstance(beta, list) and isinstance(npsd, list):
            if function is None:
                function = [decompose_spectrum_pade] * len(spectrum_sp)
            if len(spectrum_sp) != len(beta) or len(spectrum_sp) != len(npsd):
                raise ValueError("the length of bath, w_sp, beta, npsd is not equal!")
            parts = [function[i](spectrum_sp[i], w_sp, beta[i], npsd[i])
                     for i in range(len(spectrum_sp))]
            self.etal = np.concatenate([np.atleast_1d(p[0]) for p in parts]).astype(C128)
            self.etar = np.concatenate([np.atleast_1d(p[1]) for p in parts]).astype(C128)
            self.etaa = np.concatenate([np.atleast_1d(p[2]) for p in parts]).astype(C128)
            self.expn = np.concatenate([np.atleast_1d(p[3]) for p in parts]).astype(C128)
            if mode is None:
                raise ValueError("mode is not set!")
            if len(mode) != len(self.expn):
                raise ValueError("the length of mode is not equal to the number of dissipatons!")
            self.mode = np.array(mode, dtype=np.int64)
        else:
            etal, etar, etaa, expn = decompose_spectrum_pade(spectrum_sp, w_sp, beta, npsd)
            self.etal, self.etar, self.etaa, self.expn = (
                np.array(x, dtype=C128) for x in (etal, etar, etaa, expn))
            self.mode = np.zeros(len(self.expn), dtype=np.int64)
