"""Spectral-density decompositions with the reference's call signatures.

``decompose_spectrum_pade(spe, w_sp, beta, npsd, pade=1)`` takes a sympy
expression for J(w) like the reference (``pyqed/heom/deom.py:226-307``) but only
uses sympy to read off the numerator/denominator polynomials; the exponents are
then computed numerically by ``bath.rational_exponents``.  Return order is the
reference's: ``(etal, etar, etaa, expn)``.
"""
from __future__ import annotations

import numpy as np

from .bath import (rational_exponents, rational_residues, fit_t, prony_fitting,
                   single_oscillator as _single_oscillator)


def _poly_ratio(spe, w_sp):
    import sympy as sp
    numer, denom = sp.fraction(sp.cancel(sp.together(spe)))
    pn = sp.Poly(sp.expand(numer), w_sp)
    pd = sp.Poly(sp.expand(denom), w_sp)
    return ([complex(c) for c in pn.all_coeffs()], [complex(c) for c in pd.all_coeffs()])


def decompose_spectrum_pade(spe, w_sp, beta, npsd, pade=1, bose_fermi=1):
    if bose_fermi != 1:
        raise ValueError("only bosonic baths are supported")
    numer, denom = _poly_ratio(spe, w_sp)
    expn, etal, etar, etaa = rational_exponents(numer, denom, beta, npsd, pade)
    return etal, etar, etaa, expn


def decompose_spectrum_matsubara(spe, w_sp, beta, npsd):
    """Convenience wrapper: ``pade=0`` (the wrapper SURVEY.md section 8b says the
    reference needs in order to use Matsubara terms through ``Bath``)."""
    return decompose_spectrum_pade(spe, w_sp, beta, npsd, pade=0)


def single_oscillator(omega, w_sp, beta, nind):
    """Reference signature (``deom.py:545``); ``w_sp`` and ``nind`` are unused
    there too."""
    return _single_oscillator(omega, beta)


def decompose_spectrum_pade_real(spe, w_sp):
    """Poles of J(w) without the Bose factor (``deom.py:310-367``): ``(etal, etar, etaa, expn)``."""
    return rational_residues(*_poly_ratio(spe, w_sp))


def decompose_spectrum_pade_imag(spe, w_sp):
    """Same construction as ``decompose_spectrum_pade_real`` (``deom.py:370-425``)."""
    return rational_residues(*_poly_ratio(spe, w_sp))


def decompose_spectrum_prony(spe, w_sp, beta, nind, scale=250000, n=1250, npsd=10, bose_fermi=1):
    """Prony fit of the bath correlation function (``deom.py:507-543``): C(t) is sampled from a
    Pade decomposition with ``npsd`` terms on ``2n+1`` points over ``[0, scale]`` and refitted
    with ``nind`` exponentials (an int: all from Re C; ``[n_real, n_imag]``: from Re C and Im C;
    ``[n_real, 'a']``: the imaginary part's exponents taken analytically from the poles of J).
    Returns ``(etal, etar, etaa, expn)`` as ``sort_symmetry`` orders them."""
    if bose_fermi != 1:
        raise ValueError("only bosonic baths are supported")
    etal_p, _, _, expn_p = decompose_spectrum_pade(spe, w_sp, beta, npsd)
    t = np.linspace(0, 1, 2 * n + 1)
    samples = fit_t(scale * t, expn_p, etal_p)
    if isinstance(nind, list):
        nind = list(nind)
        if nind[0] == 'a':
            # the reference stops the interpreter here (exit(), deom.py:529-531)
            raise ValueError("for a bosonic bath only the imaginary part of C(t) is analytic: use [n, 'a']")
        if nind[1] == 'a':
            _, _, _, expn_imag = decompose_spectrum_pade_imag(spe, w_sp)
            gamma_imag = np.exp(-expn_imag * scale / (2 * n))
            nind[1] = len(gamma_imag)
            return prony_fitting(samples, t, nind, scale, n, gamma_imag=gamma_imag)
    return prony_fitting(samples, t, nind, scale, n)
