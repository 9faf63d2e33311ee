"""Spectral-density decompositions with the reference's call signatures.

``decompose_spectrum_pade(spe, w_sp, beta, npsd, pade=1)`` takes a sympy
expression for J(w) like the reference (``pyqed/heom/deom.py:226-307``) but only
uses sympy to read off the numerator/denominator polynomials; the exponents are
then computed numerically by ``bath.rational_exponents``.  Return order is the
reference's: ``(etal, etar, etaa, expn)``.
"""
from __future__ import annotations

import numpy as np

from .bath import rational_exponents, single_oscillator as _single_oscillator


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
