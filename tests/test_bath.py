"""Bath decomposition against the reference's outputs.

The reference's ``decompose_spectrum_pade`` evaluates ``numer / prod(w - poles)``
(``pyqed/heom/deom.py:255-305``), i.e. it drops the leading coefficient of the
denominator sympy hands it.  For monic denominators (every example in the
reference: gamma = 1) the two agree to rounding; otherwise the reference's
``etal/etar`` are the correct ones times that leading coefficient (and ``etaa``
likewise), which is what the scaled comparison below checks.
"""
import numpy as np
import pytest
import sympy as sp

from conftest import golden
from pyqed_b200.heom import bath as B
from pyqed_b200.heom.spectrum import decompose_spectrum_pade, single_oscillator

w = sp.symbols("omega", real=True)
CASES = {
    "drude_m1": (2 * 0.2 * 1.0 * w / (1.0 ** 2 + w ** 2), 1.0, 1, 0),
    "drude_m2": (2 * 6.593 * 20.0 * w / (20.0 ** 2 + w ** 2), 1 / 39.276, 2, 0),
    "drude_p1": (2 * 0.05 * 1.0 * w / (1.0 ** 2 + w ** 2), 1.0, 1, 1),
    "drude_p2": (2 * 1.0 * 1.0 * w / (1.0 ** 2 + w ** 2), 1.0, 2, 1),
    "drude_p5": (2 * 0.5 * 2.0 * w / (2.0 ** 2 + w ** 2), 0.7, 5, 1),
    "bo_p3": (2 * 0.3 * 0.4 * 1.5 ** 2 * w / ((w ** 2 - 1.5 ** 2) ** 2 + 0.4 ** 2 * w ** 2), 0.8, 3, 1),
}


def _reference_lead(spe):
    """Leading denominator coefficient in the normal form the reference uses."""
    _, denom = sp.cancel(sp.factor(spe)).as_numer_denom()
    return complex(sp.Poly(denom, w).all_coeffs()[0])


@pytest.mark.parametrize("n", [1, 2, 3, 6])
def test_pade_poles(n):
    g = golden("bath")
    p, r = B.bose_poles(n, 1)
    assert np.allclose(p, g[f"psd_pole_{n}"], rtol=1e-12, atol=0)
    assert np.allclose(r, g[f"psd_resi_{n}"], rtol=1e-11, atol=0)


@pytest.mark.parametrize("name", sorted(CASES))
def test_decomposition_matches_reference(name):
    g = golden("bath")
    spe, beta, npsd, pade = CASES[name]
    etal, etar, etaa, expn = decompose_spectrum_pade(spe, w, beta, npsd, pade=pade)
    lead = _reference_lead(spe)
    assert np.allclose(expn, g[f"{name}_expn"], rtol=1e-12, atol=1e-13)
    assert np.allclose(etal * lead, g[f"{name}_etal"], rtol=1e-11, atol=1e-14)
    assert np.allclose(etar * np.conj(lead), g[f"{name}_etar"], rtol=1e-11, atol=1e-14)
    assert np.allclose(etaa * abs(lead), g[f"{name}_etaa"], rtol=1e-11, atol=1e-14)
    if abs(lead - 1) < 1e-14:  # monic: identical to the reference without any scaling
        assert np.allclose(etal, g[f"{name}_etal"], rtol=1e-11, atol=1e-14)


def test_drude_closed_form():
    lam, gam, beta = 0.2, 1.0, 1.0
    expn, etal, etar, etaa = B.drude_exponents(lam, gam, beta, 1, 0)
    nu = 2 * np.pi / beta
    eta0 = 2 * lam / beta - 1j * lam * gam - 4 * lam * gam * beta * gam / (nu ** 2 * beta ** 2 - beta ** 2 * gam ** 2)
    assert np.allclose(expn, [gam, nu])
    assert np.allclose(etal, [eta0, 4 * lam * gam * nu / (beta * (nu ** 2 - gam ** 2))], rtol=1e-13)
    assert np.allclose(etar, np.conj(etal)) and np.allclose(etaa, np.abs(etal))
    # SURVEY.md section 8b records these values from the reference
    assert np.allclose(etal, [0.37920912 - 0.2j, 0.13063293], atol=1e-8)


def test_single_oscillator_and_bath_container():
    g = golden("bath")
    etal, etar, etaa, expn = single_oscillator(1.3, w, 0.9, 2)
    for got, key in ((etal, "so_etal"), (etar, "so_etar"), (etaa, "so_etaa"), (expn, "so_expn")):
        assert np.allclose(got, g[key], rtol=1e-14)
    b = B.Bath([CASES["drude_p2"][0]], w, [1.0], [2], [0, 0, 0])
    assert np.allclose(b.etal, g["drude_p2_etal"], rtol=1e-11)
    assert b.mode.tolist() == [0, 0, 0] and b.expn.dtype == np.complex128
    with pytest.raises(ValueError):
        B.Bath([CASES["drude_p2"][0]], w, [1.0], [2], [0, 0])
    with pytest.raises(ValueError):
        B.Bath([CASES["drude_p2"][0]], w, [1.0], [2], None)
    single = B.Bath(CASES["drude_p1"][0], w, 1.0, 1)
    assert np.allclose(single.etal, g["drude_p1_etal"], rtol=1e-11) and single.mode.tolist() == [0, 0]
