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


# ---- Pade [N/N] and [N+1/N], poles without the Bose factor, Prony refits (deom.py:119-207, 310-543) ----
@pytest.mark.parametrize("pade", [2, 3])
@pytest.mark.parametrize("n", [1, 2, 3, 6])
def test_higher_pade_poles(n, pade):
    g = golden("bath")
    p, r = B.bose_poles(n, pade)
    assert np.allclose(p, g[f"psd{pade}_pole_{n}"], rtol=1e-12, atol=0)
    assert np.allclose(r, g[f"psd{pade}_resi_{n}"], rtol=1e-11, atol=0)


def test_pade_argument_checks():
    for bad in ((-1, 1), (2, 4), (2, -1)):
        with pytest.raises(ValueError):
            B.bose_poles(*bad)
    assert B.bose_poles(0, 2)[0].size == 0


@pytest.mark.parametrize("name,case,npsd,pade,beta", [("drude_q2", "drude_p2", 2, 2, 1.0), ("drude_q3", "drude_p2", 3, 3, 1.0),
                                                      ("bo_q2", "bo_p3", 3, 2, 0.8), ("bo_q3", "bo_p3", 2, 3, 0.8)])
def test_higher_pade_decomposition_matches_reference(name, case, npsd, pade, beta):
    g = golden("bath")
    spe = CASES[case][0]
    etal, etar, etaa, expn = decompose_spectrum_pade(spe, w, beta, npsd, pade=pade)
    lead = _reference_lead(spe)
    assert np.allclose(expn, g[f"{name}_expn"], rtol=1e-12, atol=1e-13)
    assert np.allclose(etal * lead, g[f"{name}_etal"], rtol=1e-10, atol=1e-14)
    assert np.allclose(etar * np.conj(lead), g[f"{name}_etar"], rtol=1e-10, atol=1e-14)
    assert np.allclose(etaa * abs(lead), g[f"{name}_etaa"], rtol=1e-10, atol=1e-14)


def test_poles_without_bose_factor():
    from pyqed_b200.heom.spectrum import decompose_spectrum_pade_real, decompose_spectrum_pade_imag
    g = golden("bath")
    for name, fn, case in (("real_drude", decompose_spectrum_pade_real, "drude_p2"),
                           ("imag_bo", decompose_spectrum_pade_imag, "bo_p3")):
        spe = CASES[case][0]
        lead = _reference_lead(spe)
        etal, etar, etaa, expn = fn(spe, w)
        assert np.allclose(expn, g[f"{name}_expn"], rtol=1e-12, atol=1e-13)
        assert np.allclose(etal * lead, g[f"{name}_etal"], rtol=1e-11, atol=1e-14)
        assert np.allclose(etar * np.conj(lead), g[f"{name}_etar"], rtol=1e-11, atol=1e-14)
        assert np.allclose(etaa * abs(lead), g[f"{name}_etaa"], rtol=1e-11, atol=1e-14)


def _canonical(t):
    etal, etar, etaa, expn = (np.asarray(x) for x in t)
    o = np.lexsort((np.round(expn.imag, 8), np.round(expn.real, 8)))
    return etal[o], etar[o], etaa[o], expn[o]


@pytest.mark.parametrize("name,nind,kw", [("prony_3", 3, dict(scale=20, n=200, npsd=4)),
                                          ("prony_4", 4, dict(scale=30, n=400, npsd=6)),
                                          ("prony_3a", [3, "a"], dict(scale=20, n=200, npsd=4))])
def test_prony_refit_matches_reference(name, nind, kw):
    """Prony refit of a Drude bath's C(t): the exponents and coefficients of the reference's
    ``decompose_spectrum_prony`` (the order of purely damped terms is decided there by
    rounding-size imaginary parts, so the comparison is up to that order), and the fit itself
    against the sampled correlation function."""
    from pyqed_b200.heom.spectrum import decompose_spectrum_prony
    g = golden("bath")
    spe = 2 * 0.5 * 1.0 * w / (1.0 ** 2 + w ** 2)
    got = decompose_spectrum_prony(spe, w, 0.7, list(nind) if isinstance(nind, list) else nind, **kw)
    ref = tuple(g[f"{name}_{k}"] for k in ("etal", "etar", "etaa", "expn"))
    for a, b in zip(_canonical(got), _canonical(ref)):
        assert np.allclose(a, b, rtol=1e-6, atol=1e-8)
    t = np.linspace(0, kw["scale"], 2 * kw["n"] + 1)
    etal_p, _, _, expn_p = decompose_spectrum_pade(spe, w, 0.7, kw["npsd"])
    exact = B.fit_t(t, expn_p, etal_p)
    assert np.max(np.abs(B.fit_t(t, got[3], got[0]) - exact)) < 3e-3 * np.max(np.abs(exact))
    # the solver's conventions: etar = conj(etal) for damped terms, etaa = sqrt(|etal|)
    assert np.allclose(got[1], np.conj(got[0])) and np.allclose(got[2], np.sqrt(np.abs(got[0])))


def test_prony_helpers():
    # a signal with many decay rates (the method takes the (nind+1)-th Hankel singular vector, so it
    # is meant for signals richer than the fit): few exponentials reproduce it
    n, scale = 150, 12.0
    t = np.linspace(0, scale, 2 * n + 1)
    rates = 0.25 * 1.7 ** np.arange(9)
    h = sum(0.5 ** k * np.exp(-r * t) for k, r in enumerate(rates))
    gam = B.prony_find_gamma(h, n, 4)
    assert gam.shape == (4,) and np.all(np.abs(gam) < 1) and np.all(np.abs(gam.imag) < 1e-8)
    etal, etar, etaa, expn = B.prony_fitting(h.astype(complex), np.linspace(0, 1, 2 * n + 1), 4, scale, n)
    assert np.max(np.abs(B.fit_t(t, expn, etal) - h)) < 2e-3 * h[0]
    assert np.all(expn.real > 0) and np.allclose(etar, np.conj(etal)) and np.allclose(etaa, np.sqrt(np.abs(etal)))
    amp, ex, err = B.prony_decomposition(t, h, 4)
    assert err < 1e-6 and np.all(ex.real > 0)
    with pytest.raises(ValueError):
        from pyqed_b200.heom.spectrum import decompose_spectrum_prony
        decompose_spectrum_prony(2 * w / (1 + w ** 2), w, 1.0, ["a", 2], scale=10, n=50, npsd=2)
    assert np.allclose(B.spectrum_exp(np.array([0.5]), [1.0], [2.0]), 2.0 / (1.0 - 0.5j))
    # sort_symmetry: conjugate pairs first (by |Im expn|), partner's conjugate as etar
    el, er, ea, ex = B.sort_symmetry(np.array([1.0, 2 + 1j, 2 - 1j]), np.array([0.5, 1 + 3j, 1 - 3j]), if_sqrt=False)
    assert np.allclose(np.abs(ex.imag), [3, 3, 0]) and np.allclose(er[:2], np.conj(el[:2][::-1])) and ea[2] == 1.0
