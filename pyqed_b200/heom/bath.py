"""Bath correlation-function exponents for the DEOM/HEOM solvers (host side).

Mirrors the reference's ``Bath`` container and spectrum decompositions
(``pyqed/heom/deom.py:84-307, 545-552, 895-942``): a bath is five arrays
``expn, etal, etar, etaa`` (complex128[K]) and ``mode`` (int[K]) such that
``C(t) = sum_k etal_k exp(-expn_k t)``.  This is host-only, O(ms) work done
once per run; the device path only ever sees the five arrays.

Implemented here from the published formulas rather than by symbolic algebra:

* ``bose_poles``      - Matsubara poles (``deom.py:84-102``) and the [N-1/N], [N/N], [N+1/N]
  Pade spectrum decompositions of the Bose function (Hu, Xu, Yan, JCP 133, 101106 (2010);
  Hu et al., JCP 134, 244106 (2011); ``deom.py:104-207`` with ``pade=1, 2, 3``).
* ``sort_symmetry``, ``prony_find_gamma``, ``prony_fitting`` - exponential fits of a sampled
  correlation function (``deom.py:36-58, 428-504``).
* ``rational_exponents`` - exponents of any spectral density given as a ratio
  of polynomials in omega (what ``decompose_spectrum_pade`` does via sympy,
  ``deom.py:226-307``).
* ``drude_exponents`` - the Drude-Lorentz special case in closed form.
* ``single_oscillator`` - ``deom.py:545-552``.
"""
from __future__ import annotations

import numpy as np

C128 = np.complex128


def _tridiag_spectrum(offdiag):
    """Eigenvalues, largest first, of the symmetric tridiagonal matrix with zero diagonal
    and the given off-diagonal (``tseig``, ``deom.py:74-79``)."""
    mat = np.diag(offdiag, 1) + np.diag(offdiag, -1)
    return np.sort(np.linalg.eigvalsh(mat))[::-1]


def _psd_offdiag(size, first):
    b = first + 2.0 * np.arange(size)
    return 1.0 / np.sqrt(b[:-1] * b[1:])


def bose_poles(n: int, pade: int = 1):
    """Poles ``p_j`` and residues ``r_j`` of
    ``1/(1-exp(-x)) ~ 1/x + 1/2 + sum_j 2 r_j x / (x^2 + p_j^2)``
    (``pade_approximation_distribution``, ``deom.py:104-207``, bosonic case).

    ``pade=0``: Matsubara, ``p_j = 2 pi j``, ``r_j = 1``.  ``pade=1, 2, 3``: the [N-1/N],
    [N/N] and [N+1/N] Pade spectrum decompositions of Hu, Luo, Jiang, Xu, Yan, JCP 134, 244106
    (2011): the poles are ``2/lambda`` for the positive eigenvalues ``lambda`` of a symmetric
    tridiagonal matrix; the residues follow from the zeros of the numerator polynomial
    (pade 1, 2) or from its three-term recurrence (pade 3).
    """
    if n < 0 or pade not in (0, 1, 2, 3):
        raise ValueError("N or BoseFermi or pade has wrong value!")
    if n == 0:
        return np.zeros(0), np.zeros(0)
    if pade == 0:
        return 2.0 * np.pi * np.arange(1, n + 1), np.ones(n)
    if pade in (1, 2):
        size = 2 * n + (1 if pade == 2 else 0)
        xi2 = (2.0 / _tridiag_spectrum(_psd_offdiag(size, 3.0))[:n]) ** 2          # poles squared
        nz = (size - 1) // 2                                                       # numerator zeros
        zeta2 = (2.0 / _tridiag_spectrum(_psd_offdiag(size - 1, 5.0))[:nz]) ** 2
        lead = 0.5 * n * (2.0 * n + 3.0) if pade == 1 else 0.125 / ((n + 1.0) * (2.0 * n + 3.0))
        resi = np.empty(n)
        for j in range(n):
            others = np.delete(xi2, j)
            resi[j] = lead * np.prod(zeta2 - xi2[j]) / np.prod(others - xi2[j])
        return np.sqrt(xi2), resi
    # pade == 3: coefficients d_m of the continued fraction, b = 3 for bosons
    b = 3.0
    m = n + 1
    d = np.empty(2 * m)
    d[0] = 0.25 / b
    for i in range(1, m + 1):
        d[2 * i - 1] = -4.0 * i * i * (b + 2.0 * i - 2.0) ** 2 * (b + 4.0 * i - 2.0)
    for i in range(1, m):
        d[2 * i] = -0.25 * (b + 4.0 * i) / (i * (i + 1.0) * (b + 2.0 * i - 2.0) * (b + 2.0 * i))
    odd_sum = np.cumsum(d[1::2])                       # d_1, d_1 + d_3, ...
    size = 2 * n + 1
    off = 1.0 / np.sqrt(d[1:size] * d[2:size + 1])
    pole = 2.0 / _tridiag_spectrum(off)[:n]
    resi = np.empty(n)
    for j in range(n):
        x2 = pole[j] * pole[j]
        prev_r, t = 0.0, 0.25 / d[1]
        e0, e1, e2 = 0.0, 0.5, 0.0
        for i in range(m):
            q = t if (i == j or i == n) else t / (pole[i] * pole[i] - x2)
            r_even = 2.0 * np.sqrt(abs(q))
            r_odd = r_even if q > 0 else -r_even
            e2 = d[2 * i] * r_even * e1 - 0.25 * r_even * prev_r * x2 * e0
            e0, e1 = e1, e2
            e2 = d[2 * i + 1] * r_odd * e1 - 0.25 * r_odd * r_even * x2 * e0
            e0, e1 = e1, e2
            prev_r = r_odd
            if i != n:
                t = odd_sum[i] / odd_sum[i + 1]
        resi[j] = e2
    return pole, resi


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


def rational_residues(numer, denom):
    """Exponents ``expn = i z`` and coefficients ``-i Res-like(J, z)`` for the lower-half-plane
    poles ``z`` of ``J = numer/denom`` WITHOUT the Bose factor: what
    ``decompose_spectrum_pade_real`` / ``_imag`` (``deom.py:310-425``) compute.  Returns
    ``(etal, etar, etaa, expn)``; ``etal/etar/etaa`` are ordered conjugate pairs first (by
    decreasing ``|Im expn|``), then damped poles, while ``expn`` stays in pole order - the
    reference returns its unsorted list (``deom.py:366-367``), which only matters when there is
    more than one pole."""
    numer = np.atleast_1d(np.asarray(numer, dtype=C128))
    denom = np.atleast_1d(np.asarray(denom, dtype=C128))
    lead = denom[0]
    poles = np.roots(denom)
    scale = max(1.0, float(np.max(np.abs(poles)))) if len(poles) else 1.0
    poles = np.array(sorted(poles, key=lambda z: (round(z.real / scale, 10), z.imag)))
    expn_raw = np.array([1j * z for z in poles if z.imag < 0], dtype=C128)
    order = np.argsort(np.abs(expn_raw.imag), kind="stable")[::-1]
    srt = expn_raw[order]
    osc = lambda e: abs(e.imag) > 1e-12 * max(1.0, abs(e))   # (np.roots leaves rounding-size phases)
    paired = [e for e in srt if osc(e)]
    single = [complex(e.real, 0.0) for e in srt if not osc(e)]
    expn_raw = np.array([e if osc(e) else complex(e.real, 0.0) for e in expn_raw], dtype=C128)

    def coeff(e):
        z0 = -1j * e
        rest = np.prod([z0 - p for p in poles if abs(p - z0) > 1e-12 * max(1.0, abs(z0))])
        return -1j * np.polyval(numer, z0) / (lead * rest)

    etal, etar, etaa = [], [], []
    for i in range(0, len(paired) - 1, 2):
        a, b2 = coeff(paired[i]), coeff(paired[i + 1])
        etal += [a, b2]
        etar += [np.conj(b2), np.conj(a)]
        etaa += [np.sqrt(abs(a) * abs(b2)), np.sqrt(abs(b2) * abs(a))]
    for e in single:
        a = coeff(e)
        etal.append(a)
        etar.append(np.conj(a))
        etaa.append(abs(a))
    return np.array(etal, C128), np.array(etar, C128), np.array(etaa), expn_raw


def fit_t(t, expn, etal):
    """``sum_k etal_k exp(-expn_k t)`` on the time grid ``t`` (``deom.py:61-64``)."""
    t = np.asarray(t, dtype=float)
    return (np.asarray(etal, C128)[None, :] * np.exp(-np.outer(t, np.asarray(expn, C128)))).sum(axis=1)


def spectrum_exp(w, expn, etal, sigma=-1):
    """``sum_k etal_k / (expn_k + sigma i w)`` (``deom.py:21-24``)."""
    w = np.asarray(w, dtype=float)
    return (np.asarray(etal, C128)[None, :] / (np.asarray(expn, C128)[None, :] + sigma * 1j * w[:, None])).sum(axis=1)


def sort_symmetry(etal, expn, if_sqrt=True):
    """Order fitted exponentials as the solver expects (``deom.py:36-58``): by decreasing
    ``|Im expn|``, oscillating terms (``|Im| > 1e-10``) taken as consecutive conjugate pairs with
    ``etar`` = conjugate of the partner's ``etal``, damped terms with ``etar = conj(etal)``;
    ``etaa = |etal|`` (its square root with ``if_sqrt``, the reference's default).  Returns
    ``(etal, etar, etaa, expn)``."""
    etal, expn = np.asarray(etal, C128), np.asarray(expn, C128)
    order = np.argsort(np.abs(expn.imag))[::-1]
    etal, expn = etal[order], expn[order]
    ncc = int(np.count_nonzero(np.abs(expn.imag) > 1e-10))
    if ncc % 2 and ncc == len(etal):
        raise ValueError("oscillating exponents must come in conjugate pairs")
    etar = np.conj(etal)
    for i in range(0, ncc, 2):
        etar[i], etar[i + 1] = np.conj(etal[i + 1]), np.conj(etal[i])
    # an odd count (a fitted decay factor with a rounding-size phase) pairs the last oscillating
    # term with the first damped one in the reference's loop, whose second loop then resets the
    # damped term; the same here
    etar[ncc:] = np.conj(etal[ncc:])
    etaa = np.abs(etal)
    return etal, etar, (np.sqrt(etaa) if if_sqrt else etaa), expn


def prony_find_gamma(h, n_sample, nind):
    """The ``nind`` decay factors ``gamma`` of ``h_j ~ sum_i c_i gamma_i^j`` from the Hankel
    matrix of the real samples ``h`` (``deom.py:428-447``): con-eigenvector number ``nind``
    (by decreasing singular value) of ``H[i, j] = h[i + j]`` is the coefficient vector of a
    polynomial whose ``nind`` roots of smallest modulus are the decay factors."""
    h = np.asarray(h, dtype=float)
    hank = np.array([h[i:i + n_sample] for i in range(n_sample)])
    vals, vecs = np.linalg.eigh(hank)               # real symmetric: Takagi vectors = eigenvectors x phase
    order = np.argsort(np.abs(vals))[::-1]
    vec = vecs[:, order[nind]] * np.exp(-0.5j * np.angle(vals[order[nind]] + 0j))
    roots = np.roots(vec[::-1])
    return roots[np.argsort(np.abs(roots))[:nind]]


def prony_fitting(h, t, nind, scale, n, gamma_real=None, gamma_imag=None):
    """Fit ``h(t) ~ sum_i etal_i exp(-expn_i t)`` on ``2n+1`` equidistant samples over a window of
    length ``scale`` (``deom.py:450-504``).  ``nind``: number of exponentials found from the real
    part, or ``[n_real, n_imag]`` to take decay factors from the real and from the imaginary part
    separately (either set may be supplied).  The amplitudes are the complex least-squares
    solution (the reference forms the normal equations of the same problem).  Returns
    ``sort_symmetry(etal, expn)``."""
    h = np.asarray(h, C128)
    if isinstance(nind, list):
        g_r = prony_find_gamma(h.real, n, nind[0]) if gamma_real is None else np.asarray(gamma_real)
        g_i = prony_find_gamma(h.imag, n, nind[1]) if gamma_imag is None else np.asarray(gamma_imag)
        gamma = np.append(g_r, g_i)
    else:
        gamma = prony_find_gamma(h.real, n, nind)
    gamma = np.asarray(gamma, C128)
    powers = gamma[None, :] ** np.arange(len(t))[:, None]
    amp, *_ = np.linalg.lstsq(powers, h, rcond=None)
    expn = -2.0 * n * np.log(gamma) / scale
    return sort_symmetry(amp, expn)


def prony_decomposition(x, ct, nexp):
    """Real-part Prony fit of samples ``ct`` on the equidistant grid ``x`` (``pyqed/heom/prony.py:38-141``,
    which reads the window length from a module global and plots; here the window is ``x[-1] - x[0]``
    and nothing is drawn).  Returns ``(etal, expn, err)`` with ``err`` the mean squared residual."""
    x, ct = np.asarray(x, dtype=float), np.asarray(ct)
    n = (len(x) - 1) // 2
    gamma = prony_find_gamma(ct.real, n + 1, nexp)
    powers = gamma[None, :] ** np.arange(2 * n + 1)[:, None]
    amp, *_ = np.linalg.lstsq(powers, ct.real[:2 * n + 1].astype(C128), rcond=None)
    expn = -2.0 * n * np.log(gamma) / (x[-1] - x[0])
    err = float((np.abs(ct - fit_t(x - x[0], expn, amp)) ** 2).sum() / len(x))
    return amp, expn, err


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
