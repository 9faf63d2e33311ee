"""Synthetic inputs for the BASELINE.json configurations.

Pure NumPy, no device code: these builders only produce the arrays a user would
hand to ``DEOMSolver`` / ``HEOMSolver``.  They are shared by ``bench.py``, the
parity tests and ``tests/golden/make_golden.py`` so that every leg (CUDA path,
oracle, unmodified reference) sees identical inputs.

Each builder returns a dict with keys
``system, system_dipole, coupling, coupling_dipole, expn, etal, etar, etaa,
mode, lmax, rho0, dt, nt`` (+ ``pulse_system_func`` / ``pulse_coupling_func``
where the configuration has a field).
"""
from __future__ import annotations

import numpy as np

from .heom.bath import drude_exponents

C128 = np.complex128
CM2RADPS = 0.188365  # rad ps^-1 per cm^-1

_FMO_CM = np.array([
    [200.0, -87.7, 5.5, -5.9, 6.7, -13.7, -9.9],
    [-87.7, 320.0, 30.8, 8.2, 0.7, 11.8, 4.3],
    [5.5, 30.8, 0.0, -53.5, -2.2, -9.6, 6.0],
    [-5.9, 8.2, -53.5, 110.0, -70.7, -17.0, -63.3],
    [6.7, 0.7, -2.2, -70.7, 270.0, 81.1, -1.3],
    [-13.7, 11.8, -9.6, -17.0, 81.1, 420.0, 39.7],
    [-9.9, 4.3, 6.0, -63.3, -1.3, 39.7, 230.0],
])


def _pauli():
    s0 = np.eye(2, dtype=C128)
    sx = np.array([[0, 1], [1, 0]], dtype=C128)
    sy = np.array([[0, -1j], [1j, 0]], dtype=C128)
    sz = np.array([[1, 0], [0, -1]], dtype=C128)
    return s0, sx, sy, sz


def _pack(H, Q, expn, etal, etar, etaa, mode, lmax, rho0, dt, nt, mu=None, qd=None,
          f=None, g=None, name=""):
    H = np.asarray(H, dtype=C128)
    n = H.shape[0]
    Q = np.asarray(Q, dtype=C128)
    return dict(
        name=name,
        system=H,
        system_dipole=np.zeros((n, n), C128) if mu is None else np.asarray(mu, C128),
        coupling=Q,
        coupling_dipole=np.zeros_like(Q) if qd is None else np.asarray(qd, C128),
        expn=np.asarray(expn, C128), etal=np.asarray(etal, C128),
        etar=np.asarray(etar, C128), etaa=np.asarray(etaa, C128),
        mode=np.asarray(mode, np.int64), lmax=int(lmax),
        rho0=np.asarray(rho0, C128), dt=float(dt), nt=int(nt),
        pulse_system_func=f, pulse_coupling_func=g,
    )


# --------------------------------------------------------------------------
def spin_boson(lmax=10, npsd=1, pade=0, lam=0.2, gam=1.0, beta=1.0, dt=0.01, nt=1000):
    """Config 1 (DEOM form): H = -sx/2 - sz/2, Q = sz, Drude bath + ``npsd``
    Matsubara (``pade=0``) or Pade terms; K = 1 + npsd."""
    _, sx, _, sz = _pauli()
    expn, etal, etar, etaa = drude_exponents(lam, gam, beta, npsd, pade)
    rho0 = np.zeros((2, 2), C128)
    rho0[1, 1] = 1
    return _pack(-0.5 * sx - 0.5 * sz, [sz], expn, etal, etar, etaa,
                 np.zeros(len(expn), np.int64), lmax, rho0, dt, nt, name="spin_boson")


def spin_boson_deom_example(lmax=10, dt=0.01, nt=200):
    """``examples/deom.py:23-74`` of the reference: H = sz + sx, Q = sx,
    beta = lambda = gamma = 1, Pade npsd = 2 (K = 3)."""
    _, sx, _, sz = _pauli()
    expn, etal, etar, etaa = drude_exponents(1.0, 1.0, 1.0, 2, 1)
    rho0 = np.zeros((2, 2), C128)
    rho0[0, 0] = 1
    return _pack(sz + sx, [sx], expn, etal, etar, etaa, [0, 0, 0], lmax, rho0, dt, nt,
                 name="spin_boson_deom_example")


def fmo_hamiltonian():
    """7-site FMO exciton Hamiltonian in rad/ps."""
    return (_FMO_CM * CM2RADPS).astype(C128)


def fmo(lmax=4, n_matsubara=0, dt=None, nt=1000, lam_cm=35.0, gam_cm=106.18,
        kT_cm=208.51, sites=7):
    """Configs 2 and 3: 7-site FMO, one Drude bath per site, Q_m = |m><m|.

    ``n_matsubara = 0`` -> config 2: the high-temperature single exponent per
    bath, eta = lambda (2kT - i gamma)  (K = 7).
    ``n_matsubara = 2`` -> config 3: Drude pole + 2 Matsubara terms per bath
    (K = 21, ``mode = repeat(arange(7), 3)``); dt defaults to 0.25 fs for RK4
    stability.
    """
    H = fmo_hamiltonian()[:sites, :sites]
    Q = np.zeros((sites, sites, sites), C128)
    for m in range(sites):
        Q[m, m, m] = 1
    lam, gam, kT = lam_cm * CM2RADPS, gam_cm * CM2RADPS, kT_cm * CM2RADPS
    if n_matsubara == 0:
        eta = lam * (2 * kT - 1j * gam)
        e1, l1, r1, a1 = [gam], [eta], [np.conj(eta)], [abs(eta)]
        if dt is None:
            dt = 1e-3
    else:
        e1, l1, r1, a1 = drude_exponents(lam, gam, 1.0 / kT, n_matsubara, 0)
        if dt is None:
            dt = 0.25e-3
    per = len(e1)
    expn = np.tile(e1, sites)
    etal = np.tile(l1, sites)
    etar = np.tile(r1, sites)
    etaa = np.tile(a1, sites)
    mode = np.repeat(np.arange(sites), per)
    rho0 = np.zeros((sites, sites), C128)
    rho0[0, 0] = 1
    return _pack(H, Q, expn, etal, etar, etaa, mode, lmax, rho0, dt, nt,
                 name=f"fmo{sites}_K{len(expn)}_L{lmax}")


def polariton(lmax=6, nfock=16, wc=1.0, w0=1.0, g=0.1, lam=0.05, gam=1.0, beta=1.0,
              dt=0.005, nt=500, dense_h=False):
    """Config 4: two-level molecule x ``nfock`` photon states (N = 2 nfock),
    non-RWA dipole-gauge coupling i g mu (a - a^dag) + g^2/wc mu^2 as in the
    reference's ``Polariton.getH`` (``pyqed/polariton/cavity.py:608-678``);
    Q_1 = sz x I, Q_2 = I x (a + a^dag); each bath Drude + 1 Pade term (K = 4)."""
    s0, sx, _, sz = _pauli()
    a = np.diag(np.sqrt(np.arange(1, nfock)), 1).astype(C128)
    ad = a.conj().T
    ic = np.eye(nfock, dtype=C128)
    hmol = 0.5 * w0 * sz
    hcav = wc * (ad @ a)
    H = (np.kron(hmol, ic) + np.kron(s0, hcav) + 1j * g * np.kron(sx, a - ad)
         + g * g / wc * np.kron(sx @ sx, ic))
    if dense_h:  # SURVEY 8d's stress variant: a dense random Hermitian perturbation (seed 0)
        rng = np.random.default_rng(0)
        r = rng.normal(size=H.shape) + 1j * rng.normal(size=H.shape)
        H = H + 0.05 * (r + r.conj().T) / 2
    Q = np.stack([np.kron(sz, ic), np.kron(s0, a + ad)])
    e1, l1, r1, a1 = drude_exponents(lam, gam, beta, 1, 1)
    expn, etal, etar, etaa = (np.tile(x, 2) for x in (e1, l1, r1, a1))
    mode = np.repeat(np.arange(2), len(e1))
    n = 2 * nfock
    rho0 = np.zeros((n, n), C128)
    rho0[0, 0] = 1  # |e, 0> in the (sz = +1) x Fock ordering
    return _pack(H, Q, expn, etal, etar, etaa, mode, lmax, rho0, dt, nt,
                 name=f"polariton{n}_K{len(expn)}_L{lmax}")


def aggregate_2des(lmax=6, waiting_index=0, dt=0.01, nt=700, lam=0.1, gam=1.0, beta=1.0):
    """Config 5: 3-site Frenkel aggregate with ground, 3 single and 3 double
    excitons (N = 7), site-occupation couplings (diagonal, M = 3), each bath
    Drude + 1 Pade (K = 6), driven through ``system_dipole`` by a pump/probe pair
    separated by the waiting time ``0.05 * waiting_index``."""
    eps = np.array([10.0, 10.5, 11.0])
    J = np.array([[0, 0.3, 0.1], [0.3, 0, 0.3], [0.1, 0.3, 0]])
    pairs = [(0, 1), (0, 2), (1, 2)]
    n = 7
    H = np.zeros((n, n), C128)
    for m in range(3):
        H[1 + m, 1 + m] = eps[m]
        for k in range(3):
            if k != m:
                H[1 + m, 1 + k] = J[m, k]
    for p, (m, k) in enumerate(pairs):
        H[4 + p, 4 + p] = eps[m] + eps[k]
    for p, (m, k) in enumerate(pairs):
        for q, (r, s) in enumerate(pairs):
            if p == q:
                continue
            common = set((m, k)) & set((r, s))
            if len(common) == 1:
                (x,) = set((m, k)) - common
                (y,) = set((r, s)) - common
                H[4 + p, 4 + q] = J[x, y]
    mu = np.zeros((n, n), C128)
    for m in range(3):
        mu[1 + m, 0] = mu[0, 1 + m] = 1
    for p, (m, k) in enumerate(pairs):
        for site in (m, k):
            other = k if site == m else m
            mu[4 + p, 1 + other] = mu[1 + other, 4 + p] = 1
    Q = np.zeros((3, n, n), C128)
    for m in range(3):
        Q[m, 1 + m, 1 + m] = 1
        for p, pr in enumerate(pairs):
            if m in pr:
                Q[m, 4 + p, 4 + p] = 1
    e1, l1, r1, a1 = drude_exponents(lam, gam, beta, 1, 1)
    expn, etal, etar, etaa = (np.tile(x, 3) for x in (e1, l1, r1, a1))
    mode = np.repeat(np.arange(3), len(e1))
    rho0 = np.zeros((n, n), C128)
    rho0[0, 0] = 1
    amp, sig, w, tp = 0.05, 0.5, 10.5, 1.5
    tb = 0.05 * waiting_index

    def field(t, tb=tb):
        return amp * (np.exp(-(t - tp) ** 2 / (2 * sig ** 2))
                      + np.exp(-(t - tp - tb) ** 2 / (2 * sig ** 2))) * np.cos(w * t)

    out = _pack(H, Q, expn, etal, etar, etaa, mode, lmax, rho0, dt, nt, mu=mu,
                f=field, g=None, name=f"aggregate7_K{len(expn)}_L{lmax}_T{waiting_index}")
    out["observable"] = mu
    return out


def random_dense(n=4, nmod=2, nind=3, lmax=3, seed=0, dt=0.01, nt=20, hermitian=True):
    """Stress input: random (optionally non-Hermitian) H, dense Q, complex
    exponents; seeded ``np.random.default_rng(seed)``."""
    rng = np.random.default_rng(seed)

    def rnd(*shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    H = rnd(n, n)
    Q = rnd(nmod, n, n)
    mu = rnd(n, n)
    qd = rnd(nmod, n, n)
    if hermitian:
        H = (H + H.conj().T) / 2
        Q = (Q + Q.conj().transpose(0, 2, 1)) / 2
        mu = (mu + mu.conj().T) / 2
        qd = (qd + qd.conj().transpose(0, 2, 1)) / 2
    expn = rng.uniform(0.5, 2.0, nind) + 1j * rng.uniform(-1, 1, nind)
    etal = rnd(nind) * 0.3
    etar = rnd(nind) * 0.3
    etaa = rng.uniform(0.1, 0.5, nind)
    mode = rng.integers(0, nmod, nind)
    mode[0] = nmod - 1  # make sure max(mode)+1 == nmod
    psi = rnd(n)
    rho0 = np.outer(psi, psi.conj())
    rho0 /= np.trace(rho0)
    return _pack(H, Q, expn, etal, etar, etaa, mode, lmax, rho0, dt, nt, mu=mu, qd=qd,
                 f=lambda t: 0.3 * np.sin(2.0 * t), g=lambda t: 0.1 * np.cos(1.5 * t),
                 name=f"random{n}_M{nmod}_K{nind}_L{lmax}_s{seed}")
