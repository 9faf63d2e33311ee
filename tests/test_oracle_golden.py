"""Pins the CPU oracle to outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest

from conftest import golden, deom_golden_names, pulse_from_samples
from oracle import deom_oracle as DO
from oracle import chain_oracle as CO

TOL = 1e-12  # oracle vs reference: same arithmetic up to summation order


def _oracle_from(g):
    dt = float(g["dt"])
    return DO.DeomOracle(g["system"], g["system_dipole"], g["coupling"], g["coupling_dipole"],
                         g["expn"], g["etal"], g["etar"], g["etaa"], g["mode"], int(g["lmax"]),
                         pulse_from_samples(g["pulse_system"], dt),
                         pulse_from_samples(g["pulse_coupling"], dt))


@pytest.mark.parametrize("name", deom_golden_names())
def test_deom_batched_matches_reference(name):
    g = golden(name)
    o = _oracle_from(g)
    assert np.array_equal(o.keys, g["keys"].astype(np.int64))
    p1 = g["p1"] if "p1" in g else None
    ts, traj = o.run(g["rho0"], float(g["dt"]), int(g["nt"]), p1=p1, batched=True)
    traj = np.asarray(traj)
    assert np.allclose(ts, g["t_save"], rtol=0, atol=1e-15)
    assert np.max(np.abs(traj - g["traj"])) < TOL
    if "ados_final" in g:
        assert np.max(np.abs(o.ddos - g["ados_final"])) < TOL


@pytest.mark.parametrize("name", ["deom_random4_herm", "deom_random5_nonherm", "deom_random3_K1",
                                  "deom_polariton8_L4", "deom_fmo_K21_L2"])
def test_deom_loop_matches_reference(name):
    g = golden(name)
    o = _oracle_from(g)
    nt = min(int(g["nt"]), 4)
    ts, traj = o.run(g["rho0"], float(g["dt"]), nt, p1=None, batched=False)
    assert np.max(np.abs(np.asarray(traj) - g["traj"][:nt + 1])) < TOL


def test_hash_is_tier_major_bijection():
    for K, L in [(1, 5), (2, 10), (3, 4), (7, 4), (21, 2)]:
        tab = DO.pascal_table(K, L)
        keys = DO.build_keys(K, L, tab)
        ids = [DO.ado_id(k, tab) for k in keys]
        assert ids == list(range(len(keys)))
        tiers = keys.sum(axis=1)
        assert np.all(np.diff(tiers) >= 0)
        minus, plus = DO.build_neighbours(keys, L, tab)
        for n in range(len(keys)):
            for k in range(K):
                if minus[n, k] >= 0:
                    assert plus[minus[n, k], k] == n


@pytest.mark.parametrize("tag,fn", [("rk4_nado5", CO.heom_chain_rk4), ("rk4_nado12", CO.heom_chain_rk4),
                                    ("euler_nado5", CO.heom_chain_euler)])
def test_chain_matches_reference(tag, fn):
    g = golden("chain_" + tag)
    obs = fn(g["H"], g["rho0"], [g["c_op"]], list(g["e_ops"]), float(g["temperature"]),
             float(g["cutoff"]), float(g["reorganization"]), int(g["nado"]), float(g["dt"]),
             int(g["nt"]))
    assert np.max(np.abs(obs - g["observables"])) < TOL


def test_known_answers():
    """KAT values recorded independently in SURVEY.md section 8c."""
    g = golden("chain_rk4_nado5")
    assert abs(g["observables"][0, -1].real - (-0.9688066943939262)) < 1e-15
    g = golden("chain_rk4_nado12")
    assert abs(g["observables"][0, -1].real - (-0.9165103413782143)) < 1e-15
    g = golden("chain_euler_nado5")
    assert abs(g["observables"][0, -1].real - (-0.9690784389436947)) < 1e-15
    g = golden("deom_example_L10_p1")
    assert abs(g["traj"][20] - 0.8697707701043433) < 1e-14
    g = golden("deom_fmo_K7_L4")
    pops = np.real(np.diagonal(g["traj"][60]))
    assert np.allclose(pops[:3], [0.50655946, 0.45133341, 0.0160922], atol=1e-8)


def test_chain_equals_deom_form():
    """KAT-3: the chain is the DEOM form with K=1."""
    g = golden("chain_rk4_nado5")
    nado, dt, nt = int(g["nado"]), float(g["dt"]), int(g["nt"])
    D0 = CO.d0_high_temperature(float(g["temperature"]), float(g["cutoff"]), float(g["reorganization"]))
    expn, etal, etar, etaa, mode, lmax = CO.chain_as_deom(D0, float(g["cutoff"]), nado)
    n = g["H"].shape[0]
    o = DO.DeomOracle(g["H"], np.zeros((n, n)), [g["c_op"]], None, expn, etal, etar, etaa, mode, lmax)
    _, traj = o.run(g["rho0"], dt, nt)
    sz = g["e_ops"][0]
    got = np.array([np.trace(sz @ r) for r in traj[1:]])
    assert np.max(np.abs(got - g["observables"][0])) < 1e-13


def test_difference_form_rk4_is_classical_rk4():
    """The async CUDA kernel keeps the three stage inputs and combines them in
    the last stage, y' = -y/3 + S1/3 + 2 S2/3 + S3/3 + dt/6 k4;
    algebraically this is the reference's RK4 (deom.py:725-766)."""
    g = golden("deom_random5_nonherm")
    o = _oracle_from(g)
    rng = np.random.default_rng(11)
    y = rng.standard_normal((o.nmax, o.nsys, o.nsys)) + 1j * rng.standard_normal((o.nmax, o.nsys, o.nsys))
    dt, t = float(g["dt"]), 2 * float(g["dt"])   # on the fixture's pulse grid
    ref = o.rk4_step(y, dt, t, o.rhs_batched)
    s1 = y + dt / 2 * o.rhs_batched(y, t)
    s2 = y + dt / 2 * o.rhs_batched(s1, t + dt / 2)
    s3 = y + dt * o.rhs_batched(s2, t + dt / 2)
    k4 = o.rhs_batched(s3, t + dt)
    new = -y / 3 + s1 / 3 + 2 * s2 / 3 + s3 / 3 + dt / 6 * k4
    assert np.max(np.abs(new - ref)) < 1e-13


# --------------------------------------------------------------------------
# C / OpenMP restatement (oracle/heom_oracle.c)
# --------------------------------------------------------------------------
def _c_run(g, threads=0, nt=None):
    from oracle import c_oracle
    dt = float(g["dt"])
    nt = int(g["nt"]) if nt is None else nt
    traj, ados = c_oracle.run(g["system"], g["system_dipole"], g["coupling"], g["coupling_dipole"],
                              g["expn"], g["etal"], g["etar"], g["etaa"], g["mode"], int(g["lmax"]),
                              g["rho0"], dt, nt, pulse_from_samples(g["pulse_system"], dt),
                              pulse_from_samples(g["pulse_coupling"], dt), threads=threads)
    if "p1" in g:
        traj = np.array([np.trace(g["p1"] @ r) for r in traj])
    return traj, ados


@pytest.mark.parametrize("name", deom_golden_names())
def test_c_oracle_matches_reference(name):
    g = golden(name)
    traj, ados = _c_run(g)
    assert np.max(np.abs(traj - g["traj"])) < TOL
    if "ados_final" in g:
        assert np.max(np.abs(ados - g["ados_final"])) < TOL


def test_c_oracle_keys_and_thread_independence():
    from oracle import c_oracle
    for K, L in [(1, 5), (2, 10), (3, 4), (7, 4), (21, 2)]:
        assert np.array_equal(c_oracle.keys(K, L), DO.build_keys(K, L))
    g = golden("deom_fmo_K7_L4")
    a, _ = _c_run(g, threads=1, nt=5)
    b, _ = _c_run(g, threads=4, nt=5)
    assert np.array_equal(a, b)   # each ADO is reduced by one thread in a fixed order
