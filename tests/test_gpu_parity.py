"""CUDA path vs the reference's own outputs (tests/golden) and vs the oracle.

Everything here goes through the public drop-in classes, which call the C ABI.
Tolerance: the north star asks for max |rho_sys - rho_ref| <= 1e-10 over the
trajectory in FP64; the kernels differ from NumPy only in summation order, so
the tests hold them to 1e-12.
"""
import numpy as np
import pytest

from conftest import golden, deom_golden_names, pulse_from_samples

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _solver_from(g, **kw):
    from pyqed_b200.heom import DEOMSolver, Bath
    dt = float(g["dt"])
    bath = Bath(expn=g["expn"], etal=g["etal"], etar=g["etar"], etaa=g["etaa"], mode=g["mode"])
    return DEOMSolver(system=g["system"], system_dipole=g["system_dipole"], bath=bath,
                      coupling=g["coupling"], coupling_dipole=g["coupling_dipole"],
                      pulse_system_func=pulse_from_samples(g["pulse_system"], dt),
                      pulse_coupling_func=pulse_from_samples(g["pulse_coupling"], dt),
                      lmax=int(g["lmax"]), **kw)


def _check_against_golden(g, s):
    p1 = g["p1"] if "p1" in g else None
    ts, traj = s.run(g["rho0"].copy(), float(g["dt"]), int(g["nt"]), p1=p1)
    traj = np.asarray(traj)
    assert np.allclose(ts, g["t_save"], rtol=0, atol=1e-15)
    err = np.max(np.abs(traj - g["traj"]))
    assert err < TOL, err
    assert np.array_equal(s.keys, g["keys"].astype(np.int64))
    if "ados_final" in g:
        err = np.max(np.abs(s.ddos - g["ados_final"]))
        assert err < TOL, err


@pytest.mark.parametrize("name", deom_golden_names())
def test_deom_matches_reference(name):
    g = golden(name)
    _check_against_golden(g, _solver_from(g))


@pytest.mark.parametrize("name", ["deom_fmo_K7_L4", "deom_random4_herm", "deom_random5_nonherm",
                                  "deom_spin_boson_L10", "deom_fmo_K21_L2"])
@pytest.mark.parametrize("order", [1, 2])
def test_lexicographic_storage_order(name, order):
    g = golden(name)
    _check_against_golden(g, _solver_from(g, order=order))


@pytest.mark.parametrize("name", ["deom_fmo_K7_L4", "deom_random4_herm", "deom_random5_nonherm",
                                  "deom_aggregate_L3_T37", "deom_random3_K1"])
def test_generic_kernel_small_n(name):
    g = golden(name)
    s = _solver_from(g)
    s.tuning = dict(kernel=2, warps_per_cta=0, use_graph=0)
    _check_against_golden(g, s)


@pytest.mark.parametrize("opts", [{"qdiag": 0}, {"hermitian": 0}, {"qdiag": 0, "hermitian": 0},
                                  {"real_h": 0}, {"rk13": 0, "resident": 0}, {"rk13": 1, "resident": 0}])
@pytest.mark.parametrize("name", ["deom_fmo_K7_L4", "deom_fmo_K21_L2", "deom_spin_boson_L10",
                                  "deom_aggregate_L3_T37"])
def test_structure_fast_paths_can_be_disabled(name, opts):
    """The diagonal-Q and Hermitian fast paths are optimisations only: with
    them off the general code must give the same trajectory."""
    g = golden(name)
    s = _solver_from(g)
    s.options = dict(opts)
    _check_against_golden(g, s)
    assert s._plan.info("qdiag") == (0 if opts.get("qdiag") == 0 else 1)


def test_fast_path_detection():
    for name, qdiag, herm in [("deom_fmo_K7_L4", 1, 1), ("deom_random4_herm", 0, 0),
                              ("deom_random5_nonherm", 0, 0), ("deom_polariton8_L4", 0, 1),
                              ("deom_spin_boson_L10", 1, 1)]:
        g = golden(name)
        s = _solver_from(g)
        s.run(g["rho0"].copy(), float(g["dt"]), 1)
        assert s._plan.info("q_diagonal") == qdiag, name
        assert s._plan.info("hermitian") == herm, name


@pytest.mark.parametrize("kernel", [1, 3])
@pytest.mark.parametrize("warps", [1, 2, 4, 8])
def test_warps_per_cta(warps, kernel):
    g = golden("deom_fmo_K21_L2")
    s = _solver_from(g)
    s.tuning = dict(kernel=kernel, warps_per_cta=warps, use_graph=0)
    _check_against_golden(g, s)


@pytest.mark.parametrize("kernel", [1, 3])
@pytest.mark.parametrize("name", ["deom_fmo_K7_L4", "deom_fmo_K21_L3", "deom_spin_boson_L10",
                                  "deom_aggregate_L3_T0"])
@pytest.mark.parametrize("herm,sym", [(-1, -1), (-1, 0), (0, -1)])
def test_row_kernels_on_diagonal_coupling(name, kernel, herm, sym):
    """Both row kernels (plain loads / cp.async staging), with and without
    the Hermitian row fetch and the Hermitian-symmetric shortcuts of the async
    kernel, on projector, sigma_z and occupation couplings."""
    g = golden(name)
    s = _solver_from(g)
    s.tuning = dict(kernel=kernel, warps_per_cta=0, use_graph=0)
    s.options = {"hermitian": herm, "sym": sym}
    _check_against_golden(g, s)
    assert s._plan.info("resident_launches") == 0
    if "fmo" in name:   # projector couplings: one support row per mode
        assert s._plan.info("sym") == (1 if herm != 0 and sym != 0 else 0)
    if "spin_boson" in name:   # sigma_z has two non-zero diagonal entries
        assert s._plan.info("sym") == 0


@pytest.mark.parametrize("name", ["deom_fmo_K7_L4", "deom_fmo_K21_L2", "deom_spin_boson_L10",
                                  "deom_aggregate_L3_T0", "deom_aggregate_L3_T37"])
@pytest.mark.parametrize("resident", [0, 1, 4])
def test_cluster_resident_kernel(name, resident):
    """Small hierarchies are propagated by one cluster-resident launch (1: the
    element-parallel kernel 5, 4: the row-per-lane kernel 4); with the option off
    the per-stage kernels must give the same trajectory."""
    g = golden(name)
    s = _solver_from(g)
    s.options = {"resident": resident}
    if resident:
        s.tuning = dict(kernel=4, warps_per_cta=0, use_graph=0)
    _check_against_golden(g, s)
    assert (s._plan.info("resident_launches") > 0) == bool(resident)


def test_cluster_resident_batch_and_restart():
    """Batch of trajectories (one cluster each) and two consecutive run() calls."""
    ga, gb = golden("deom_aggregate_L3_T0"), golden("deom_aggregate_L3_T37")
    s = _solver_from(ga)
    dt, nt = float(ga["dt"]), int(ga["nt"])
    fa = pulse_from_samples(ga["pulse_system"], dt)
    fb = pulse_from_samples(gb["pulse_system"], dt)
    for _ in range(2):
        ts, out = s.run_batch([ga["rho0"]] * 5, dt, nt, p1=ga["p1"],
                              pulse_system_funcs=[fa, fb, None, fb, fa])
        assert np.max(np.abs(out[0] - ga["traj"])) < TOL and np.max(np.abs(out[4] - ga["traj"])) < TOL
        assert np.max(np.abs(out[1] - gb["traj"])) < TOL and np.max(np.abs(out[3] - gb["traj"])) < TOL
        assert np.max(np.abs(out[2])) < TOL
    assert s._plan.info("resident_launches") == 2


@pytest.mark.parametrize("tag", ["rk4_nado5", "rk4_nado12"])
def test_chain_solver_matches_reference(tag):
    from pyqed_b200.heom import HEOMSolver
    g = golden("chain_" + tag)
    sol = HEOMSolver(g["H"], c_ops=[g["c_op"]], e_ops=list(g["e_ops"]), verbose=False)
    obs = sol.run(rho0=g["rho0"], dt=float(g["dt"]), nt=int(g["nt"]),
                  temperature=float(g["temperature"]), cutoff=float(g["cutoff"]),
                  reorganization=float(g["reorganization"]), nado=int(g["nado"]))
    assert obs.shape == g["observables"].shape
    err = np.max(np.abs(obs - g["observables"]))
    assert err < TOL, err


@pytest.mark.parametrize("kernel", [0, 1, 3])
def test_batch_of_waiting_times(kernel):
    """Config 5 shape: trajectories that differ only in their field table
    (kernel 0: one resident launch; 1, 3: per-stage row kernels - the async one
    is launched once per trajectory with that trajectory's pointers)."""
    ga, gb = golden("deom_aggregate_L3_T0"), golden("deom_aggregate_L3_T37")
    s = _solver_from(ga)
    if kernel:
        s.tuning = dict(kernel=kernel, warps_per_cta=0, use_graph=0)
    dt, nt = float(ga["dt"]), int(ga["nt"])
    fa = pulse_from_samples(ga["pulse_system"], dt)
    fb = pulse_from_samples(gb["pulse_system"], dt)
    ts, out = s.run_batch([ga["rho0"], gb["rho0"], ga["rho0"]], dt, nt, p1=ga["p1"],
                          pulse_system_funcs=[fa, fb, None])
    assert np.max(np.abs(out[0] - ga["traj"])) < TOL
    assert np.max(np.abs(out[1] - gb["traj"])) < TOL
    # no field: the ground state never leaves |g><g|, Tr(mu rho) stays 0
    assert np.max(np.abs(out[2])) < TOL


def test_mid_size_against_oracle():
    """Config 3 shape at depth 4 (12650 ADOs): too slow for the reference,
    seconds for the batched oracle (pinned to the reference at depth 2-3)."""
    from oracle.deom_oracle import DeomOracle
    from pyqed_b200 import workloads as W
    from pyqed_b200.heom import DEOMSolver, Bath
    w = W.fmo(lmax=4, n_matsubara=2)
    nt = 3
    o = DeomOracle(w["system"], w["system_dipole"], w["coupling"], w["coupling_dipole"], w["expn"],
                   w["etal"], w["etar"], w["etaa"], w["mode"], w["lmax"])
    _, ref = o.run(w["rho0"], w["dt"], nt)
    for order in (0, 1, 2):
        s = DEOMSolver(w["system"], w["system_dipole"],
                       Bath(expn=w["expn"], etal=w["etal"], etar=w["etar"], etaa=w["etaa"],
                            mode=w["mode"]), w["coupling"], w["coupling_dipole"], lmax=w["lmax"],
                       order=order)
        _, got = s.run(w["rho0"].copy(), w["dt"], nt)
        assert np.max(np.abs(np.asarray(got) - np.asarray(ref))) < TOL
        assert np.max(np.abs(s.ddos - o.ddos)) < TOL
        assert np.array_equal(s.keys, o.keys)


def test_beyond_l2_against_c_oracle():
    """Config 3 shape at depth 6 (296 010 ADOs, 232 MB per array: the state no
    longer fits L2, the per-stage async kernel runs from HBM) against the
    C/OpenMP oracle, every ADO compared; with and without pulses."""
    from oracle import c_oracle
    from pyqed_b200 import workloads as W
    from pyqed_b200.heom import DEOMSolver, Bath
    w = W.fmo(lmax=6, n_matsubara=2)
    nt = 4
    mu = np.diag(np.linspace(-1.0, 1.0, 7)) * 20.0
    for pulse in (None, lambda t: np.exp(-((t - 0.02) / 0.01) ** 2)):
        dip = w["system_dipole"] if pulse is None else mu
        ref, ref_ados = c_oracle.run(w["system"], dip, w["coupling"], w["coupling_dipole"], w["expn"],
                                     w["etal"], w["etar"], w["etaa"], w["mode"], w["lmax"], w["rho0"],
                                     w["dt"], nt, pulse_system=pulse)
        s = DEOMSolver(w["system"], dip,
                       Bath(expn=w["expn"], etal=w["etal"], etar=w["etar"], etaa=w["etaa"],
                            mode=w["mode"]), w["coupling"], w["coupling_dipole"],
                       pulse_system_func=pulse, lmax=w["lmax"])
        _, got = s.run(w["rho0"].copy(), w["dt"], nt)
        assert s._plan.info("resident_launches") == 0
        assert np.max(np.abs(np.asarray(got) - ref)) < TOL
        assert np.max(np.abs(s.ddos - ref_ados)) < TOL
        del s


def test_invariants_large():
    """Size-independent properties on a hierarchy the oracle cannot reach
    (K=21, L=5: 65780 ADOs): trace 1, Hermitian rho_sys, populations in [0,1]."""
    from pyqed_b200 import workloads as W
    from pyqed_b200.heom import DEOMSolver, Bath
    w = W.fmo(lmax=5, n_matsubara=2)
    s = DEOMSolver(w["system"], w["system_dipole"],
                   Bath(expn=w["expn"], etal=w["etal"], etar=w["etar"], etaa=w["etaa"],
                        mode=w["mode"]), w["coupling"], w["coupling_dipole"], lmax=w["lmax"])
    _, traj = s.run(w["rho0"].copy(), w["dt"], 40)
    traj = np.asarray(traj)
    tr = np.trace(traj, axis1=1, axis2=2)
    assert np.max(np.abs(tr - 1.0)) < 1e-12
    assert np.max(np.abs(traj - traj.conj().transpose(0, 2, 1))) < 1e-12
    pops = np.real(np.diagonal(traj, axis1=1, axis2=2))
    assert pops.min() > -1e-9 and pops.max() < 1 + 1e-9
    assert pops[-1, 0] < 1.0 - 1e-4  # something actually moved


def test_empty_and_tiny_inputs():
    """nt = 0 returns just the initial state; depth 0 is plain von Neumann."""
    from pyqed_b200.heom import DEOMSolver, Bath
    g = golden("deom_random4_herm")
    s = _solver_from(g)
    ts, traj = s.run(g["rho0"].copy(), 0.01, 0)
    assert len(traj) == 1 and np.allclose(traj[0], g["rho0"]) and ts.shape == (1,)
    H = g["system"]
    bath = Bath(expn=g["expn"], etal=g["etal"], etar=g["etar"], etaa=g["etaa"], mode=g["mode"])
    s = DEOMSolver(system=H, bath=bath, coupling=g["coupling"], lmax=0)
    dt, nt = 0.01, 50
    _, traj = s.run(g["rho0"].copy(), dt, nt)
    import scipy.linalg as la
    U = la.expm(-1j * H * dt * nt)
    # RK4 truncation error at dt = 0.01 is ~2e-9 here; the check is against the exact propagator
    assert np.max(np.abs(traj[-1] - U @ g["rho0"] @ U.conj().T)) < 1e-7


@pytest.mark.parametrize("name", ["deom_random4_herm", "deom_fmo_K7_L4", "deom_spin_boson_L10"])
def test_euler_method_against_oracle(name):
    from oracle.deom_oracle import DeomOracle
    from pyqed_b200._cabi import Plan
    g = golden(name)
    o = DeomOracle(g["system"], g["system_dipole"], g["coupling"], g["coupling_dipole"], g["expn"],
                   g["etal"], g["etar"], g["etaa"], g["mode"], int(g["lmax"]))
    rho = np.zeros((o.nmax, o.nsys, o.nsys), dtype=np.complex128)
    rho[0] = g["rho0"]
    dt, nt = 0.003, 25
    for i in range(nt):
        rho = rho + dt * o.rhs_batched(rho, 0.0)
    p = Plan(o.nsys, o.nind, int(g["coupling"].shape[0]), int(g["lmax"]))
    p.set_system(g["system"], None)
    p.set_coupling(g["coupling"], None)
    p.set_bath(g["expn"], g["etal"], g["etar"], g["etaa"], g["mode"])
    p.build()
    p.set_state(g["rho0"][None])
    p.propagate(dt, nt, method=1)
    got = p.get_ados()[0]
    assert np.max(np.abs(got - rho)) < TOL


def test_euler_chain_solver_matches_reference():
    """pyqed/oqs.py HEOMSolver (what examples/heom.py runs): KAT-1e."""
    from pyqed_b200.oqs import HEOMSolver
    g = golden("chain_euler_nado5")
    sol = HEOMSolver(g["H"], c_ops=[g["c_op"]], e_ops=list(g["e_ops"]), verbose=False)
    obs = sol.run(rho0=g["rho0"], dt=float(g["dt"]), nt=int(g["nt"]), temperature=float(g["temperature"]),
                  cutoff=float(g["cutoff"]), reorganization=float(g["reorganization"]), nado=int(g["nado"]))
    assert obs.shape == g["observables"].shape
    assert np.max(np.abs(obs - g["observables"])) < TOL
    assert abs(obs[0, -1].real - (-0.9690784389436947)) < 1e-13


def test_liouville_propagators_match_reference():
    from pyqed_b200.heom import HEOMSolver as RK4Solver
    from pyqed_b200.oqs import HEOMSolver as EulerSolver, liouville_propagator
    g = golden("chain_propagator")
    args = dict(dt=float(g["dt"]), nt=int(g["nt"]), temperature=float(g["temperature"]),
                cutoff=float(g["cutoff"]), reorganization=float(g["reorganization"]), nado=int(g["nado"]))
    u = RK4Solver(g["H"], [g["c_op"]], [g["c_op"]], verbose=False).propagator(**args)
    assert u.shape == g["u_heom"].shape and np.max(np.abs(u - g["u_heom"])) < TOL
    u = EulerSolver(g["H"], [g["c_op"]], [g["c_op"]], verbose=False).propagator(**args)
    assert np.max(np.abs(u - g["u_oqs"])) < TOL
    assert np.max(np.abs(g["u_oqs"] - g["u_heom"])) > 1e-6   # the two variants really differ
    T, gam, lam, nado, dt, nt = g["t3"]
    u3 = liouville_propagator(g["H3"], g["S3"], dt, int(nt), T, gam, lam, int(nado))
    assert np.max(np.abs(u3 - g["u3_heom"])) < TOL


def test_heom_space_correlation_function():
    """<A(t)B(0)> by applying B to every ADO and propagating, against the same
    sequence done with the oracle; also operator_action_ddos on both sides."""
    from oracle.deom_oracle import DeomOracle
    g = golden("deom_random4_herm")
    s = _solver_from(g)
    s.pulse_system_func = s.pulse_coupling_func = None
    o = DeomOracle(g["system"], g["system_dipole"], g["coupling"], g["coupling_dipole"], g["expn"],
                   g["etal"], g["etar"], g["etaa"], g["mode"], int(g["lmax"]))
    rng = np.random.default_rng(3)
    n = o.nsys
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    B = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    dt, n_eq, nt = 0.01, 30, 25
    # oracle: equilibrate, apply B to all ADOs, propagate, trace with A
    _, _ = o.run(g["rho0"], dt, n_eq)
    rho = B @ o.ddos
    ref = [np.trace(A @ rho[0])]
    for _ in range(nt):
        rho = o.rk4_step(rho, dt, 0.0, o.rhs_batched)
        ref.append(np.trace(A @ rho[0]))
    s.prepare(g["rho0"], dt, n_eq)
    t, c = s.correlation_2op_1t(A, B, dt, nt)
    assert np.allclose(t, np.arange(nt + 1) * dt)
    assert np.max(np.abs(c - np.array(ref))) < TOL
    # right action
    s.prepare(g["rho0"], dt, n_eq)
    s.operator_action_ddos(B, side="right")
    assert np.max(np.abs(s.ddos - o.ddos @ B)) < TOL


def _random_problem(n, nmod, nind, lmax, seed, diag, herm):
    """Random inputs: diagonal or dense couplings, Hermitian-preserving bath or a
    general one (complex exponents, unrelated etal/etar)."""
    rng = np.random.default_rng(seed)

    def rnd(*shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    H = rnd(n, n)
    H = (H + H.conj().T) / 2
    if rng.integers(2) and herm:
        H = H.real.astype(np.complex128)          # exercises the real-H kernels
    if diag:
        Q = np.zeros((nmod, n, n), np.complex128)
        for m in range(nmod):
            k = int(rng.integers(1, min(3, n) + 1))
            idx = rng.choice(n, size=k, replace=False)
            Q[m, idx, idx] = rng.uniform(0.5, 1.5, k) * rng.choice([-1, 1], k)
    else:
        Q = rnd(nmod, n, n)
        Q = (Q + Q.conj().transpose(0, 2, 1)) / 2
    mode = rng.integers(0, nmod, nind)
    mode[0] = nmod - 1
    if herm:
        expn = rng.uniform(0.5, 2.0, nind).astype(np.complex128)
        etal = rnd(nind) * 0.3
        etar = np.conj(etal)
        etaa = np.abs(etal).astype(np.complex128)
        psi = rnd(n)
        rho0 = np.outer(psi, psi.conj())
        rho0 = (rho0 + rho0.conj().T) / 2 / np.trace(rho0).real   # exactly Hermitian
    else:
        expn = rng.uniform(0.5, 2.0, nind) + 1j * rng.uniform(-1, 1, nind)
        etal, etar = rnd(nind) * 0.3, rnd(nind) * 0.3
        etaa = rng.uniform(0.1, 0.5, nind).astype(np.complex128)
        rho0 = rnd(n, n)
    rho0 = rho0 / np.trace(rho0)
    return dict(H=H, Q=Q, expn=expn, etal=etal, etar=etar, etaa=etaa, mode=mode, lmax=lmax, rho0=rho0)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("diag,herm", [(True, True), (True, False), (False, True), (False, False)])
def test_every_system_size_against_oracle(n, diag, herm):
    """All template instantiations N = 2..8 of the row kernels (and the resident
    kernels where they apply), diagonal and dense couplings, Hermitian-preserving
    and general baths, in all three storage orders, against the batched oracle."""
    from oracle.deom_oracle import DeomOracle
    from pyqed_b200.heom import DEOMSolver, Bath
    q = _random_problem(n, nmod=2, nind=3, lmax=3, seed=100 * n + 2 * diag + herm, diag=diag, herm=herm)
    o = DeomOracle(q["H"], None, q["Q"], None, q["expn"], q["etal"], q["etar"], q["etaa"], q["mode"], q["lmax"])
    dt, nt = 0.004, 12
    _, ref = o.run(q["rho0"], dt, nt)
    bath = Bath(expn=q["expn"], etal=q["etal"], etar=q["etar"], etaa=q["etaa"], mode=q["mode"])
    for order, opts in [(0, {}), (1, {"resident": 0}), (2, {"resident": 0}), (0, {"resident": 4}),
                        (0, {"resident": 0, "rk13": 0})]:
        s = DEOMSolver(q["H"], None, bath, q["Q"], None, lmax=q["lmax"], order=order)
        s.options = opts
        _, got = s.run(q["rho0"].copy(), dt, nt)
        assert np.max(np.abs(np.asarray(got) - np.asarray(ref))) < TOL, (order, opts)
        assert np.max(np.abs(s.ddos - o.ddos)) < TOL, (order, opts)
        assert s._plan.info("q_diagonal") == int(diag)
        assert s._plan.info("hermitian") == int(herm)


def test_batched_general_coupling_and_large_k():
    """Batch > 1 on the dense-coupling path and a wide hierarchy (K = 12)."""
    from oracle.deom_oracle import DeomOracle
    from pyqed_b200.heom import DEOMSolver, Bath
    q = _random_problem(3, nmod=3, nind=12, lmax=2, seed=77, diag=False, herm=True)
    o = DeomOracle(q["H"], None, q["Q"], None, q["expn"], q["etal"], q["etar"], q["etaa"], q["mode"], q["lmax"])
    bath = Bath(expn=q["expn"], etal=q["etal"], etar=q["etar"], etaa=q["etaa"], mode=q["mode"])
    s = DEOMSolver(q["H"], None, bath, q["Q"], None, lmax=q["lmax"])
    rho_b = np.diag([0.2, 0.3, 0.5]).astype(np.complex128)
    dt, nt = 0.005, 10
    _, out = s.run_batch([q["rho0"], rho_b], dt, nt)
    for b, r0 in enumerate((q["rho0"], rho_b)):
        _, ref = o.run(r0, dt, nt)
        assert np.max(np.abs(out[b] - np.asarray(ref))) < TOL


@pytest.mark.parametrize("tag", ["random3_K1", "random4_herm", "spin_boson_L3"])
def test_dense_generator_matches_reference(tag):
    """gen_generate_propgator: the dense generator equals the reference's
    (deom.py:1116-1125) and exponentiating it reproduces the RK4 trajectory,
    the consistency check of pyqed/heom/propagator.py:67-84."""
    import scipy.linalg as la
    from pyqed_b200.heom import DEOMSolver, Bath
    g = golden("generator")
    get = lambda k: g[f"{tag}_{k}"]
    bath = Bath(expn=get("expn"), etal=get("etal"), etar=get("etar"), etaa=get("etaa"), mode=get("mode"))
    s = DEOMSolver(system=get("system"), bath=bath, coupling=get("coupling"), lmax=int(get("lmax")))
    gen = s.gen_generate_propgator(chunk=37)     # chunk that does not divide the dimension
    ref = get("generator")
    assert gen.shape == ref.shape
    assert np.max(np.abs(gen - ref)) < 1e-12
    # exp(G t) on vec(rho0, 0, 0, ...) against the RK4 trajectory
    dt, nt = 0.002, 100
    _, traj = s.run(get("rho0").copy(), dt, nt)
    n = s.nsys
    v0 = np.zeros(gen.shape[0], dtype=np.complex128)
    v0[:n * n] = get("rho0").ravel()
    vt = la.expm(gen * dt * nt) @ v0
    assert np.max(np.abs(vt[:n * n].reshape(n, n) - traj[-1])) < 1e-9


def test_examples_deom_script_flow():
    """The call sequence of the reference's examples/deom.py:23-74 (Mol.deom
    factory, setters, run with p1) against KAT-2."""
    import sympy as sp
    from pyqed_b200.mol import Mol
    from pyqed_b200.heom import Bath
    from pyqed_b200.heom.spectrum import decompose_spectrum_pade as pade
    g = golden("deom_example_L10_p1")
    sz = np.array([[1, 0], [0, -1]], dtype=complex)
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    H = sz + sx
    mol = Mol(H, np.zeros_like(H))
    sdip = np.zeros((2, 2), np.complex128)
    rho = np.zeros((2, 2), dtype=np.complex128)
    rho[0, 0] = 1
    w_sp, lamd_sp, gams_sp = sp.symbols(r"\omega , \lambda, \gamma", real=True)
    spe_sp = (2 * lamd_sp * gams_sp * w_sp / (gams_sp ** 2 + w_sp ** 2)).subs({lamd_sp: 1, gams_sp: 1})
    bath = Bath([spe_sp], w_sp, [1.0], [2], [0, 0, 0], [pade])
    assert np.allclose(bath.expn, g["expn"], rtol=1e-12) and np.allclose(bath.etal, g["etal"], rtol=1e-11)
    solver = mol.deom(bath, [sx])
    solver.set_hierarchy(10)
    solver.set_pulse_system_func(lambda t: 0)
    solver.set_coupling_dipole(sdip)
    solver.set_pulse_coupling_func(lambda t: 0)
    t_save, ddos_save = solver.run(rho0=rho, dt=0.01, nt=20, p1=[[1, 0], [0, 0]])
    assert np.max(np.abs(ddos_save - g["traj"])) < 1e-11
    assert abs(ddos_save[20] - 0.8697707701043433) < 1e-11
    assert solver.nmax == 286 and np.array_equal(solver.keys[:4], [[0, 0, 0], [0, 0, 1], [0, 1, 0], [1, 0, 0]])
    assert abs(rho[0, 0] - ddos_save[20]) < 1e-12   # rho0 is aliased and advanced in place, as in the reference


@pytest.mark.parametrize("tag", ["random3_K1", "spin_boson_L3"])
@pytest.mark.parametrize("lcr", ["llll", "lrcl"])
def test_frequency_domain_four_operator_response(tag, lcr):
    """correlation_4op_3t (deom.py:1127-1210) on the GPU-built generator."""
    from pyqed_b200.heom import DEOMSolver, Bath
    g = golden("generator")
    get = lambda k: g[f"{tag}_{k}"]
    bath = Bath(expn=get("expn"), etal=get("etal"), etar=get("etar"), etaa=get("etaa"), mode=get("mode"))
    s = DEOMSolver(system=get("system"), bath=bath, coupling=get("coupling"), lmax=int(get("lmax")))
    ops = get("c4_ops")
    c = s.correlation_4op_3t(ops[0], ops[1], ops[2], ops[3], get("rho0"), 0.7, get("c4_wx"), get("c4_wy"), lcr=lcr)
    ref = get(f"c4_{lcr}")
    assert c.shape == ref.shape
    assert np.max(np.abs(c - ref)) < 1e-8 * max(1.0, np.max(np.abs(ref)))


# ---- kernel 8: persistent, ADO-to-ADO synchronised propagation (csrc/heom_dataflow.cuh) ----
@pytest.mark.parametrize("name", ["deom_polariton32_L6", "deom_polariton32_L2", "deom_polariton8_L4",
                                  "deom_fmo_K7_L4", "deom_spin_boson_L10", "deom_fmo_K21_L2"])
def test_dataflow_kernel_matches_reference(name):
    """All steps in one cooperative launch, stages ordered by per-ADO release/acquire flags instead
    of kernel boundaries: the reference's trajectory and every final ADO (N = 32 at full depth 6,
    and smaller systems forced through the same kernel)."""
    g = golden(name)
    s = _solver_from(g)
    s.tuning = dict(kernel=8, warps_per_cta=0, use_graph=0)
    _check_against_golden(g, s)
    assert s._plan.info("dataflow_launches") == 1 and s._plan.info("resident_launches") == 0


# ---- kernel 9: the same scheme for Hermitian problems, one CTA per ADO (csrc/heom_dataflow_tma.cuh) ----
@pytest.mark.parametrize("name", ["deom_polariton32_L6", "deom_polariton32_L2", "deom_polariton8_L4",
                                  "deom_spin_boson_L10", "deom_fmo_K21_L2"])
def test_dataflow_tma_kernel_matches_reference(name):
    """Kernel 9: y and the accumulator in registers for the whole run, k = W + W^dagger, neighbour
    matrices by bulk copies behind per-link flags; N = 32 at full depth 6 (210 CTAs on 148 SMs: the
    per-SM placement with spare CTAs) and smaller systems forced through the same kernel."""
    g = golden(name)
    s = _solver_from(g)
    s.tuning = dict(kernel=9, warps_per_cta=0, use_graph=0)
    _check_against_golden(g, s)
    assert s._plan.info("dataflow_tma_launches") == 1 and s._plan.info("resident_launches") == 0


def test_dataflow_tma_kernel_repeated_runs_and_fallback():
    """Two propagations in a row continue from the evolved state (bit-identical to kernel 8's
    restart behaviour within rounding); a non-Hermitian initial state is not taken by kernel 9."""
    g = golden("deom_polariton32_L2")
    n, dt, nt = g["rho0"].shape[0], float(g["dt"]), int(g["nt"])
    a = _solver_from(g)
    a.tuning = dict(kernel=9, warps_per_cta=0, use_graph=0)
    b = _solver_from(g)
    b.tuning = dict(kernel=2, warps_per_cta=0, use_graph=0)
    for s in (a, b):
        s.run(g["rho0"].copy(), dt, nt)
    assert a._plan.info("dataflow_tma_launches") == 1
    assert np.max(np.abs(a.ddos - b.ddos)) < 1e-13
    rng = np.random.default_rng(11)
    r = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))       # not Hermitian
    c = _solver_from(g)
    _, out = c.run(r.copy(), dt, nt)
    assert c._plan.info("dataflow_tma_launches") == 0 and c._plan.info("dataflow_launches") == 1
    d = _solver_from(g)
    d.tuning = dict(kernel=2, warps_per_cta=0, use_graph=0)
    _, ref = d.run(r.copy(), dt, nt)
    assert np.max(np.abs(np.asarray(out) - np.asarray(ref))) < 1e-12


def test_dataflow_tma_kernel_long_run_agrees_with_kernel8():
    """Config 4 as benchmarked (N = 32, 210 ADOs) over 1500 RK4 steps: kernel 9 (k = W + W^dagger, packed
    units, S from four real sums) against kernel 8 (the reference's product order) on every final ADO;
    rho_sys stays Hermitian with unit trace."""
    from pyqed_b200 import workloads as W
    from pyqed_b200.heom import DEOMSolver, Bath
    w = W.polariton(lmax=6)
    bath = Bath(expn=w["expn"], etal=w["etal"], etar=w["etar"], etaa=w["etaa"], mode=w["mode"])
    out = {}
    for kern in (9, 8):
        s = DEOMSolver(w["system"], w["system_dipole"], bath, w["coupling"], w["coupling_dipole"], lmax=w["lmax"])
        s.tuning = dict(kernel=kern, warps_per_cta=0, use_graph=0)
        _, traj = s.run(w["rho0"].copy(), w["dt"], 1500)
        assert s._plan.info("dataflow_tma_launches") == (1 if kern == 9 else 0)
        out[kern] = (np.asarray(traj[-1]), s.ddos.copy())
    assert np.max(np.abs(out[9][1] - out[8][1])) < 1e-12
    rho = out[9][0]
    assert abs(np.trace(rho) - 1) < 1e-12 and np.array_equal(rho, rho.conj().T)


def test_dense_operators_at_n32_against_the_oracle():
    """SURVEY 8d's stress variant of config 4: N = 32 with a dense random Hermitian H (seed 0) and the
    polariton couplings.  Dense rows do not fit kernel 9's operator table: the automatic choice is its
    DENSE_H instantiation (A = -iH in parameter space, the own term as a matrix product by rows); it,
    kernel 8 and the per-stage generic kernel must match the oracle (test infrastructure, here as the
    checker) on the trajectory and on every ADO."""
    from oracle.deom_oracle import DeomOracle
    from pyqed_b200 import workloads as W
    from pyqed_b200.heom import DEOMSolver, Bath
    w = W.polariton(lmax=2)
    n = w["system"].shape[0]
    rng = np.random.default_rng(0)
    a = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    H = 0.1 * (a + a.conj().T) / 2 + w["system"]
    nt, dt = 8, w["dt"]
    o = DeomOracle(H, w["system_dipole"], w["coupling"], w["coupling_dipole"], w["expn"], w["etal"], w["etar"],
                   w["etaa"], w["mode"], w["lmax"])
    _, ref = o.run(w["rho0"], dt, nt)
    bath = Bath(expn=w["expn"], etal=w["etal"], etar=w["etar"], etaa=w["etaa"], mode=w["mode"])
    for kern in (0, 8, 2):
        s = DEOMSolver(H, w["system_dipole"], bath, w["coupling"], w["coupling_dipole"], lmax=w["lmax"])
        s.tuning = dict(kernel=kern, warps_per_cta=0, use_graph=0)
        _, got = s.run(w["rho0"].copy(), dt, nt)
        assert np.max(np.abs(np.asarray(got) - np.asarray(ref))) < TOL
        assert np.max(np.abs(s.ddos - o.ddos)) < TOL
        if kern == 0:
            assert s._plan.info("dataflow_dense_launches") == 1
        if kern == 8:
            assert s._plan.info("dataflow_launches") == 1 and s._plan.info("dataflow_tma_launches") == 0


def test_dense_h_instantiation_at_full_depth_agrees_with_kernel8():
    """The dense-H variant of config 4 at depth 6 (210 ADOs on 148 SMs), 200 steps: kernel 9<DENSE_H>
    against kernel 8 on every ADO."""
    from pyqed_b200 import workloads as W
    from pyqed_b200.heom import DEOMSolver, Bath
    w = W.polariton(lmax=6, dense_h=True)
    bath = Bath(expn=w["expn"], etal=w["etal"], etar=w["etar"], etaa=w["etaa"], mode=w["mode"])
    out = {}
    for kern in (0, 8):
        s = DEOMSolver(w["system"], w["system_dipole"], bath, w["coupling"], w["coupling_dipole"], lmax=w["lmax"])
        s.tuning = dict(kernel=kern, warps_per_cta=0, use_graph=0)
        s.run(w["rho0"].copy(), w["dt"], 200)
        assert s._plan.info("dataflow_dense_launches") == (1 if kern == 0 else 0)
        out[kern] = s.ddos.copy()
    assert np.max(np.abs(out[0] - out[8])) < 1e-12


def test_dataflow_tma_kernel_handles_batches():
    """Several trajectories (different Hermitian initial states) in one kernel-9 launch: 3 x 15 ADOs,
    each against its own per-stage run of the generic kernel."""
    g = golden("deom_polariton32_L2")
    n, nt, dt = g["rho0"].shape[0], int(g["nt"]), float(g["dt"])
    rng = np.random.default_rng(7)
    rhos = []
    for _ in range(3):
        a = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        r = a @ a.conj().T
        rhos.append(r / np.trace(r))
    b = _solver_from(g)
    _, out = b.run_batch(rhos, dt, nt)
    assert b._plan.info("dataflow_tma_launches") == 1
    for i, r in enumerate(rhos):
        one = _solver_from(g)
        one.tuning = dict(kernel=2, warps_per_cta=0, use_graph=0)
        _, ref = one.run(r.copy(), dt, nt)
        assert np.max(np.abs(out[i] - np.asarray(ref))) < 1e-13


def test_dataflow_kernel_is_the_default_for_config4_and_handles_batches():
    g = golden("deom_polariton32_L6")
    s = _solver_from(g)
    _check_against_golden(g, s)
    assert s._plan.info("dataflow_launches") == 1          # chosen automatically for N = 32
    assert s._plan.info("dataflow_tma_launches") == 1      # Hermitian problem, 210 ADOs <= 2 CTAs per SM: kernel 9
    # a batch of trajectories (different initial states) in the same launch, twice in a row
    n, nt, dt = g["rho0"].shape[0], 12, float(g["dt"])
    rng = np.random.default_rng(5)
    rhos = []
    for _ in range(3):
        a = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        r = a @ a.conj().T
        rhos.append(r / np.trace(r))
    b = _solver_from(g)
    _, out = b.run_batch(rhos, dt, nt)
    assert b._plan.info("dataflow_launches") == 1
    for i, r in enumerate(rhos):
        one = _solver_from(g)
        one.tuning = dict(kernel=2, warps_per_cta=0, use_graph=0)     # per-stage launches of the generic kernel
        _, ref = one.run(r.copy(), dt, nt)
        assert np.max(np.abs(out[i] - np.asarray(ref))) < 1e-13
    _, out2 = b.run_batch(rhos, dt, nt)
    assert np.array_equal(out, out2)


def test_waiting_time_scan_batch_matches_reference_at_depth_6():
    """BASELINE configs[4] at its full depth (924 ADOs): two waiting times of the pump/probe scan in
    one ``run_batch`` against the reference's two separate ``DEOMSolver.run`` calls."""
    gs = [golden("deom_aggregate_L6_T0"), golden("deom_aggregate_L6_T37")]
    g = gs[0]
    dt, nt = float(g["dt"]), int(g["nt"])
    from pyqed_b200.heom import DEOMSolver, Bath
    bath = Bath(expn=g["expn"], etal=g["etal"], etar=g["etar"], etaa=g["etaa"], mode=g["mode"])
    s = DEOMSolver(g["system"], g["system_dipole"], bath, g["coupling"], g["coupling_dipole"], lmax=int(g["lmax"]))
    fields = [pulse_from_samples(x["pulse_system"], dt) for x in gs]
    ts, sig = s.run_batch([g["rho0"], g["rho0"]], dt, nt, p1=g["p1"], pulse_system_funcs=fields)
    for i, x in enumerate(gs):
        assert np.max(np.abs(sig[i] - x["traj"])) < TOL
    assert np.max(np.abs(sig[0] - sig[1])) > 1e-4      # the two waiting times really differ
