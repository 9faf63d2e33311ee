"""Cavity-molecule model builders and the Result container.

The Hamiltonians are pinned to outputs of the reference's ``Cavity`` /
``Polariton.getH`` (``tests/golden/polariton.npz``, made by
``tests/golden/make_golden.py polariton``)."""
import os
import pickle

import numpy as np
import pytest

from conftest import golden
from pyqed_b200.mol import Mol, Result
from pyqed_b200.polariton import Cavity, Polariton, Env, env_bath, ham_ho, heom


def _case(g, tag):
    wc, ncav, coup = g[f"{tag}_params"]
    mol = Mol(g[f"{tag}_hmol"], g[f"{tag}_edip"], lowering=g[f"{tag}_lowering"])
    cav = Cavity(float(wc), int(ncav))
    return mol, cav, Polariton(mol, cav, g=float(coup))


@pytest.mark.parametrize("tag", ["two_level", "ladder3"])
def test_polariton_hamiltonian_matches_reference(tag):
    g = golden("polariton")
    mol, cav, pol = _case(g, tag)
    for rwa in (0, 1):
        assert np.array_equal(pol.getH(RWA=bool(rwa)), g[f"{tag}_H_rwa{rwa}"])
    assert np.array_equal(cav.create(), g[f"{tag}_create"])
    assert np.array_equal(cav.annihilate(), g[f"{tag}_annihilate"])
    assert np.array_equal(cav.num(), g[f"{tag}_num"])
    assert np.array_equal(cav.getH(), g[f"{tag}_hcav"])
    assert np.array_equal(cav.get_dm(), g[f"{tag}_vacuum_dm"])
    H = pol.getH(RWA=False)
    assert np.max(np.abs(H - H.conj().T)) == 0.0
    assert pol.dim == mol.dim * cav.n_cav and pol.dims == [mol.dim, cav.n_cav]


def test_config4_workload_is_the_reference_model():
    """BASELINE config 4's H is Polariton.getH(RWA=False) of a two-level molecule."""
    from pyqed_b200 import workloads as W
    g = golden("polariton")
    w = W.polariton()
    assert np.max(np.abs(w["system"] - g["two_level_H_rwa0"])) < 1e-14
    mol, cav, pol = _case(g, "two_level")
    sz = np.diag([1.0, -1.0])
    assert np.array_equal(w["coupling"][0], pol.promote_op(sz, "mol"))
    assert np.array_equal(w["coupling"][1], pol.promote_op(cav.quadrature(), "cav"))
    assert np.array_equal(w["rho0"], pol.get_dm())


def test_builder_errors():
    mol = Mol(np.diag([0.0, 1.0]), np.array([[0, 1], [1, 0]]))
    with pytest.raises(ValueError):
        Cavity(1.0, 0)
    pol = Polariton(mol, Cavity(1.0, 3))
    with pytest.raises(ValueError):
        pol.getH()                      # g not set
    pol.g = 0.1
    with pytest.raises(ValueError):
        pol.getH(RWA=True)              # no lowering operator given
    with pytest.raises(NotImplementedError):
        Polariton(mol, Cavity(1.0, 3), g=0.1, gauge="velocity")
    with pytest.raises(TypeError):
        Polariton(object(), Cavity(1.0, 3))
    assert np.array_equal(ham_ho(2.0, 3), np.diag([0, 2.0, 4.0]))


def test_env_bath_is_the_high_temperature_chain():
    """One bath: eta = lambda (2 kT - i gamma) is the D0 of HEOM/heom.py:312."""
    from oracle import chain_oracle as CO
    env = Env(3.0, [0.7], [0.2])
    b = env_bath(env)
    d0 = CO.d0_high_temperature(3.0, 0.7, 0.2)
    expn, etal, etar, etaa, mode, _ = CO.chain_as_deom(d0, 0.7, 5)
    assert np.allclose(b.expn, expn) and np.allclose(b.etal, etal)
    assert np.allclose(b.etar, etar) and np.allclose(b.etaa, etaa)
    assert list(b.mode) == list(mode)
    with pytest.raises(ValueError):
        env_bath(Env(1.0, [1.0, 2.0], [0.1]))


def test_result_container(tmp_path):
    r = Result(description="x", rho0=np.eye(2), dt=0.5, Nt=10, nout=2)
    assert np.allclose(r.times, [0, 1, 2, 3, 4, 5]) and r.nt == r.timesteps == 10
    assert r.expect() is None
    with pytest.raises(ValueError):
        r.analyze()
    r.observables = np.arange(12.0).reshape(6, 2)
    r.solver = object()
    path = os.path.join(tmp_path, "r.pkl")
    r.save(path)
    assert r.solver is not None
    with open(path, "rb") as f:
        back = pickle.load(f)
    assert np.array_equal(back.expect(), r.observables) and not hasattr(back, "solver")


def _two_bath_model():
    sz, sx = np.diag([1.0, -1.0]), np.array([[0, 1.0], [1.0, 0]])
    mol = Mol(0.5 * sz, sx, lowering=np.array([[0, 0], [1.0, 0]]))
    cav = Cavity(1.0, 4)
    pol = Polariton(mol, cav, g=0.1)
    hs = pol.getH(RWA=False)
    env = Env(1.5, [1.0, 0.6], [0.05, 0.08])
    env.set_c_ops([pol.promote_op(sz, "mol"), pol.promote_op(cav.quadrature(), "cav")])
    obs_ops = [pol.promote_op(np.diag([1.0, 0.0]), "mol"), pol.promote_op(cav.num(), "cav"),
               pol.promote_op(sx, "mol")]
    return pol, env, hs, obs_ops


def _oracle_observables(env, hs, rho0, obs_ops, nt, dt, lmax):
    from oracle.deom_oracle import DeomOracle
    b = env_bath(env)
    o = DeomOracle(hs, np.zeros_like(hs), env.c_ops, None, b.expn, b.etal, b.etar, b.etaa, b.mode, lmax)
    _, traj = o.run(rho0, dt, nt)
    traj = np.asarray(traj)
    return np.array([[np.trace(a @ r) for a in obs_ops] for r in traj]), traj


def test_two_bath_heom_glue_on_cpu(monkeypatch):
    """``polariton.heom`` with the CUDA solver replaced by the oracle: checks the
    model -> Bath mapping, the observable contraction and the Result layout."""
    from oracle.deom_oracle import DeomOracle
    import pyqed_b200.heom.deom as D

    class FakeSolver:
        def __init__(self, system, system_dipole, bath, coupling, lmax=None, **kw):
            self.o = DeomOracle(system, system_dipole, coupling, None, bath.expn, bath.etal, bath.etar,
                                bath.etaa, bath.mode, lmax)

        def run(self, rho0, dt, nt):
            return self.o.run(rho0, dt, nt)

    monkeypatch.setattr(D, "DEOMSolver", FakeSolver)
    pol, env, hs, obs_ops = _two_bath_model()
    rho0 = pol.get_dm()
    res = heom(env, hs, rho0, obs_ops, 12, 0.01, lmax=3)
    ref, traj = _oracle_observables(env, hs, rho0, obs_ops, 12, 0.01, 3)
    assert res.observables.shape == (13, 3) and np.max(np.abs(res.observables - ref)) < 1e-14
    assert np.array_equal(res.rho, traj[-1]) and np.allclose(res.times, 0.01 * np.arange(13))
    assert abs(res.observables[0, 0] - 1.0) < 1e-15      # starts in |e, 0>
    with pytest.raises(ValueError):
        heom(Env(1.0, [1.0], [0.1]), hs, rho0, obs_ops, 1, 0.01)


@pytest.mark.gpu
def test_two_bath_heom_on_gpu():
    pol, env, hs, obs_ops = _two_bath_model()
    rho0 = pol.get_dm()
    res = heom(env, hs, rho0, obs_ops, 40, 0.01, lmax=4)
    ref, traj = _oracle_observables(env, hs, rho0, obs_ops, 40, 0.01, 4)
    assert np.max(np.abs(res.observables - ref)) < 1e-12
    assert np.max(np.abs(res.rholist - traj)) < 1e-12
    # the same model through the factory of the composite object
    s = pol.deom(env_bath(env), env.c_ops, lmax=4)
    _, traj2 = s.run(rho0.copy(), 0.01, 40)
    assert np.max(np.abs(np.asarray(traj2) - traj)) < 1e-12
