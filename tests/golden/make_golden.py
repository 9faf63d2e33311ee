"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):

    python tests/golden/make_golden.py

The reference cannot be imported as a package here, so
the three entry points on the path are loaded by file:

* ``pyqed/heom/deom.py``  -> ``DEOMSolver``, ``Bath``, decompositions (spec loader)
* ``pyqed/HEOM/heom.py``  -> ``_heom`` RK4 chain  (AST-extracted, helpers from phys.py)
* ``pyqed/oqs.py``        -> ``_heom`` Euler chain (AST-extracted)

Nothing from the reference is written into the repo except its numerical
outputs.  Inputs come from ``pyqed_b200.workloads`` so that the oracle, the
CUDA path and the reference all see the same arrays; the inputs are stored in
each fixture too, so the tests do not depend on the builders staying unchanged.
"""
from __future__ import annotations

import ast
import importlib.util
import io
import contextlib
import os
import sys
import time
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from pyqed_b200 import workloads as W  # noqa: E402


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ref_deom = _load("ref_deom", f"{REF}/pyqed/heom/deom.py")
ref_deom.tqdm = lambda x: x
ref_phys = _load("ref_phys", f"{REF}/pyqed/phys.py")


def _extract(path, fname, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == fname:
            code = compile(ast.Module([node], []), path, "exec")
            exec(code, ns)
            return ns[fname]
    raise KeyError(fname)


_ns = dict(np=np, comm=ref_phys.comm, commutator=ref_phys.commutator,
           anticommutator=ref_phys.anticommutator, rk4=ref_phys.rk4, coth=ref_phys.coth,
           obs=lambda rho, a: np.vdot(ref_phys.dag(a).ravel(), rho))
ref_heom_rk4 = _extract(f"{REF}/pyqed/HEOM/heom.py", "_heom", dict(_ns))
ref_heom_euler = _extract(f"{REF}/pyqed/oqs.py", "_heom", dict(_ns))


def run_ref_deom(w, nt, p1=None, keep_ados=True):
    bath = SimpleNamespace(expn=w["expn"].copy(), etal=w["etal"].copy(), etar=w["etar"].copy(),
                           etaa=w["etaa"].copy(), mode=w["mode"].copy())
    f = w["pulse_system_func"] or (lambda t: 0.0)
    g = w["pulse_coupling_func"] or (lambda t: 0.0)
    s = ref_deom.DEOMSolver(system=w["system"].copy(), system_dipole=w["system_dipole"].copy(),
                            bath=bath, coupling=w["coupling"].copy(),
                            coupling_dipole=w["coupling_dipole"].copy(),
                            pulse_system_func=f, pulse_coupling_func=g, lmax=w["lmax"])
    t0 = time.time()
    ts, dd = s.run(w["rho0"].copy(), w["dt"], nt, p1=p1)
    wall = time.time() - t0
    if p1 is None:
        dd = np.stack([np.asarray(x) for x in dd])
    ados = np.stack([np.asarray(x) if not hasattr(x, "toarray") else x.toarray()
                     for x in s.ddos]) if keep_ados else None
    return ts, np.asarray(dd), ados, np.asarray(s.keys), wall


def save_deom(tag, w, nt, p1=None, keep_ados=True, stride=1):
    ts, dd, ados, keys, wall = run_ref_deom(w, nt, p1, keep_ados)
    f = w["pulse_system_func"]
    g = w["pulse_coupling_func"]
    # pulses sampled on the half-step grid so the fixture is self-contained
    grid = np.arange(2 * nt + 1) * (w["dt"] / 2)
    fs = np.array([f(t) for t in grid]) if f else np.zeros(2 * nt + 1)
    gs = np.array([g(t) for t in grid]) if g else np.zeros(2 * nt + 1)
    out = dict(system=w["system"], system_dipole=w["system_dipole"], coupling=w["coupling"],
               coupling_dipole=w["coupling_dipole"], expn=w["expn"], etal=w["etal"],
               etar=w["etar"], etaa=w["etaa"], mode=w["mode"], lmax=w["lmax"], rho0=w["rho0"],
               dt=w["dt"], nt=nt, t_save=ts[::stride], traj=dd[::stride], stride=stride,
               pulse_system=fs, pulse_coupling=gs, keys=keys.astype(np.uint8),
               ref_wall_s=wall)
    if p1 is not None:
        out["p1"] = np.asarray(p1, dtype=np.complex128)
    if keep_ados:
        out["ados_final"] = ados
    np.savez_compressed(os.path.join(HERE, f"deom_{tag}.npz"), **out)
    nmax = keys.shape[0]
    print(f"deom_{tag}: nmax={nmax} nt={nt} ref wall {wall:.2f}s "
          f"({nmax * nt / wall:.0f} ADO-steps/s)")


def save_chain(tag, fn, kind, nado, dt, nt):
    s0, sx, sy, sz = ref_phys.pauli()
    H = -0.5 * sx - 0.5 * sz
    rho0 = np.zeros((2, 2))
    rho0[1, 1] = 1
    with contextlib.redirect_stdout(io.StringIO()):
        obs = fn(H, rho0.copy(), [sz], [sz, sx], temperature=600, cutoff=5,
                 reorganization=0.2, nado=nado, dt=dt, nt=nt)
    np.savez_compressed(os.path.join(HERE, f"chain_{tag}.npz"), H=H, rho0=rho0, c_op=sz,
                        e_ops=np.stack([sz, sx]), temperature=600.0, cutoff=5.0,
                        reorganization=0.2, nado=nado, dt=dt, nt=nt, observables=obs, kind=kind)
    print(f"chain_{tag}: last <sz> = {obs[0, -1].real!r}")


def save_bath():
    import sympy as sp
    w_sp = sp.symbols("omega", real=True)
    out = {}
    cases = {
        "drude_m1": (2 * 0.2 * 1.0 * w_sp / (1.0 ** 2 + w_sp ** 2), 1.0, 1, 0),
        "drude_m2": (2 * 6.593 * 20.0 * w_sp / (20.0 ** 2 + w_sp ** 2), 1 / 39.276, 2, 0),
        "drude_p1": (2 * 0.05 * 1.0 * w_sp / (1.0 ** 2 + w_sp ** 2), 1.0, 1, 1),
        "drude_p2": (2 * 1.0 * 1.0 * w_sp / (1.0 ** 2 + w_sp ** 2), 1.0, 2, 1),
        "drude_p5": (2 * 0.5 * 2.0 * w_sp / (2.0 ** 2 + w_sp ** 2), 0.7, 5, 1),
        "bo_p3": (2 * 0.3 * 0.4 * 1.5 ** 2 * w_sp / ((w_sp ** 2 - 1.5 ** 2) ** 2 + 0.4 ** 2 * w_sp ** 2),
                  0.8, 3, 1),
    }
    for name, (spe, beta, npsd, pade) in cases.items():
        etal, etar, etaa, expn = ref_deom.decompose_spectrum_pade(spe, w_sp, beta, npsd, pade=pade)
        out[f"{name}_etal"], out[f"{name}_etar"] = np.asarray(etal), np.asarray(etar)
        out[f"{name}_etaa"], out[f"{name}_expn"] = np.asarray(etaa), np.asarray(expn)
        out[f"{name}_args"] = np.array([beta, npsd, pade], dtype=float)
    for n in (1, 2, 3, 6):
        p, r = ref_deom.pade_approximation_distribution(n, 1, 1)
        out[f"psd_pole_{n}"], out[f"psd_resi_{n}"] = np.asarray(p), np.asarray(r)
        for pade in (2, 3):   # [N/N] and [N+1/N] (deom.py:119-207)
            p, r = ref_deom.pade_approximation_distribution(n, 1, pade)
            out[f"psd{pade}_pole_{n}"], out[f"psd{pade}_resi_{n}"] = np.asarray(p), np.asarray(r)
    more = {
        "drude_q2": (cases["drude_p2"][0], 1.0, 2, 2),
        "drude_q3": (cases["drude_p2"][0], 1.0, 3, 3),
        "bo_q2": (cases["bo_p3"][0], 0.8, 3, 2),
        "bo_q3": (cases["bo_p3"][0], 0.8, 2, 3),
    }
    for name, (spe, beta, npsd, pade) in more.items():
        etal, etar, etaa, expn = ref_deom.decompose_spectrum_pade(spe, w_sp, beta, npsd, pade=pade)
        out[f"{name}_etal"], out[f"{name}_etar"] = np.asarray(etal), np.asarray(etar)
        out[f"{name}_etaa"], out[f"{name}_expn"] = np.asarray(etaa), np.asarray(expn)
        out[f"{name}_args"] = np.array([beta, npsd, pade], dtype=float)
    # poles of J without the Bose factor (deom.py:310-425)
    for name, fn, spe in (("real_drude", ref_deom.decompose_spectrum_pade_real, cases["drude_p2"][0]),
                          ("imag_bo", ref_deom.decompose_spectrum_pade_imag, cases["bo_p3"][0])):
        etal, etar, etaa, expn = fn(spe, w_sp)
        out[f"{name}_etal"], out[f"{name}_etar"] = np.asarray(etal), np.asarray(etar)
        out[f"{name}_etaa"], out[f"{name}_expn"] = np.asarray(etaa), np.asarray(expn)
    # Prony refits of a Drude bath (deom.py:428-543); monic denominator (gamma = 1)
    prony = {"prony_3": (3, dict(scale=20, n=200, npsd=4)), "prony_4": (4, dict(scale=30, n=400, npsd=6)),
             "prony_3a": ([3, "a"], dict(scale=20, n=200, npsd=4))}
    spe = 2 * 0.5 * 1.0 * w_sp / (1.0 ** 2 + w_sp ** 2)
    for name, (nind, kw) in prony.items():
        with contextlib.redirect_stdout(io.StringIO()):
            etal, etar, etaa, expn = ref_deom.decompose_spectrum_prony(
                spe, w_sp, 0.7, list(nind) if isinstance(nind, list) else nind, **kw)
        out[f"{name}_etal"], out[f"{name}_etar"] = np.asarray(etal), np.asarray(etar)
        out[f"{name}_etaa"], out[f"{name}_expn"] = np.asarray(etaa), np.asarray(expn)
    etal, etar, etaa, expn = ref_deom.single_oscillator(1.3, w_sp, 0.9, 2)
    out["so_etal"], out["so_etar"], out["so_etaa"], out["so_expn"] = etal, etar, etaa, expn
    np.savez_compressed(os.path.join(HERE, "bath.npz"), **out)
    print("bath: saved", len(out), "arrays")


def save_propagators():
    """Liouville-space Euler propagators: HEOM/heom.py:349-413 and oqs.py:1877-1941."""
    from scipy.sparse import kron, identity
    sup_ns = dict(np=np, kron=kron, identity=identity)
    o2s = _extract(f"{REF}/pyqed/superoperator.py", "operator_to_superoperator", sup_ns)
    au2k = 315775.13  # pyqed/units.py:6
    ns = dict(np=np, operator_to_superoperator=o2s, au2k=au2k)
    prop_heom = _extract(f"{REF}/pyqed/HEOM/heom.py", "_heom_propagator", dict(ns))
    prop_oqs = _extract(f"{REF}/pyqed/oqs.py", "_heom_propagator", dict(ns))
    s0, sx, sy, sz = ref_phys.pauli()
    H = -0.5 * sx - 0.5 * sz
    out = dict(H=H, c_op=sz, temperature=300.0 * 100, cutoff=5.0, reorganization=0.2, nado=5, dt=0.001,
               nt=40, au2k=au2k)
    for tag, fn in (("heom", prop_heom), ("oqs", prop_oqs)):
        with contextlib.redirect_stdout(io.StringIO()):
            u = fn(H, [sz], [sz], out["temperature"], out["cutoff"], out["reorganization"],
                   out["nado"], out["dt"], out["nt"])
        out[f"u_{tag}"] = np.asarray(u)
    # a 3-level case with a complex Hermitian H and non-diagonal coupling
    rng = np.random.default_rng(5)
    A = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
    H3 = (A + A.conj().T) / 2
    S3 = np.diag([1.0, 0.0, -1.0]).astype(complex)
    S3[0, 1] = S3[1, 0] = 0.3
    with contextlib.redirect_stdout(io.StringIO()):
        u3 = prop_heom(H3, [S3], [S3], 2.0e5, 3.0, 0.1, 6, 0.002, 25)
    out.update(H3=H3, S3=S3, u3_heom=np.asarray(u3), t3=np.array([2.0e5, 3.0, 0.1, 6, 0.002, 25]))
    np.savez_compressed(os.path.join(HERE, "chain_propagator.npz"), **out)
    print("chain_propagator: |u_heom[0]| max", np.abs(out["u_heom"][0]).max(),
          "diff heom/oqs", np.abs(out["u_heom"] - out["u_oqs"]).max())


def save_generator():
    """Dense HEOM generator (gen_generate_propgator, deom.py:1116-1125) for small cases."""
    out = {}
    for tag, w in (("random3_K1", W.random_dense(3, 1, 1, 5, seed=2, hermitian=True)),
                   ("random4_herm", W.random_dense(4, 2, 3, 2, seed=0, hermitian=True)),
                   ("spin_boson_L3", W.spin_boson(lmax=3))):
        bath = SimpleNamespace(expn=w["expn"].copy(), etal=w["etal"].copy(), etar=w["etar"].copy(),
                               etaa=w["etaa"].copy(), mode=w["mode"].copy())
        s = ref_deom.DEOMSolver(system=w["system"].copy(), system_dipole=w["system_dipole"].copy(),
                                bath=bath, coupling=w["coupling"].copy(),
                                coupling_dipole=w["coupling_dipole"].copy(),
                                pulse_system_func=lambda t: 0.0, pulse_coupling_func=lambda t: 0.0,
                                lmax=w["lmax"])
        s.gen_generate_propgator()
        for k in ("system", "coupling", "expn", "etal", "etar", "etaa", "mode", "rho0"):
            out[f"{tag}_{k}"] = w[k]
        out[f"{tag}_lmax"] = w["lmax"]
        out[f"{tag}_generator"] = np.asarray(s.propgator)
        print("generator", tag, s.propgator.shape, np.abs(s.propgator).max())
        if tag in ("spin_boson_L3", "random3_K1"):
            # frequency-domain four-operator response (correlation_4op_3t, deom.py:1127-1210)
            n = w["system"].shape[0]
            rng = np.random.default_rng(9)
            ops = [rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) for _ in range(4)]
            w_x, w_y = np.linspace(0.3, 2.3, 5), np.linspace(-2.1, -0.2, 4)  # away from the zero mode
            ref_deom.tqdm = lambda x: x
            for lcr in ("llll", "lrcl"):
                with contextlib.redirect_stdout(io.StringIO()):
                    s.Δ = None
                    c = s.correlation_4op_3t(ops[0], ops[1], ops[2], ops[3], w["rho0"], 0.7, w_x, w_y, lcr=lcr)
                out[f"{tag}_c4_{lcr}"] = np.asarray(c)
            out[f"{tag}_c4_ops"] = np.stack(ops)
            out[f"{tag}_c4_wx"], out[f"{tag}_c4_wy"] = w_x, w_y
    np.savez_compressed(os.path.join(HERE, "generator.npz"), **out)


def _extract_class_member(path, cls, member, ns):
    """Compile one method of a reference class as a plain function (``member`` None:
    the whole class)."""
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            if member is None:
                exec(compile(ast.Module([node], []), path, "exec"), ns)
                return ns[cls]
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == member:
                    exec(compile(ast.Module([sub], []), path, "exec"), ns)
                    return ns[member]
    raise KeyError((cls, member))


def save_polariton():
    """Hamiltonians of the reference's cavity-molecule model builder
    (``Cavity``, ``Polariton.getH``; ``pyqed/polariton/cavity.py:404-678``) for a
    two-level molecule and for a three-level ladder, RWA on and off."""
    import warnings
    from scipy.sparse import identity, kron, lil_matrix
    path = f"{REF}/pyqed/polariton/cavity.py"
    ns = dict(np=np, identity=identity, kron=kron, lil_matrix=lil_matrix, dag=ref_phys.dag,
              ket2dm=ref_phys.ket2dm, Mol=object)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _extract(path, "ham_ho", ns)
        Cavity = _extract_class_member(path, "Cavity", None, ns)
        getH = _extract_class_member(path, "Polariton", "getH", ns)
    out = {}
    cases = {
        "two_level": (np.diag([0.5, -0.5]).astype(complex), np.array([[0, 1], [1, 0]], dtype=complex),
                      np.array([[0, 0], [1, 0]], dtype=complex), 1.0, 16, 0.1),
        "ladder3": (np.diag([0.0, 0.9, 2.1]).astype(complex),
                    np.array([[0, 1, 0.2], [1, 0, 0.7], [0.2, 0.7, 0]], dtype=complex),
                    np.array([[0, 1, 0.2], [0, 0, 0.7], [0, 0, 0]], dtype=complex), 0.95, 5, 0.3),
    }
    for tag, (hmol, edip, lowering, wc, ncav, g) in cases.items():
        mol = SimpleNamespace(getH=lambda h=hmol: h, edip=edip, idm=identity(hmol.shape[0]),
                              lowering=lowering, raising=lowering.conj().T, dim=hmol.shape[0])
        cav = Cavity(wc, ncav)
        for rwa in (False, True):
            me = SimpleNamespace(mol=mol, cav=cav, _g=g, gauge="length", H=None)
            H = getH(me, RWA=rwa)
            H = H.toarray() if hasattr(H, "toarray") else np.asarray(H)
            out[f"{tag}_H_rwa{int(rwa)}"] = H
        out[f"{tag}_hmol"], out[f"{tag}_edip"], out[f"{tag}_lowering"] = hmol, edip, lowering
        out[f"{tag}_params"] = np.array([wc, ncav, g])
        out[f"{tag}_create"] = cav.create().toarray()
        out[f"{tag}_annihilate"] = cav.annihilate().toarray()
        out[f"{tag}_num"] = cav.get_number_operator().toarray()
        out[f"{tag}_hcav"] = np.asarray(cav.getH())
        out[f"{tag}_vacuum_dm"] = np.asarray(cav.get_dm())
    np.savez_compressed(os.path.join(HERE, "polariton.npz"), **out)
    print("polariton.npz", sorted(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "polariton":
        save_polariton()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "propagators":
        save_propagators()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "aggregate_L6":
        # config 5 at full depth (924 ADOs), pump/probe field on, observable Tr(mu rho); the reference
        # needs ~5 minutes per waiting time.  Optional second argument: one waiting index.
        for b in ([int(sys.argv[2])] if len(sys.argv) > 2 else [0, 37]):
            w = W.aggregate_2des(lmax=6, waiting_index=b)
            save_deom(f"aggregate_L6_T{b}", w, 400, p1=w["observable"], keep_ados=False, stride=1)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "polariton_L6":
        # config 4 at full depth (N = 32, 210 ADOs): trajectory and every final ADO
        save_deom("polariton32_L6", W.polariton(lmax=6), 40)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "bath":
        save_bath()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "generator":
        save_generator()
        return
    save_propagators()
    save_bath()
    # KAT-1 / KAT-1b / KAT-1e: examples/heom.py inputs
    save_chain("rk4_nado5", ref_heom_rk4, "rk4", 5, 0.02, 100)
    save_chain("rk4_nado12", ref_heom_rk4, "rk4", 12, 0.01, 200)
    save_chain("euler_nado5", ref_heom_euler, "euler", 5, 0.02, 100)
    # config 1 (DEOM form), depth 10, shortened trajectory
    save_deom("spin_boson_L10", W.spin_boson(lmax=10), 100)
    # KAT-2: examples/deom.py inputs, population observable
    save_deom("example_L10_p1", W.spin_boson_deom_example(lmax=10), 20,
              p1=np.array([[1, 0], [0, 0]], dtype=np.complex128), keep_ados=False)
    # config 2, 60 steps (KAT-4)
    save_deom("fmo_K7_L4", W.fmo(lmax=4, n_matsubara=0), 60)
    # config 3 shape at reduced depth
    save_deom("fmo_K21_L2", W.fmo(lmax=2, n_matsubara=2), 8)
    save_deom("fmo_K21_L3", W.fmo(lmax=3, n_matsubara=2), 2)
    # config 4 shape at reduced depth (N=32, dense-ish Q, complex H)
    save_deom("polariton32_L2", W.polariton(lmax=2), 6)
    save_deom("polariton8_L4", W.polariton(lmax=4, nfock=4), 10)
    # config 5 with the pump/probe field on, two waiting times, observable Tr(mu rho)
    for b in (0, 37):
        w = W.aggregate_2des(lmax=3, waiting_index=b)
        save_deom(f"aggregate_L3_T{b}", w, 300, p1=w["observable"], keep_ados=False, stride=1)
    # stress: random dense complex inputs with both pulses on, Hermitian and not
    save_deom("random4_herm", W.random_dense(4, 2, 3, 3, seed=0, hermitian=True), 20)
    save_deom("random5_nonherm", W.random_dense(5, 3, 4, 2, seed=1, hermitian=False), 12)
    save_deom("random3_K1", W.random_dense(3, 1, 1, 5, seed=2, hermitian=True), 15)


if __name__ == "__main__":
    main()
