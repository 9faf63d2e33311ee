"""Tiny propagation used under compute-sanitizer (racecheck is slow): the async row
kernel, the plain-load row kernel, both resident kernels, the generic kernel, kernels 6 and 7
(full and packed Hermitian storage) and the persistent dataflow kernels 8 and 9 on a 36-ADO
hierarchy, two RK4 steps each, checked against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.deom_oracle import DeomOracle  # noqa: E402
from pyqed_b200 import workloads as W  # noqa: E402
from pyqed_b200.heom import DEOMSolver, Bath  # noqa: E402

w = W.fmo(lmax=2, n_matsubara=0)
o = DeomOracle(w["system"], w["system_dipole"], w["coupling"], w["coupling_dipole"], w["expn"], w["etal"],
               w["etar"], w["etaa"], w["mode"], w["lmax"])
_, ref = o.run(w["rho0"], w["dt"], 2)
bath = Bath(expn=w["expn"], etal=w["etal"], etar=w["etar"], etaa=w["etaa"], mode=w["mode"])
for tuning, opts in [(dict(kernel=3), {"resident": 0}), (dict(kernel=1), {"resident": 0}),
                     (dict(kernel=2), {"resident": 0}), (dict(kernel=0), {"resident": 1}),
                     (dict(kernel=0), {"resident": 4}), (dict(kernel=6), {"resident": 0}),
                     (dict(kernel=7), {"resident": 0}), (dict(kernel=8), {}), (dict(kernel=9), {})]:
    s = DEOMSolver(w["system"], w["system_dipole"], bath, w["coupling"], w["coupling_dipole"], lmax=w["lmax"])
    s.tuning = dict(kernel=tuning["kernel"], warps_per_cta=0, use_graph=0)
    s.options = opts
    _, got = s.run(w["rho0"].copy(), w["dt"], 2)
    err = np.max(np.abs(np.asarray(got) - np.asarray(ref)))
    assert err < 1e-12, (tuning, opts, err)
    print("ok", tuning, opts, err)
