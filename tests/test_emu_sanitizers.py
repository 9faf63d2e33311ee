"""The emulated kernels 3, 6 and 7 under AddressSanitizer and ThreadSanitizer.

``tests/_shim/emu_sanitize_main.cpp`` runs the three kernels (both timing models
of the asynchronous copies, two owned ranges, rotated visiting order) on a
problem dumped here and cross-checks their results.  Built with
``-fsanitize=address`` it reports any out-of-bounds access to the (exactly
sized) global arrays or to the emulated shared memory; built with
``-fsanitize=thread`` it reports data races between lanes, i.e. what a missing
``__syncwarp`` / ``__syncthreads`` / copy wait would cause on the GPU - a CPU
stand-in for compute-sanitizer's memcheck and racecheck.
"""
import os
import subprocess

import numpy as np
import pytest

from pyqed_b200 import workloads as W
from test_sym_kernel_emu import host_tables, projector_problem, ROOT


def dump(w, nt, path):
    o, t = host_tables(w)
    H = np.ascontiguousarray(o.H0)
    meta = np.array([t["N"], t["K"], t["M"], o.lmax, o.nmax, len(t["links"]), nt,
                     int(np.all(H.imag == 0))], dtype=np.int64)
    arrays = dict(meta=meta, dt=np.array([w["dt"]]), H=H, ops=t["ops"], cbase=t["cbase"], damp=t["damp"],
                  rho0=np.ascontiguousarray(w["rho0"], dtype=np.complex128), kmode=t["kmode"],
                  link_ptr=t["link_ptr"], links=t["links"], supp=t["supp"])
    for name, a in arrays.items():
        np.ascontiguousarray(a).tofile(os.path.join(path, name + ".bin"))


@pytest.fixture(scope="module")
def problems(tmp_path_factory):
    out = []
    for i, (w, nt) in enumerate([(W.fmo(lmax=2, n_matsubara=1), 1),            # N=7, K=14, 120 ADOs, 3 chunks
                                 (projector_problem(4, 2, 3, seed=3, complex_h=True), 1)]):  # even N, padded tiles
        d = tmp_path_factory.mktemp(f"problem{i}")
        dump(w, nt, str(d))
        out.append(str(d))
    return out


@pytest.fixture(scope="module")
def binaries(tmp_path_factory):
    """Both instrumented builds, compiled side by side."""
    from concurrent.futures import ThreadPoolExecutor
    d = tmp_path_factory.mktemp("sanitized")
    src = os.path.join(ROOT, "tests", "_shim", "emu_sanitize_main.cpp")

    def build(sanitizer):
        exe = str(d / f"emu_{sanitizer}")
        subprocess.check_call(["g++", "-O0", "-g", "-std=c++17", "-pthread", f"-fsanitize={sanitizer}",
                               "-fno-omit-frame-pointer", "-DHEOM_EMU_FEW_N", "-o", exe, src])
        return exe
    with ThreadPoolExecutor(2) as pool:
        return dict(zip(("address", "thread"), pool.map(build, ("address", "thread"))))


def run_sanitized(exe, problems):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", TSAN_OPTIONS="halt_on_error=1")
    return [subprocess.run([exe, d], capture_output=True, text=True, timeout=900, env=env) for d in problems]


@pytest.fixture(scope="module")
def runs(problems, binaries):
    """Both instrumented programs run side by side (they are independent processes)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(2) as pool:
        futures = {s: pool.submit(run_sanitized, binaries[s], problems) for s in ("address", "thread")}
        return {s: f.result() for s, f in futures.items()}


@pytest.mark.parametrize("sanitizer", ["address", "thread"])
def test_emulated_kernels_under_sanitizer(sanitizer, runs):
    for res in runs[sanitizer]:
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
        assert "emulated kernels agree" in res.stdout
        assert "Sanitizer" not in res.stderr, res.stderr[-4000:]
