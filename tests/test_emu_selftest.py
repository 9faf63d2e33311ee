"""The CUDA emulation harness must reject what the hardware rejects.

``tests/_shim/cuda_emu.h`` runs kernel sources on the CPU; a plain ``memcpy`` behind
``cp.async`` / ``cp.async.bulk`` would forgive misaligned addresses, sizes that are
not a multiple of 16 bytes and shared-memory operands outside the CTA's allocation,
which fault or hang on the GPU.  ``emu_selftest.cpp`` issues each such copy once:
the legal sequence must run through, every illegal one must abort with a message.
The kernel emulation tests (kernels 3, 6, 7, hierarchy builder) run with the same
checks switched on, for every system size they cover.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ILLEGAL = {
    1: "cp.async 16: shared destination",
    2: "cp.async 16: global source",
    3: "global->shared: size",
    4: "global->shared: shared destination",
    5: "global->shared: global source",
    6: "shared->global: global destination",
    7: "shared->global: shared source",
    8: "mbarrier.init",
    9: "global->shared: size",
    10: "illegal launch",
}


@pytest.fixture(scope="module")
def selftest(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emu_selftest") / "emu_selftest")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-o", exe,
                           os.path.join(ROOT, "tests", "_shim", "emu_selftest.cpp")])
    return exe


def test_legal_copies_run_through(selftest):
    res = subprocess.run([selftest, "0"], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr


@pytest.mark.parametrize("case", sorted(ILLEGAL))
def test_illegal_operands_abort(selftest, case):
    res = subprocess.run([selftest, str(case)], capture_output=True, text=True, timeout=60)
    assert res.returncode != 0
    assert "cuda_emu:" in res.stderr and ILLEGAL[case] in res.stderr, res.stderr
