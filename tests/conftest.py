import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def deom_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "deom_*.npz")))


def pulse_from_samples(samples, dt):
    """Callable t -> value reproducing a field stored on the half-step grid."""
    samples = np.asarray(samples)

    def f(t):
        return float(samples[int(round(t / (dt / 2)))])

    return f


@pytest.fixture(scope="session")
def has_cuda():
    import torch
    return torch.cuda.is_available()
