"""HEOM / DEOM solvers with the reference's interfaces (pyqed/heom, pyqed/HEOM)."""
from .bath import Bath, drude_exponents, bose_poles, rational_exponents  # noqa: F401
from .deom import DEOMSolver  # noqa: F401
from .heom import HEOMSolver  # noqa: F401
from .spectrum import (decompose_spectrum_pade, decompose_spectrum_matsubara,  # noqa: F401
                       single_oscillator)
