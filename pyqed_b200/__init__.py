"""pyqed_b200 - B200-native HEOM/DEOM time propagation behind pyqed's solver API.

Only the hot path of binggu56/pyqed's ``pyqed/heom`` is implemented (SURVEY.md
section 8); the work is done by hand-written sm_100a CUDA kernels in
``csrc/heom_kernels.cu`` reached through the C ABI of ``include/pyqed_heom.h``.
"""
__version__ = "0.1.0"
