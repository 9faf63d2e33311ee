"""ctypes binding of ``include/pyqed_heom.h`` plus a thin RAII wrapper.

There is no CPU fallback: if the shared library is missing or no CUDA device
is present, ``load()`` / ``Plan`` raise.  PyTorch is used only to own device
memory and to supply the CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

_c_double_p = C.POINTER(C.c_double)
_c_int64_p = C.POINTER(C.c_int64)
_c_uint8_p = C.POINTER(C.c_uint8)
_c_size_p = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); mirrors include/pyqed_heom.h one to one
SIGNATURES = {
    "pyqed_heom_version": (C.c_int, []),
    "pyqed_heom_last_error": (C.c_char_p, []),
    "pyqed_heom_hierarchy_size": (C.c_int64, [C.c_int, C.c_int]),
    "pyqed_heom_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int]),
    "pyqed_heom_plan_destroy": (None, [C.c_void_p]),
    "pyqed_heom_set_system": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p]),
    "pyqed_heom_set_coupling": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p]),
    "pyqed_heom_set_bath": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p, _c_double_p,
                                      _c_double_p, _c_int64_p]),
    "pyqed_heom_set_order": (C.c_int, [C.c_void_p, C.c_int]),
    "pyqed_heom_table_bytes": (C.c_int, [C.c_void_p, _c_size_p]),
    "pyqed_heom_state_bytes": (C.c_int, [C.c_void_p, _c_size_p]),
    "pyqed_heom_bind": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                  C.c_void_p]),
    "pyqed_heom_build_hierarchy": (C.c_int, [C.c_void_p]),
    "pyqed_heom_get_keys": (C.c_int, [C.c_void_p, _c_uint8_p]),
    "pyqed_heom_set_state": (C.c_int, [C.c_void_p, _c_double_p]),
    "pyqed_heom_load_ados": (C.c_int, [C.c_void_p, _c_double_p]),
    "pyqed_heom_get_ados": (C.c_int, [C.c_void_p, _c_double_p]),
    "pyqed_heom_propagate": (C.c_int, [C.c_void_p, C.c_double, C.c_int64, _c_double_p, _c_double_p,
                                       C.c_void_p, C.c_int]),
    "pyqed_heom_propagate_begin": (C.c_int, [C.c_void_p, C.c_double, C.c_int64, _c_double_p, _c_double_p,
                                             C.c_void_p]),
    "pyqed_heom_propagate_stage": (C.c_int, [C.c_void_p, C.c_int64, C.c_int]),
    "pyqed_heom_set_partition": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "pyqed_heom_halo_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                       C.c_int]),
    "pyqed_heom_halo_push": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, _c_int64_p,
                                       C.POINTER(C.c_uint64), C.c_int]),
    "pyqed_heom_set_push_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                                            C.c_int]),
    "pyqed_heom_shared_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p), _c_uint8_p]),
    "pyqed_heom_shared_open": (C.c_int, [C.c_int, _c_uint8_p, C.POINTER(C.c_void_p)]),
    "pyqed_heom_shared_close": (C.c_int, [C.c_int, C.c_void_p]),
    "pyqed_heom_shared_free": (C.c_int, [C.c_int, C.c_void_p]),
    "pyqed_heom_shard_state_bytes": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _c_size_p, _c_size_p]),
    "pyqed_heom_shard_setup": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                         C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_size_t, C.POINTER(C.c_uint64), C.c_int]),
    "pyqed_heom_shard_set_state": (C.c_int, [C.c_void_p, _c_double_p]),
    "pyqed_heom_shard_get_owned": (C.c_int, [C.c_void_p, _c_double_p, C.POINTER(C.c_int32)]),
    "pyqed_heom_shard_begin": (C.c_int, [C.c_void_p, C.c_double, C.c_int64, C.c_void_p]),
    "pyqed_heom_shard_stage": (C.c_int, [C.c_void_p, C.c_int64, C.c_int]),
    "pyqed_heom_shard_end": (C.c_int, [C.c_void_p]),
    "pyqed_heom_shard_barrier": (C.c_int, [C.c_void_p]),
    "pyqed_heom_shard_propagate": (C.c_int, [C.c_void_p, C.c_double, C.c_int64, C.c_void_p]),
    "pyqed_heom_shard_error": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "pyqed_heom_apply_operator": (C.c_int, [C.c_void_p, _c_double_p, C.c_int]),
    "pyqed_heom_chain_euler": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, _c_double_p,
                                         _c_double_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_int64, C.c_int, C.c_void_p, _c_double_p, C.c_int, C.c_void_p]),
    "pyqed_heom_expectation": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, _c_double_p, C.c_int,
                                         C.c_void_p]),
    "pyqed_heom_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pyqed_heom_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pyqed_heom_synchronize": (C.c_int, [C.c_void_p]),
    "pyqed_heom_launch_count": (C.c_int64, [C.c_void_p]),
    "pyqed_heom_stage_timing": (C.c_int, [C.c_void_p, C.c_int, _c_double_p, _c_int64_p]),
    "pyqed_heom_set_tuning": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "pyqed_heom_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "pyqed_heom_get_info": (C.c_int64, [C.c_void_p, C.c_char_p]),
}


class HeomError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB


def load():
    """Load ``libpyqed_heom.so`` (building it first if sources are newer and
    nvcc is available).  Raises if it cannot be loaded."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("PYQED_HEOM_LIB") or _build.LIB  # override: tuning variants
    if path == _build.LIB and _build.is_stale():
        try:
            _build.build_extension()
        except Exception as exc:  # no nvcc on this box and no prebuilt library
            if not os.path.exists(path):
                raise HeomError(
                    f"CUDA extension {path} is missing and could not be built: {exc}") from exc
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _dptr(a):
    return None if a is None else a.ctypes.data_as(_c_double_p)


def _c128(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.complex128))
    if shape is not None:
        a = np.ascontiguousarray(np.broadcast_to(a, shape))
    return a


class Plan:
    """Owns a ``pyqed_heom_plan`` and the two torch byte buffers bound to it."""

    def __init__(self, nsys, nind, nmod, lmax, batch=1, device=0, order=0):
        import torch
        if not torch.cuda.is_available():
            raise HeomError("no CUDA device: the HEOM path has no CPU fallback")
        self.lib = load()
        self.torch = torch
        self.device = int(device)
        self.nsys, self.nind, self.nmod, self.lmax, self.batch = (int(nsys), int(nind), int(nmod),
                                                                    int(lmax), int(batch))
        h = C.c_void_p()
        self._h = None
        self._check(self.lib.pyqed_heom_plan_create(C.byref(h), self.device, self.nsys, self.nind,
                                                    self.nmod, self.lmax, self.batch))
        self._h = h
        self.nmax = int(self.lib.pyqed_heom_hierarchy_size(self.nind, self.lmax))
        self._check(self.lib.pyqed_heom_set_order(self._h, int(order)))
        self.order = int(order)
        self._tables = self._state = None
        self._keep = []

    # -- plumbing ----------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise HeomError(self.lib.pyqed_heom_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) is not None:
            self.lib.pyqed_heom_plan_destroy(self._h)
            self._h = None
        self._tables = self._state = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup -------------------------------------------------------------
    def set_system(self, H, mu=None):
        n = self.nsys
        H = _c128(H, (n, n))
        mu = None if mu is None else _c128(mu, (n, n))
        self._check(self.lib.pyqed_heom_set_system(self._h, _dptr(H.view(np.float64)),
                                                   None if mu is None else _dptr(mu.view(np.float64))))

    def set_coupling(self, Q, Qdip=None):
        n, m = self.nsys, self.nmod
        Q = _c128(Q, (m, n, n))
        Qdip = None if Qdip is None else _c128(Qdip, (m, n, n))
        self._check(self.lib.pyqed_heom_set_coupling(
            self._h, _dptr(Q.view(np.float64)), None if Qdip is None else _dptr(Qdip.view(np.float64))))

    def set_bath(self, expn, etal, etar, etaa, mode):
        k = self.nind
        arrs = [_c128(x, (k,)) for x in (expn, etal, etar, etaa)]
        mode = np.ascontiguousarray(np.asarray(mode, dtype=np.int64))
        if mode.shape != (k,):
            raise ValueError("mode must have one entry per dissipaton")
        self._check(self.lib.pyqed_heom_set_bath(self._h, *[_dptr(a.view(np.float64)) for a in arrs],
                                                 mode.ctypes.data_as(_c_int64_p)))

    def set_tuning(self, kernel=0, warps_per_cta=0, use_graph=0):
        self._check(self.lib.pyqed_heom_set_tuning(self._h, int(kernel), int(warps_per_cta),
                                                   int(use_graph)))

    def set_option(self, name, value):
        self._check(self.lib.pyqed_heom_set_option(self._h, name.encode(), int(value)))

    def info(self, name):
        return int(self.lib.pyqed_heom_get_info(self._h, name.encode()))

    def state_bytes(self):
        sb = C.c_size_t()
        self._check(self.lib.pyqed_heom_state_bytes(self._h, C.byref(sb)))
        return sb.value

    def build(self, state=None, tables_only=False):
        """Allocate the device buffers (torch), bind them and build the tables.
        ``state``: optional caller-provided uint8 tensor for the ADO arrays
        (e.g. symmetric memory shared with peer ranks).  ``tables_only``: bind no
        state buffer (sharded runs bind a rank-local one in ``shard_setup``)."""
        torch = self.torch
        tb, sb = C.c_size_t(), C.c_size_t()
        self._check(self.lib.pyqed_heom_table_bytes(self._h, C.byref(tb)))
        self._check(self.lib.pyqed_heom_state_bytes(self._h, C.byref(sb)))
        dev = torch.device("cuda", self.device)
        self._tables = torch.empty(tb.value, dtype=torch.uint8, device=dev)
        self.table_nbytes, self.state_nbytes = tb.value, sb.value
        stream = torch.cuda.current_stream(dev).cuda_stream
        if tables_only:
            self._state = None
            self._check(self.lib.pyqed_heom_bind(self._h, self._tables.data_ptr(), tb.value, None, 0,
                                                 C.c_void_p(stream)))
        else:
            self._state = state if state is not None else torch.empty(sb.value, dtype=torch.uint8, device=dev)
            assert self._state.numel() >= sb.value and self._state.dtype == torch.uint8
            self._check(self.lib.pyqed_heom_bind(self._h, self._tables.data_ptr(), tb.value,
                                                 self._state.data_ptr(), sb.value, C.c_void_p(stream)))
        self._check(self.lib.pyqed_heom_build_hierarchy(self._h))

    # -- state -------------------------------------------------------------
    def set_state(self, rho0):
        rho0 = _c128(rho0, (self.batch, self.nsys, self.nsys))
        self._check(self.lib.pyqed_heom_set_state(self._h, _dptr(rho0.view(np.float64))))

    def load_ados(self, ados):
        ados = _c128(ados, (self.batch, self.nmax, self.nsys, self.nsys))
        self._check(self.lib.pyqed_heom_load_ados(self._h, _dptr(ados.view(np.float64))))

    def get_ados(self):
        out = np.empty((self.batch, self.nmax, self.nsys, self.nsys), dtype=np.complex128)
        self._check(self.lib.pyqed_heom_get_ados(self._h, _dptr(out.view(np.float64))))
        return out

    def get_keys(self):
        out = np.empty((self.nmax, self.nind), dtype=np.uint8)
        self._check(self.lib.pyqed_heom_get_keys(self._h, out.ctypes.data_as(_c_uint8_p)))
        return out

    # -- propagation -------------------------------------------------------
    def propagate(self, dt, nt, fsys=None, fcoup=None, traj=None, method=0):
        """``traj``: torch complex128 tensor [batch, nt+1, N, N] on the plan's
        device, or None.  ``fsys`` / ``fcoup``: float64 [batch, nt, 3] or None."""
        def field(f):
            if f is None:
                return None
            f = np.ascontiguousarray(np.broadcast_to(np.asarray(f, dtype=np.float64),
                                                     (self.batch, nt, 3)))
            self._keep.append(f)
            return f
        self._keep = []
        fs, fc = field(fsys), field(fcoup)
        tp = None
        if traj is not None:
            assert traj.is_cuda and traj.is_contiguous() and traj.dtype == self.torch.complex128
            assert tuple(traj.shape) == (self.batch, nt + 1, self.nsys, self.nsys)
            tp = C.c_void_p(traj.data_ptr())
        self._check(self.lib.pyqed_heom_propagate(self._h, float(dt), int(nt), _dptr(fs), _dptr(fc),
                                                  tp, int(method)))

    def expectation(self, rho, ops):
        """``rho``: torch complex128 [batch, npts, N, N] (device); ``ops``:
        [n_ops, N, N] host.  Returns torch complex128 [batch, n_ops, npts]."""
        torch = self.torch
        ops = _c128(ops)
        ops = ops.reshape(-1, self.nsys, self.nsys)
        npts = rho.shape[1]
        out = torch.empty((self.batch, ops.shape[0], npts), dtype=torch.complex128, device=rho.device)
        self._check(self.lib.pyqed_heom_expectation(self._h, C.c_void_p(rho.data_ptr()), npts,
                                                    _dptr(ops.view(np.float64)), ops.shape[0],
                                                    C.c_void_p(out.data_ptr())))
        return out

    def apply_operator(self, op, side="left"):
        """All ADOs <- op @ ADO (``left``) or ADO @ op (``right``), in place."""
        op = _c128(op, (self.nsys, self.nsys))
        self._check(self.lib.pyqed_heom_apply_operator(self._h, _dptr(op.view(np.float64)),
                                                       {"left": 0, "right": 1}[side]))

    def synchronize(self):
        self._check(self.lib.pyqed_heom_synchronize(self._h))

    def launch_count(self):
        return int(self.lib.pyqed_heom_launch_count(self._h))

    def stage_timing(self, enable):
        """Returns ``(total_ms, launches)`` accumulated since the last call and
        switches per-launch CUDA-event timing on or off."""
        ms, n = C.c_double(), C.c_int64()
        self._check(self.lib.pyqed_heom_stage_timing(self._h, int(bool(enable)), C.byref(ms),
                                                     C.byref(n)))
        return ms.value, n.value


def chain_euler(H, S, ado0, gamma, c, dt, nt, e_ops=None, double_update0=False, device=0):
    """Explicit-Euler chain HEOM on the GPU (``pyqed_heom_chain_euler``).

    ``ado0``: complex [batch, nado, N, N] initial ADOs.  Returns
    ``(ado_final[batch, nado, N, N], obs[batch, n_e, nt] or None)``."""
    import torch
    if not torch.cuda.is_available():
        raise HeomError("no CUDA device: the HEOM path has no CPU fallback")
    lib = load()
    ado0 = _c128(ado0)
    batch, nado, n, _ = ado0.shape
    H, S = _c128(H, (n, n)), _c128(S, (n, n))
    dev = torch.device("cuda", device)
    d_ado = torch.from_numpy(ado0).to(dev)
    n_e = 0 if e_ops is None else len(e_ops)
    ops = None if n_e == 0 else _c128(np.stack([np.asarray(e) for e in e_ops]), (n_e, n, n))
    d_obs = None if n_e == 0 else torch.empty((batch, n_e, nt), dtype=torch.complex128, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.pyqed_heom_chain_euler(device, C.c_void_p(stream), n, nado, batch, _dptr(H.view(np.float64)),
                                    _dptr(S.view(np.float64)), float(gamma), float(np.real(c)),
                                    float(np.imag(c)), float(dt), int(nt), int(bool(double_update0)),
                                    C.c_void_p(d_ado.data_ptr()),
                                    None if ops is None else _dptr(ops.view(np.float64)), n_e,
                                    None if d_obs is None else C.c_void_p(d_obs.data_ptr()))
    if rc != 0:
        raise HeomError(lib.pyqed_heom_last_error().decode())
    return d_ado.cpu().numpy(), (None if d_obs is None else d_obs.cpu().numpy())
