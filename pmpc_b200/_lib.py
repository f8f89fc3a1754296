"""ctypes binding of ``libpmpc_b200.so`` (C ABI in ``include/pmpc_b200.h``).

Replaces the reference's ``pmpc/import_pmpcjl.py:30-38`` (which preloads the AOT Julia
libraries and imports the pybind11 ``pmpcjl`` extension).  There is **no CPU fallback**: if the
shared library is missing, or no CUDA device is present at solve time, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMPC_B200_LIB", os.path.join(_HERE, "libpmpc_b200.so"))  # override: A/B builds

c_double_p = C.POINTER(C.c_double)
c_ll_p = C.POINTER(C.c_longlong)


class Settings(C.Structure):
    """Mirror of ``pmpc_settings`` (include/pmpc_b200.h)."""

    _fields_ = [
        ("rho0", C.c_double),
        ("alpha", C.c_double),
        ("eps_abs", C.c_double),
        ("eps_rel", C.c_double),
        ("kkt_tol", C.c_double),
        ("adapt_tol", C.c_double),
        ("max_admm_iter", C.c_int),
        ("check_every", C.c_int),
        ("polish", C.c_int),
        ("polish_iters", C.c_int),
        ("adapt_rho", C.c_int),
        ("warm_start", C.c_int),
        ("fused", C.c_int),
        ("ring", C.c_int),
    ]


class PmpcError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None

_DP = c_double_p
_SIG = {
    "pmpc_default_settings": (None, [C.POINTER(Settings)]),
    "pmpc_last_error": (C.c_char_p, []),
    "pmpc_version": (C.c_char_p, []),
    "pmpc_device_count": (C.c_int, []),
    "pmpc_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                              C.c_size_t, C.c_longlong]),
    "pmpc_destroy": (C.c_int, [C.c_void_p]),
    "pmpc_set_settings": (C.c_int, [C.c_void_p, C.POINTER(Settings)]),
    "pmpc_set_problem": (C.c_int, [C.c_void_p] + [_DP] * 9 + [C.c_double, C.c_double] + [_DP] * 3),
    "pmpc_set_warm_start": (C.c_int, [C.c_void_p, _DP, _DP]),
    "pmpc_set_layout": (C.c_int, [C.c_void_p, C.c_int]),
    "pmpc_set_broadcast": (C.c_int, [C.c_void_p, C.c_uint, C.c_uint]),
    "pmpc_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "pmpc_host_free": (C.c_int, [C.c_void_p]),
    "pmpc_set_dynamics": (C.c_int, [C.c_void_p, C.c_int, _DP, C.c_size_t, C.c_double]),
    "pmpc_set_linearization": (C.c_int, [C.c_void_p, _DP, _DP, _DP]),
    "pmpc_set_ltv": (C.c_int, [C.c_void_p, _DP, _DP, _DP]),
    "pmpc_register_dynamics": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, _DP, C.c_size_t]),
    "pmpc_linearize": (C.c_int, [C.c_void_p]),
    "pmpc_get_linearization": (C.c_int, [C.c_void_p, _DP, _DP, _DP]),
    "pmpc_qp_solve": (C.c_int, [C.c_void_p, C.c_int]),
    "pmpc_scp_solve": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "pmpc_scp_solve_async": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "pmpc_sync": (C.c_int, [C.c_void_p]),
    "pmpc_get_solution": (C.c_int, [C.c_void_p, _DP, _DP]),
    "pmpc_get_trajectory": (C.c_int, [C.c_void_p, _DP, _DP]),
    "pmpc_get_info": (C.c_int, [C.c_void_p, c_ll_p]),
    "pmpc_get_hist": (C.c_int, [C.c_void_p, _DP, C.c_int]),
    "pmpc_last_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "pmpc_timer_start": (C.c_int, [C.c_void_p]),
    "pmpc_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "pmpc_comm_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pmpc_comm_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "pmpc_comm_connect_local": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_size_t]),
    "pmpc_comm_set_timeout": (C.c_int, [C.c_void_p, C.c_double]),
    "pmpc_launch_count": (C.c_longlong, []),
    "pmpc_get_occupancy": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pmpc_microbench": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "pmpc_lqp_solve": (C.c_int, [_DP, _DP, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_longlong] + [_DP] * 14
                       + [C.c_double, C.c_double] + [_DP] * 3 + [C.POINTER(Settings), c_ll_p]),
    "init_julia": (None, [C.c_int, C.c_void_p]),
    "shutdown_julia": (None, [C.c_int]),
    "c_lqp_solve": (None, [_DP, _DP, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_longlong] + [_DP] * 14
                    + [C.c_double, C.c_double] + [_DP] * 3 + [C.c_longlong]),
    "c_lcone_solve": (None, [_DP, _DP, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_longlong] + [_DP] * 14
                      + [C.c_double, C.c_double] + [_DP] * 3 + [C.c_longlong, C.c_double, C.c_char_p]),
}
EXPORTED_SYMBOLS = tuple(_SIG.keys())


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PmpcError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C pmpc_b200/csrc`.  pmpc_b200 has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIG.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().pmpc_last_error().decode()
        raise PmpcError(f"pmpc_b200 {what} failed (code {rc}): {msg}")


def dptr(a: Optional[np.ndarray]):
    """C pointer of a float64 C-contiguous array, or NULL."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(c_double_p)


def f64c(a) -> Optional[np.ndarray]:
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def mat_cm(a) -> Optional[np.ndarray]:
    """(..., rows, cols) Python matrix stack -> the reference C-ABI memory layout (column-major
    blocks), i.e. what ``np.asfortranarray(py2jl(a, 2))`` lays out (static_backend.py:71-74)."""
    return None if a is None else np.ascontiguousarray(np.swapaxes(np.asarray(a, dtype=np.float64), -1, -2))


def default_settings(**overrides) -> Settings:
    s = Settings()
    load().pmpc_default_settings(C.byref(s))
    for k, v in overrides.items():
        if not hasattr(s, k):
            raise KeyError(f"unknown solver setting {k!r}")
        setattr(s, k, v)
    return s


def device_count() -> int:
    return int(load().pmpc_device_count())


class _Pinned:
    """Owner of one cudaHostAlloc block; freed when the last array view on it dies."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        check(load().pmpc_host_alloc(C.byref(self.ptr), max(int(nbytes), 8)), "host_alloc")
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                load().pmpc_host_free(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """NumPy array in page-locked host memory (``pmpc_host_alloc``): H2D / D2H copies of the handle API
    from such arrays run at full PCIe speed.  The memory lives as long as the array (or any view of it)."""
    shape = tuple(int(x) for x in np.atleast_1d(shape))
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    owner = _Pinned(n)
    buf = (C.c_char * max(n, 8)).from_address(owner.ptr.value)
    buf._owner = owner  # keep-alive chain: array -> ctypes buffer -> owner
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def pinned_copy(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    out = pinned_empty(a.shape)
    out[...] = a
    return out
