"""Drop-in for the reference's ``pmpcjl`` pybind11 extension (``PMPC.jl/pmpcjl/module.cpp:28-120``).

Same Python-level signatures, same flat return values, but the work is done by
``libpmpc_b200.so`` (CUDA) instead of the embedded Julia runtime.  Arrays are passed exactly as
the reference passes them: Fortran-ordered ``(n,N,M)`` / ``(n,n,N,M)`` (static_backend.py:71-74);
callers that follow the reference are zero-copy, anything else is converted to that order first.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _f(a):
    """Fortran-contiguous float64 array -> (keep-alive, pointer): the memory the C ABI expects."""
    a = np.asfortranarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_lib.c_double_p)


def _call(fn_name, xdim, udim, N, M, Nc, arrays, reg_x, reg_u, slews, verbose, extra=()):
    lib = _lib.load()
    X = np.empty(int(xdim) * int(N) * int(M))
    U = np.empty(int(udim) * int(N) * int(M))
    keep = [_f(a) for a in arrays]
    keep_s = [_f(a) for a in slews]
    getattr(lib, fn_name)(
        X.ctypes.data_as(_lib.c_double_p), U.ctypes.data_as(_lib.c_double_p), int(xdim), int(udim), int(N), int(M),
        int(Nc), *[k[1] for k in keep], float(reg_x), float(reg_u), *[k[1] for k in keep_s], int(verbose), *extra)
    return X, U


def lqp_solve(xdim, udim, N, M, Nc, x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx, ux, lu, uu, reg_x, reg_u,
              slew_reg, slew_reg0, slew_um1, verbose):
    """module.cpp:28-72 — returns flat ``(X_out, U_out)`` in ``(n,N,M)`` column-major order."""
    return _call("c_lqp_solve", xdim, udim, N, M, Nc, [x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx, ux, lu, uu],
                 reg_x, reg_u, [slew_reg, slew_reg0, slew_um1], verbose)


def lcone_solve(xdim, udim, N, M, Nc, x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx, ux, lu, uu, reg_x, reg_u,
                slew_reg, slew_reg0, slew_um1, verbose, smooth_alpha, solver):
    """module.cpp:74-120."""
    return _call("c_lcone_solve", xdim, udim, N, M, Nc,
                 [x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx, ux, lu, uu], reg_x, reg_u,
                 [slew_reg, slew_reg0, slew_um1], verbose, extra=(float(smooth_alpha), str(solver).encode()))
