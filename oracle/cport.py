"""ctypes binding of the plain-C oracle (``oracle/c/pmpc_oracle.c``).  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Same restatement as ``oracle/lqp.py`` + ``oracle/osqp_port.py`` + ``oracle/dubins.py`` + ``oracle/scp.py`` (reference
``PMPC.jl/src/lqp_utils.jl:2-423``, OSQP 0.6.2 published algorithm, ``tests/dubins_car.py:48-90``,
``pmpc/scp_mpc.py:337-428``), compiled, so that the CPU baseline beside the GPU numbers is not a Python-speed one and can
use every host core (one problem per thread, like the reference's worker pool ``pmpc/remote.py``).
Build: ``make -C oracle/c`` (done by ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "c", "libpmpc_oracle.so")
_DP = C.POINTER(C.c_double)
_lib = None


class Settings(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("rho", "sigma", "alpha", "eps_abs", "eps_rel", "adaptive_rho_tolerance",
                                          "adaptive_rho_fraction")] + \
               [(k, C.c_int) for k in ("max_iter", "scaling", "adaptive_rho", "adaptive_rho_interval", "check_termination")]


class Info(C.Structure):
    _fields_ = [("iters", C.c_int), ("status", C.c_int), ("rho_updates", C.c_int), ("r_prim", C.c_double),
                ("r_dual", C.c_double), ("setup_time", C.c_double), ("solve_time", C.c_double)]


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle/c`")
        _lib = C.CDLL(LIB_PATH)
        _lib.pmpc_oracle_lqp_solve.restype = C.c_int
        _lib.pmpc_oracle_scp_dubins.restype = C.c_int
        _lib.pmpc_oracle_scp_dubins_batch.restype = C.c_int
    return _lib


def settings(**kw) -> Settings:
    s = Settings()
    load().pmpc_oracle_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise KeyError(k)
        setattr(s, k, v)
    return s


def _p(a):
    return None if a is None else a.ctypes.data_as(_DP)


def _c(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a if shape is None else np.ascontiguousarray(np.broadcast_to(a, shape))


def lqp_solve(x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx=None, ux=None, lu=None, uu=None, reg_x=0.0, reg_u=0.0,
              slew_reg=None, slew_reg0=None, slew_um1=None, Nc=-1, osqp_settings=None):
    """One sub-problem (`c_lqp_solve`, c_interface.jl:77-141) with the reference's solver restated.  Python layout in,
    X (M,N,n) without x0 and U (M,N,m) out, plus the OSQP statistics."""
    f = _c(f)
    M, N, n = f.shape
    m = np.asarray(fu).shape[-1]
    arrs = [_c(x0), f, _c(fx), _c(fu), _c(X_prev), _c(U_prev), _c(Q), _c(R), _c(X_ref), _c(U_ref), _c(lx), _c(ux), _c(lu), _c(uu)]
    sl = [_c(slew_reg, (M,)) if slew_reg is not None else None, _c(slew_reg0, (M,)) if slew_reg0 is not None else None,
          _c(slew_um1, (M, m)) if slew_um1 is not None else None]
    X, U = np.empty((M, N, n)), np.empty((M, N, m))
    st = settings(**(osqp_settings or {}))
    info = Info()
    load().pmpc_oracle_lqp_solve(_p(X), _p(U), C.c_int(n), C.c_int(m), C.c_int(N), C.c_int(M), C.c_longlong(int(Nc)),
                                 *[_p(a) for a in arrs], C.c_double(float(reg_x)), C.c_double(float(reg_u)),
                                 *[_p(a) for a in sl], C.byref(st), C.byref(info))
    return X, U, info


def scp_dubins(p, Q, R, x0, X_ref=None, U_ref=None, X_prev=None, U_prev=None, u_l=None, u_u=None, reg_x=1.0, reg_u=1e-2,
               slew_rate=0.0, max_it=100, res_tol=1e-5, Nc=-1, eps=1e-6, osqp_settings=None):
    """`pmpc.solve` for Dubins particles entirely in C (linearise, assemble, OSQP, bookkeeping).  Returns X (M,N+1,4),
    U (M,N,2), hist (its, 3) = obj, resid, OSQP iterations."""
    Q = _c(Q)
    M, N = Q.shape[:2]
    z = lambda a, d: _c(np.zeros((M, N, d)) if a is None else a, (M, N, d))
    X_ref, U_ref = z(X_ref, 4), z(U_ref, 2)
    X, U, hist = np.full((M, N + 1, 4), np.nan), np.full((M, N, 2), np.nan), np.zeros((max_it, 3))
    sr = None if not slew_rate else _c(np.full(M, float(slew_rate)))
    st = settings(**(osqp_settings or {}))
    arrs = [_c(p, (M, 3))]
    rest = [_c(x0, (M, 4)), Q, _c(R), X_ref, U_ref, None if X_prev is None else _c(X_prev, (M, N, 4)),
            None if U_prev is None else _c(U_prev, (M, N, 2)), None if u_l is None else _c(u_l, (M, N, 2)),
            None if u_u is None else _c(u_u, (M, N, 2))]
    its = load().pmpc_oracle_scp_dubins(_p(X), _p(U), _p(hist), C.c_int(N), C.c_int(M), C.c_longlong(int(Nc)), _p(arrs[0]),
                                        C.c_double(eps), *[_p(a) for a in rest], C.c_double(reg_x), C.c_double(reg_u), _p(sr),
                                        C.c_int(max_it), C.c_double(res_tol), C.byref(st))
    if its < 0:
        return None, None, None
    return X, U, hist[:its]


def scp_dubins_batch(nthreads, p, Q, R, x0, X_ref, U_ref, X_prev, U_prev, u_l, u_u, reg_x=1.0, reg_u=1e-2, max_it=100,
                     res_tol=1e-5, Nc=-1, eps=1e-6, osqp_settings=None):
    """B independent problems (leading axis B on every array) on ``nthreads`` host threads, one problem per thread at a
    time.  Returns X (B,M,N+1,4), U (B,M,N,2), iters (B,)."""
    Q = _c(Q)
    B, M, N = Q.shape[:3]
    X, U = np.full((B, M, N + 1, 4), np.nan), np.full((B, M, N, 2), np.nan)
    iters = np.zeros(B, dtype=np.int32)
    st = settings(**(osqp_settings or {}))
    arrs = [_c(p, (B, M, 3))]
    rest = [_c(x0, (B, M, 4)), Q, _c(R), _c(X_ref, (B, M, N, 4)), _c(U_ref, (B, M, N, 2)),
            None if X_prev is None else _c(X_prev, (B, M, N, 4)), None if U_prev is None else _c(U_prev, (B, M, N, 2)),
            None if u_l is None else _c(u_l, (B, M, N, 2)), None if u_u is None else _c(u_u, (B, M, N, 2))]
    load().pmpc_oracle_scp_dubins_batch(C.c_int(nthreads), C.c_int(B), _p(X), _p(U), iters.ctypes.data_as(C.POINTER(C.c_int)),
                                        C.c_int(N), C.c_int(M), C.c_longlong(int(Nc)), _p(arrs[0]), C.c_double(eps),
                                        *[_p(a) for a in rest], C.c_double(reg_x), C.c_double(reg_u), C.c_int(max_it),
                                        C.c_double(res_tol), C.byref(st))
    return X, U, iters
