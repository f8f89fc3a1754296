"""SCP outer loop restated.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Follows `pmpc/scp_mpc.py:205-442` (``scp_solve``) and the marshalling semantics of
`pmpc/static_backend.py:198-312` (``aff_solve``): defaults (:311-321), the shifted
linearisation point ``X_ = [x0, X_prev[:-1]]`` (:338), NaN failure (:391-394), the
residual / objective bookkeeping (:397-405), warm-start update (:407) and the stop
rule (:424-428).  Filters, ``lin_cost_fn`` and ``extra_cstrs_fns`` are not restated
(optional, host-side).

The QP of each iteration is solved by ``oracle.lqp`` (exact) or ``oracle.osqp_port``
(the reference's solver restated at its default settings).
Pinned against the reference's real loop by ``tests/golden/scp_*.npz``.
"""
from __future__ import annotations

import math
import time
from typing import Any, Callable, Dict, Optional

import numpy as np

from .lqp import lqp_solve as _lqp_exact
from .osqp_port import lqp_solve_osqp as _lqp_osqp


def aff_solve(
    f, fx, fu, x0, X_prev, U_prev, Q, R, X_ref, U_ref, reg_x, reg_u, slew_rate, u_slew,
    x_l, x_u, u_l, u_u, solver_settings=None, qp="exact", osqp_settings=None, lin_rows=None,
):
    """static_backend.py:198-312 semantics on Python-layout arrays; returns (X_traj, U, info)."""
    solver_settings = dict(solver_settings or {})
    M, N, n = f.shape
    Nc = solver_settings.get("Nc", -1)  # static_backend.py:257
    absent = lambda z: z is None or np.size(z) == 0
    kw = dict(
        lx=None if absent(x_l) else x_l,
        ux=None if absent(x_u) else x_u,
        lu=None if absent(u_l) else u_l,
        uu=None if absent(u_u) else u_u,
        reg_x=reg_x,
        reg_u=reg_u,
        # slew_reg <- slew_rate kwarg (:262); slew_reg0 <- solver_settings["slew_reg"] (:263-267)
        slew_reg=None if slew_rate is None else np.full(M, float(slew_rate)),
        slew_reg0=(np.full(M, float(solver_settings["slew_reg"])) if "slew_reg" in solver_settings else None),
        slew_um1=None if u_slew is None else np.asarray(u_slew, float).reshape(M, -1),
        Nc=Nc,
    )
    args = (x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref)
    if lin_rows is not None:
        assert callable(qp) or qp == "exact"
        kw["lin_rows"] = lin_rows
    if callable(qp):  # another treatment of the same sub-problem (oracle.barrier: the cone route's smoothed box rows)
        X, U, info = qp(*args, **kw)
    elif qp == "exact":
        X, U, info = _lqp_exact(*args, **kw)
    else:
        X, U, info = _lqp_osqp(*args, osqp_settings=osqp_settings, **kw)
    X_traj = np.concatenate([np.asarray(x0, float).reshape(M, 1, n), X], -2)
    return X_traj, U, info


def scp_solve(
    f_fx_fu_fn: Callable,
    Q,
    R,
    x0,
    X_ref=None,
    U_ref=None,
    X_prev=None,
    U_prev=None,
    x_l=None,
    x_u=None,
    u_l=None,
    u_u=None,
    verbose: bool = False,
    max_it: int = 100,
    time_limit: float = 1000.0,
    res_tol: float = 1e-5,
    reg_x: float = 1e0,
    reg_u: float = 1e-2,
    slew_rate: Optional[float] = 0.0,
    u0_slew=None,
    solver_settings: Optional[Dict[str, Any]] = None,
    qp: str = "exact",
    osqp_settings=None,
    lin_rows=None,
):
    t0 = time.time()
    x0, reg_x, reg_u = np.array(x0, float), float(reg_x), float(reg_u)
    Q, R = np.array(Q, float), np.array(R, float)
    single = x0.ndim == 1
    if single:  # scp_mpc.py:298-305
        nd = lambda z, d: None if z is None else np.asarray(z, float).reshape((1,) * max(d - np.ndim(z), 0) + np.shape(z))
        Q, R, x0 = nd(Q, 4), nd(R, 4), nd(x0, 2)
        X_ref, U_ref, X_prev, U_prev = [nd(z, 3) for z in (X_ref, U_ref, X_prev, U_prev)]
        x_l, x_u, u_l, u_u = [nd(z, 3) for z in (x_l, x_u, u_l, u_u)]
    M, N, n = Q.shape[:3]
    m = R.shape[-1]
    X_ref = np.zeros((M, N, n)) if X_ref is None else np.array(X_ref, float).reshape(M, N, n)
    U_ref = np.zeros((M, N, m)) if U_ref is None else np.array(U_ref, float).reshape(M, N, m)
    X_prev = X_ref if X_prev is None else np.array(X_prev, float).reshape(M, N, n)  # :313
    U_prev = U_ref if U_prev is None else np.array(U_prev, float).reshape(M, N, m)
    slew_rate = slew_rate if slew_rate is None else float(slew_rate)
    data = dict(hist=[], t_aff_solve=[], solver_data=[])
    X = U = None
    it, max_res = 0, math.inf
    while it < max_it:
        X_ = np.concatenate([x0[:, None, :], X_prev[:, :-1, :]], -2)  # :338
        f, fx, fu = f_fx_fu_fn(X_, U_prev)
        f = np.asarray(f, float).reshape(M, N, n)
        fx = np.asarray(fx, float).reshape(M, N, n, n)
        fu = np.asarray(fu, float).reshape(M, N, n, m)
        t1 = time.time()
        X, U, info = aff_solve(
            f, fx, fu, x0, X_prev, U_prev, Q, R, X_ref, U_ref, reg_x, reg_u, slew_rate, u0_slew,
            x_l, x_u, u_l, u_u, solver_settings, qp=qp, osqp_settings=osqp_settings, lin_rows=lin_rows,
        )
        t1 = time.time() - t1
        if np.isnan(X).any() or np.isnan(U).any():  # :391-394
            return None, None, None
        Xn = X[:, 1:, :]
        dX, dU = Xn - X_prev, U - U_prev
        max_res = max(np.linalg.norm(dX, 2, -1).max(), np.linalg.norm(dU, 2, -1).max())  # :402-403
        dX, dU = Xn - X_ref, U - U_ref
        obj = (
            np.sum(dX * np.einsum("ijrt,ijt->ijr", Q, dX)) + np.sum(dU * np.einsum("ijrt,ijt->ijr", R, dU))
        ) / N / M  # :404-405 (no 1/2, mean over N and M)
        X_prev, U_prev = Xn, U
        data["hist"].append(dict(it=it + 1, elaps=time.time() - t0, obj=obj, resid=max_res, reg_x=reg_x, reg_u=reg_u))
        data["t_aff_solve"].append(t1)
        data["solver_data"].append(info)
        if verbose:
            print("%04d  %8.3e  %8.3e  %8.3e" % (it + 1, time.time() - t0, obj, max_res))
        if max_res < res_tol:
            break
        it += 1
        if (time.time() - t0) * (it + 1) / it > time_limit:
            break
    if single:
        return X.reshape(N + 1, n), U.reshape(N, m), data
    return X, U, data
