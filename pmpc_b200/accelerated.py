"""Momentum-accelerated SCP (reference ``pmpc/accelerated.py:15-116``).

The outer loop is host-side: every iteration extrapolates the linearisation point from the last two iterates
(over-relaxation with factor ``ALF = 1.6``, ``accelerated.py:10-12``) and runs ONE sub-problem through
:func:`pmpc_b200.scp_solve` (``max_it=1``), i.e. one launch of the SCP kernel for built-in / registered dynamics;
the device handle is kept between iterations, so no problem data is re-allocated.
"""
from __future__ import annotations

import math
import time
from typing import Any, Dict, Optional

import numpy as np

from .scp_mpc import print_fn, scp_solve
from .utils import TablePrinter

ALF = 1.6


def momentum_update(zk: np.ndarray, zkm1: np.ndarray, it: int) -> np.ndarray:
    return ALF * zk + (1.0 - ALF) * zkm1


def accelerated_scp_solve(f_fx_fu_fn, Q, R, x0, X_ref=None, U_ref=None, X_prev=None, U_prev=None, x_l=None, x_u=None,
                          u_l=None, u_u=None, verbose=True, debug=False, max_it=100, time_limit=1000.0, res_tol=1e-5,
                          reg_x=1e0, reg_u=1e-2, slew_rate=0.0, u_slew=None, cost_fn=None, method="lqp",
                          solver_settings: Optional[Dict[str, Any]] = None, solver_state=None):
    """Same arguments and return values as the reference function; particle-form inputs only
    (``x0 (M, xdim)``, ``Q (M, N, xdim, xdim)``, ``R (M, N, udim, udim)``), as there."""
    x0, Q, R = np.asarray(x0, np.float64), np.asarray(Q, np.float64), np.asarray(R, np.float64)
    assert x0.ndim == 2 and R.ndim == 4 and Q.ndim == 4
    M, N, xdim = Q.shape[:3]
    udim = R.shape[-1]
    X_ref = np.zeros((M, N, xdim)) if X_ref is None else np.asarray(X_ref, np.float64).reshape(M, N, xdim)
    U_ref = np.zeros((M, N, udim)) if U_ref is None else np.asarray(U_ref, np.float64).reshape(M, N, udim)
    Xh = [X_ref if X_prev is None else np.asarray(X_prev, np.float64).reshape(M, N, xdim)] * 2
    Uh = [U_ref if U_prev is None else np.asarray(U_prev, np.float64).reshape(M, N, udim)] * 2
    settings = dict(solver_settings or {})
    names = ["it", "elaps", "obj", "resid", "reg_x", "reg_u"]
    tp = TablePrinter(names, fmts=["%04d", "%8.3e", "%8.3e", "%8.3e", "%8.3e", "%8.3e"])
    t0 = time.time()
    data: Dict[str, Any] = {}
    X = U = None
    if verbose:
        print_fn(tp.make_header())
    for it in range(max_it):
        Xp, Up = momentum_update(Xh[-1], Xh[-2], it), momentum_update(Uh[-1], Uh[-2], it)
        X, U, d = scp_solve(f_fx_fu_fn, Q, R, x0, X_ref=X_ref, U_ref=U_ref, X_prev=Xp, U_prev=Up, x_l=x_l, x_u=x_u, u_l=u_l,
                            u_u=u_u, verbose=False, debug=debug, max_it=1, time_limit=math.inf, res_tol=0.0, reg_x=reg_x,
                            reg_u=reg_u, slew_rate=slew_rate, u0_slew=u_slew, cost_fn=cost_fn, solver_settings=settings,
                            solver_state=solver_state)
        if X is None:
            return None, None, None
        Xh, Uh = [Xh[-1], X[..., 1:, :]], [Uh[-1], U]
        if "device_solver" in d:  # keep the HBM-resident handle for the next iteration
            settings["_device_solver"] = d.pop("device_solver")
        d.pop("device", None)
        solver_state = (d.get("solver_data") or [{}])[-1].get("solver_state", None)
        for k, v in d.items():
            data.setdefault(k, []).extend(v)
        if verbose:
            print_fn(tp.make_values([it + 1, time.time() - t0] + [d["hist"][-1][k] for k in names[2:]]))
        if data["hist"][-1]["resid"] < res_tol:
            break
        if (it + 2) / (it + 1) * (time.time() - t0) > time_limit:
            break
    if verbose:
        print_fn(tp.make_footer())
    return X, U, data
