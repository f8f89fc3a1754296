"""Momentum-accelerated SCP outer loop — drop-in for reference ``pmpc/accelerated.py:15-116``.

Each outer iteration extrapolates the linearisation point from the last two iterates and runs exactly ONE convex
sub-problem through :func:`pmpc_b200.scp_solve` (``max_it=1``): for built-in / registered dynamics that is one launch
of the SCP kernel on a handle that stays resident in HBM between iterations.
"""
from __future__ import annotations

import math
import time
from collections import deque
from typing import Any, Dict, Optional

import numpy as np

from .scp_mpc import print_fn, scp_solve
from .utils import TablePrinter

ALF = 1.6  # over-relaxation factor of the extrapolation (reference accelerated.py:10)
_COLS = ("it", "elaps", "obj", "resid", "reg_x", "reg_u")


def momentum_update(zk: np.ndarray, zkm1: np.ndarray, it: int) -> np.ndarray:
    """z_k + (ALF - 1)(z_k - z_{k-1}); ``it`` is unused, kept for signature compatibility."""
    return zk + (ALF - 1.0) * (zk - zkm1)


class _Extrapolator:
    """Two-deep history of one trajectory array and its extrapolated next linearisation point."""

    def __init__(self, z0: np.ndarray):
        self.hist = deque([z0, z0], maxlen=2)

    def point(self, it: int) -> np.ndarray:
        return momentum_update(self.hist[1], self.hist[0], it)

    def push(self, z: np.ndarray):
        self.hist.append(z)


def accelerated_scp_solve(f_fx_fu_fn, Q, R, x0, X_ref=None, U_ref=None, X_prev=None, U_prev=None, x_l=None, x_u=None,
                          u_l=None, u_u=None, verbose=True, debug=False, max_it=100, time_limit=1000.0, res_tol=1e-5,
                          reg_x=1e0, reg_u=1e-2, slew_rate=0.0, u_slew=None, cost_fn=None, method="lqp",
                          solver_settings: Optional[Dict[str, Any]] = None, solver_state=None):
    """Arguments and return values of the reference function; particle-form inputs only (``x0 (M, xdim)``,
    ``Q (M, N, xdim, xdim)``, ``R (M, N, udim, udim)``), as there.  ``method`` is accepted and ignored (one back end)."""
    x0, Q, R = (np.asarray(a, np.float64) for a in (x0, Q, R))
    if not (x0.ndim == 2 and Q.ndim == 4 and R.ndim == 4):
        raise AssertionError("accelerated_scp_solve needs particle-form inputs")
    M, N, xdim = Q.shape[:3]
    udim = R.shape[-1]
    as_traj = lambda a, d: np.zeros((M, N, d)) if a is None else np.asarray(a, np.float64).reshape(M, N, d)
    X_ref, U_ref = as_traj(X_ref, xdim), as_traj(U_ref, udim)
    ex_X = _Extrapolator(X_ref if X_prev is None else as_traj(X_prev, xdim))
    ex_U = _Extrapolator(U_ref if U_prev is None else as_traj(U_prev, udim))
    fixed = dict(X_ref=X_ref, U_ref=U_ref, x_l=x_l, x_u=x_u, u_l=u_l, u_u=u_u, verbose=False, debug=debug, max_it=1,
                 time_limit=math.inf, res_tol=0.0, reg_x=reg_x, reg_u=reg_u, slew_rate=slew_rate, u0_slew=u_slew, cost_fn=cost_fn)
    settings = dict(solver_settings or {})
    table = TablePrinter(list(_COLS), fmts=["%04d"] + ["%8.3e"] * 5)
    say = print_fn if verbose else (lambda *_: None)
    merged: Dict[str, Any] = {}
    X = U = None
    t0 = time.time()
    say(table.make_header())
    for it in range(max_it):
        X, U, step = scp_solve(f_fx_fu_fn, Q, R, x0, X_prev=ex_X.point(it), U_prev=ex_U.point(it), solver_settings=settings,
                               solver_state=solver_state, **fixed)
        if X is None:  # NaN failure convention of the solve path
            return None, None, None
        ex_X.push(X[..., 1:, :])
        ex_U.push(U)
        handle = step.pop("device_solver", None)
        if handle is not None:
            settings["_device_solver"] = handle  # the next iteration reuses the HBM-resident problem
        step.pop("device", None)
        solver_state = (step.get("solver_data") or [{}])[-1].get("solver_state")
        for key, seq in step.items():
            merged.setdefault(key, []).extend(seq)
        last = merged["hist"][-1]
        elapsed = time.time() - t0
        say(table.make_values([it + 1, elapsed] + [last[c] for c in _COLS[2:]]))
        if last["resid"] < res_tol or elapsed * (it + 2) / (it + 1) > time_limit:
            break
    say(table.make_footer())
    return X, U, merged
